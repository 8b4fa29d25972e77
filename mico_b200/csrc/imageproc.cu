// mico_b200 -- SURVEY 8(f).4 (input side of the path): image / video-frame preprocessing on the GPU.
//
// Replaces torchvision's ToTensor -> Resize((R,R)) -> Normalize(mean, std) chain that model/imageprocessor.py:24-29,52-56 and
// model/videoprocessor.py:36-41 run on the CPU per image / frame:
//     v = u8 / 255                                   (ToTensor)
//     r = bilinear resize, align_corners = False     (Resize on a tensor -> F.interpolate; torchvision >= 0.17 anti-aliases
//                                                     by default, the reference's pinned 0.15.2 does not: both are provided)
//     out[c] = (r[c] - mean[c]) / std[c]             (Normalize), written as fp32 CHW = the tower's input layout
// One thread per output pixel, all channels; reads of the HWC uint8 source are contiguous along x across a warp.
// Anti-aliased mode follows ATen's separable triangle filter (UpSampleKernel _compute_indices_weights_aa):
//   scale = in/out, support = max(scale, 1), centre = scale * (i + 0.5), taps [xmin, xmin + xsize) with
//   xmin = max(0, int(centre - support + 0.5)), xsize = min(in, int(centre + support + 0.5)) - xmin,
//   w_j = max(0, 1 - |(j + xmin - centre + 0.5) / max(scale, 1)|), normalised to sum 1.
#include "common.cuh"
#include "host_utils.h"

namespace mico {
namespace {

struct ImgNorm { float mean[4], inv_std[4]; };

__device__ __forceinline__ void aa_taps(int i, float scale, int in_size, int& xmin, int& xsize, float& centre, float& inv) {
    const float support = scale >= 1.0f ? scale : 1.0f;
    centre = scale * ((float)i + 0.5f);
    xmin = max(0, (int)(centre - support + 0.5f));
    xsize = min(in_size, (int)(centre + support + 0.5f)) - xmin;
    inv = scale >= 1.0f ? 1.0f / scale : 1.0f;
}
__device__ __forceinline__ float tri(float x) { x = fabsf(x); return x < 1.0f ? 1.0f - x : 0.0f; }

template <bool kU8Hwc>
__device__ __forceinline__ float src_at(const void* src, int64_t img_off, int C, int H, int W, int c, int y, int x) {
    if (kU8Hwc) return (float)reinterpret_cast<const uint8_t*>(src)[img_off + ((int64_t)y * W + x) * C + c] * (1.0f / 255.0f);
    return reinterpret_cast<const float*>(src)[img_off + ((int64_t)c * H + y) * W + x];
}

template <bool kU8Hwc, bool kAA>
__global__ void resize_normalize_kernel(const void* __restrict__ src, int64_t img_stride, int C, int H, int W,
                                        float* __restrict__ dst, int Ho, int Wo, ImgNorm nm) {
    const int ox = blockIdx.x * blockDim.x + threadIdx.x, oy = blockIdx.y, n = blockIdx.z;
    if (ox >= Wo) return;
    const int64_t img_off = (int64_t)n * img_stride;
    const float sx = (float)W / (float)Wo, sy = (float)H / (float)Ho;
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    if (kAA) {
        int x0, xn, y0, yn;
        float cx, cy, ix, iy;
        aa_taps(ox, sx, W, x0, xn, cx, ix);
        aa_taps(oy, sy, H, y0, yn, cy, iy);
        float wxs = 0.f, wys = 0.f;
        for (int j = 0; j < xn; ++j) wxs += tri(((float)(j + x0) - cx + 0.5f) * ix);
        for (int j = 0; j < yn; ++j) wys += tri(((float)(j + y0) - cy + 0.5f) * iy);
        for (int jy = 0; jy < yn; ++jy) {
            const float wy = tri(((float)(jy + y0) - cy + 0.5f) * iy) / wys;
            float row[4] = {0.f, 0.f, 0.f, 0.f};
            for (int jx = 0; jx < xn; ++jx) {
                const float wx = tri(((float)(jx + x0) - cx + 0.5f) * ix) / wxs;
                for (int c = 0; c < C; ++c) row[c] += wx * src_at<kU8Hwc>(src, img_off, C, H, W, c, y0 + jy, x0 + jx);
            }
            for (int c = 0; c < C; ++c) acc[c] += wy * row[c];
        }
    } else {
        // plain bilinear, align_corners = False (ATen area_pixel_compute_source_index: negative sources clamp to 0)
        float fx = sx * ((float)ox + 0.5f) - 0.5f, fy = sy * ((float)oy + 0.5f) - 0.5f;
        fx = fx < 0.f ? 0.f : fx;
        fy = fy < 0.f ? 0.f : fy;
        const int x0 = min((int)fx, W - 1), y0 = min((int)fy, H - 1);
        const int x1 = min(x0 + 1, W - 1), y1 = min(y0 + 1, H - 1);
        const float lx = fx - (float)x0, ly = fy - (float)y0;
        for (int c = 0; c < C; ++c) {
            const float v00 = src_at<kU8Hwc>(src, img_off, C, H, W, c, y0, x0), v01 = src_at<kU8Hwc>(src, img_off, C, H, W, c, y0, x1);
            const float v10 = src_at<kU8Hwc>(src, img_off, C, H, W, c, y1, x0), v11 = src_at<kU8Hwc>(src, img_off, C, H, W, c, y1, x1);
            acc[c] = (1.f - ly) * ((1.f - lx) * v00 + lx * v01) + ly * ((1.f - lx) * v10 + lx * v11);
        }
    }
    for (int c = 0; c < C; ++c)
        dst[(((int64_t)n * C + c) * Ho + oy) * Wo + ox] = (acc[c] - nm.mean[c]) * nm.inv_std[c];
}

}  // namespace
}  // namespace mico

extern "C" int mico_resize_normalize(const void* src, int src_is_u8_hwc, int n, int C, int H, int W, int64_t img_stride,
                                     float* dst, int Ho, int Wo, const float* mean, const float* std, int antialias,
                                     void* stream_) {
    using namespace mico;
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    MICO_CHECK_ARG(src && dst && mean && std);
    MICO_CHECK_ARG(n > 0 && C >= 1 && C <= 4 && H > 0 && W > 0 && Ho > 0 && Wo > 0 && Ho <= 65535 && n <= 65535);
    ImgNorm nm;
    for (int c = 0; c < 4; ++c) {
        nm.mean[c] = c < C ? mean[c] : 0.f;
        nm.inv_std[c] = c < C ? 1.0f / std[c] : 1.f;
    }
    ProfScope prof(kProfOther, (double)n * C * ((double)H * W * (src_is_u8_hwc ? 1 : 4) + (double)Ho * Wo * 4), stream);
    dim3 grid(ceil_div(Wo, 128), Ho, n);
    auto go = [&](auto k) { k<<<grid, 128, 0, stream>>>(src, img_stride, C, H, W, dst, Ho, Wo, nm); };
    if (src_is_u8_hwc) { if (antialias) go(resize_normalize_kernel<true, true>); else go(resize_normalize_kernel<true, false>); }
    else               { if (antialias) go(resize_normalize_kernel<false, true>); else go(resize_normalize_kernel<false, false>); }
    MICO_CHECK_CUDA(cudaGetLastError());
    count_launch();
    return MICO_OK;
}
