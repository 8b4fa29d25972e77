// mico_b200 -- HBM-bound helper kernels of the ViT/BERT path (coalesced 128-bit accesses).
//
//   mico_cast_f32_to_bf16   fp32 master weights -> bf16 GEMM operands (once per optimizer step)
//   mico_colsum_bf16        bias gradients: out[n] = sum_m x[m,n]            (nn.Linear bias backward)
//   mico_batch_sum_f32      d pos_embed / d cls_token: out[r] = sum_b x[b,r] (eva_vit_model.py:615-619 backward)
//   mico_patchify           K1 im2col: (B,C,H,W) fp32 -> bf16 [B*gh*gw, Kpad]   (eva_vit_model.py:440-447)
//   mico_cls_pos_row        token 0 = cls_token + pos_embed[0]                  (eva_vit_model.py:615-619)
//   mico_scale_cast_bf16    bf16(x * row_scale[row/rpg])  (gradient entering a DropPath'd residual branch)
#include <atomic>
#include <curand_kernel.h>

#include "common.cuh"
#include "host_utils.h"

namespace mico {
namespace {

__global__ void cast_f32_bf16_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst, int64_t n) {
    const int64_t nvec = n >> 3;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += stride) {
        const float4 a = reinterpret_cast<const float4*>(src)[2 * i];
        const float4 b = reinterpret_cast<const float4*>(src)[2 * i + 1];
        reinterpret_cast<uint4*>(dst)[i] = make_uint4(pack_bf16x2(a.x, a.y), pack_bf16x2(a.z, a.w),
                                                      pack_bf16x2(b.x, b.y), pack_bf16x2(b.z, b.w));
    }
    // tail
    for (int64_t i = (nvec << 3) + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
        dst[i] = __float2bfloat16(src[i]);
}

// x *= s (s = *scale_dev * scale_host): gradients of a loss group differentiated ahead of time, rescaled by the upstream
// scalar gradient when the outer backward pass reaches them (mico_b200/train_step.py)
__global__ void scale_f32_kernel(float* __restrict__ x, const float* __restrict__ scale_dev, float scale_host, int64_t n) {
    const float s = (scale_dev ? __ldg(scale_dev) : 1.0f) * scale_host;
    const int64_t nvec = n >> 2;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += stride) {
        float4 v = reinterpret_cast<float4*>(x)[i];
        v.x *= s; v.y *= s; v.z *= s; v.w *= s;
        reinterpret_cast<float4*>(x)[i] = v;
    }
    for (int64_t i = (nvec << 2) + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) x[i] *= s;
}

// bf16 -> fp32 (the reduced bf16 gradient bucket back into the fp32 gradient buffer)
__global__ void cast_bf16_f32_kernel(const __nv_bfloat16* __restrict__ src, float* __restrict__ dst, int64_t n) {
    const int64_t nvec = n >> 3;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += stride) {
        const uint4 v = reinterpret_cast<const uint4*>(src)[i];
        reinterpret_cast<float4*>(dst)[2 * i] = make_float4(bf16_lo(v.x), bf16_hi(v.x), bf16_lo(v.y), bf16_hi(v.y));
        reinterpret_cast<float4*>(dst)[2 * i + 1] = make_float4(bf16_lo(v.z), bf16_hi(v.z), bf16_lo(v.w), bf16_hi(v.w));
    }
    for (int64_t i = (nvec << 3) + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
        dst[i] = __bfloat162float(src[i]);
}

// Rotary position embedding of the EVA02 towers (model/evaclip/rope.py:79-136, applied at eva_vit_model.py:314-322): every
// token except the first (cls) is rotated pair-wise, (x1, x2) -> (x1 c - x2 s, x2 c + x1 s) with the interleaved tables
// cos / sin [tokens-1, d] (both entries of a pair hold the same angle).  inverse: the transpose rotation (backward pass).
template <typename TI, typename TO>
__global__ void rope_kernel(const TI* __restrict__ x, TO* __restrict__ y, const float* __restrict__ cosv,
                            const float* __restrict__ sinv, int64_t n_pairs, int T, int H, int d, int inverse) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int hp = d >> 1;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_pairs; i += stride) {
        const int pr = (int)(i % hp);
        const int64_t row = i / hp;                    // (b, t, h)
        const int t = (int)((row / H) % T);
        const float x1 = (float)x[2 * i], x2 = (float)x[2 * i + 1];
        float y1 = x1, y2 = x2;
        if (t > 0) {
            const float c = cosv[(int64_t)(t - 1) * d + 2 * pr];
            float sn = sinv[(int64_t)(t - 1) * d + 2 * pr];
            if (inverse) sn = -sn;
            y1 = x1 * c - x2 * sn;
            y2 = x2 * c + x1 * sn;
        }
        y[2 * i] = (TO)y1;
        y[2 * i + 1] = (TO)y2;
    }
}

// SwiGLU of the EVA02 MLP (eva_vit_model.py:201-224): g = silu(u1) * u2; backward du1 = dg u2 silu'(u1), du2 = dg silu(u1)
__global__ void swiglu_fwd_kernel(const float* __restrict__ u1, const float* __restrict__ u2, float* __restrict__ g, int64_t n) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const float a = u1[i], sg = 1.0f / (1.0f + __expf(-a));
        g[i] = a * sg * u2[i];
    }
}
__global__ void swiglu_bwd_kernel(const float* __restrict__ u1, const float* __restrict__ u2, const float* __restrict__ dg,
                                  float* __restrict__ du1, float* __restrict__ du2, int64_t n) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const float a = u1[i], sg = 1.0f / (1.0f + __expf(-a)), d = dg[i];
        du1[i] = d * u2[i] * sg * (1.0f + a * (1.0f - sg));
        du2[i] = d * a * sg;
    }
}

// rows x cols fp32 (pitch lds) -> bf16 (pitch ldd >= cols); columns cols..ldd-1 are zero-filled.
__global__ void cast_f32_bf16_2d_kernel(const float* __restrict__ src, int64_t lds, int rows, int cols,
                                        __nv_bfloat16* __restrict__ dst, int64_t ldd) {
    const int64_t total = (int64_t)rows * ldd;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = i / ldd;
        const int c = (int)(i - r * ldd);
        dst[i] = __float2bfloat16(c < cols ? src[r * lds + c] : 0.0f);
    }
}

// grid: (ceil(N/256), splits); block 256 = 8 warps; lane owns 8 consecutive columns.
// Single pass: every block writes its partial row, takes a ticket, and the LAST block of a column group sums the partials
// in split order (deterministic) and writes the result -- no finalize launch (ncu round 1: 8 us per finalize, 190 per step).
// Logical column c is physical column c + (c >= gap_start ? gap_len : 0): one launch sums the q and v thirds of the fused
// qkv gradient and skips the bias-free k third (eva_vit_model.py:307); results go to out0 (c < gap_start) / out1.
constexpr int kColsumCols = 256;
constexpr int kColsumSlots = 32, kColsumMaxBlocks = 256;
__device__ unsigned int g_colsum_tickets[kColsumSlots * kColsumMaxBlocks];   // zero at load, reset by the last block

__global__ void __launch_bounds__(256)
colsum_bf16_kernel(const __nv_bfloat16* __restrict__ x, int64_t ldx, int M, int N, int gap_start, int gap_len,
                   float* __restrict__ partials, float* __restrict__ out0, float* __restrict__ out1, int accumulate,
                   unsigned int* __restrict__ tickets) {
    __shared__ float red[8][kColsumCols];
    __shared__ int s_last;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int col = blockIdx.x * kColsumCols + lane * 8;                 // logical
    const int pcol = col + (col >= gap_start ? gap_len : 0);              // physical (gap bounds are multiples of 8)
    const int rows_per_split = (M + gridDim.y - 1) / gridDim.y;
    const int r0 = blockIdx.y * rows_per_split;
    const int r1 = min(M, r0 + rows_per_split);
    float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    if (col + 8 <= N) {
        int r = r0 + warp;
        for (; r + 8 < r1; r += 16) {      // two rows in flight per lane
            const uint4 u = *reinterpret_cast<const uint4*>(x + (int64_t)r * ldx + pcol);
            const uint4 w = *reinterpret_cast<const uint4*>(x + (int64_t)(r + 8) * ldx + pcol);
            acc[0] += bf16_lo(u.x) + bf16_lo(w.x); acc[1] += bf16_hi(u.x) + bf16_hi(w.x);
            acc[2] += bf16_lo(u.y) + bf16_lo(w.y); acc[3] += bf16_hi(u.y) + bf16_hi(w.y);
            acc[4] += bf16_lo(u.z) + bf16_lo(w.z); acc[5] += bf16_hi(u.z) + bf16_hi(w.z);
            acc[6] += bf16_lo(u.w) + bf16_lo(w.w); acc[7] += bf16_hi(u.w) + bf16_hi(w.w);
        }
        for (; r < r1; r += 8) {
            const uint4 u = *reinterpret_cast<const uint4*>(x + (int64_t)r * ldx + pcol);
            acc[0] += bf16_lo(u.x); acc[1] += bf16_hi(u.x); acc[2] += bf16_lo(u.y); acc[3] += bf16_hi(u.y);
            acc[4] += bf16_lo(u.z); acc[5] += bf16_hi(u.z); acc[6] += bf16_lo(u.w); acc[7] += bf16_hi(u.w);
        }
    } else {
        for (int r = r0 + warp; r < r1; r += 8)
            for (int j = 0; j < 8; ++j)
                if (col + j < N) acc[j] += __bfloat162float(x[(int64_t)r * ldx + pcol + j]);
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) red[warp][lane * 8 + j] = acc[j];
    __syncthreads();
    const int c = blockIdx.x * kColsumCols + threadIdx.x;
    if (c < N) {
        float s = 0.f;
#pragma unroll
        for (int w = 0; w < 8; ++w) s += red[w][threadIdx.x];
        partials[(size_t)blockIdx.y * N + c] = s;
    }
    if (tickets == nullptr) return;          // two-kernel path (more column groups than ticket counters)
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) s_last = atomicAdd(&tickets[blockIdx.x], 1u) == gridDim.y - 1;
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    if (c < N) {
        float s = 0.f;
        for (int i = 0; i < (int)gridDim.y; ++i) s += __ldcg(partials + (size_t)i * N + c);
        float* dst = c < gap_start ? out0 + c : out1 + (c - gap_start);
        *dst = accumulate == 1 ? *dst + s : (accumulate == 2 ? *dst - s : s);
    }
    if (threadIdx.x == 0) tickets[blockIdx.x] = 0;
}

__global__ void colsum_finalize_kernel(const float* __restrict__ partials, int splits, int N, float* __restrict__ out,
                                       int accumulate) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= N) return;
    float s = 0.f;
    for (int i = 0; i < splits; ++i) s += partials[(size_t)i * N + c];
    out[c] = accumulate == 1 ? out[c] + s : (accumulate == 2 ? out[c] - s : s);
}

// out[r] (+)= sum_b x[b*R + r]; R % 4 == 0
__global__ void batch_sum_f32_kernel(const float* __restrict__ x, int B, int64_t R, float* __restrict__ out,
                                     int accumulate) {
    const int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
    if (i >= R) return;
    float4 s = make_float4(0, 0, 0, 0);
    for (int b = 0; b < B; ++b) {
        const float4 v = *reinterpret_cast<const float4*>(x + (int64_t)b * R + i);
        s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
    }
    float4* o = reinterpret_cast<float4*>(out + i);
    if (accumulate) { const float4 p = *o; s.x += p.x; s.y += p.y; s.z += p.z; s.w += p.w; }
    *o = s;
}

// One block per (image, patch-row): stage C x P x W pixels through smem, emit gw patch rows of Kpad bf16.
// Column order within a patch row is (c, ky, kx) == Conv2d weight.view(out, C*P*P).
__global__ void __launch_bounds__(256)
patchify_kernel(const float* __restrict__ img, int64_t img_stride, int64_t chan_stride, int C, int H, int W, int P,
                int Kpad, int tokens_per_img, int token_off, __nv_bfloat16* __restrict__ out) {
    extern __shared__ float ps[];   // [C*P][W]
    const int gw = W / P, gh = H / P;
    const int b = blockIdx.x / gh, py = blockIdx.x % gh;
    const float* src = img + (int64_t)b * img_stride;
    const int rows = C * P;
    for (int idx = threadIdx.x; idx < rows * W; idx += blockDim.x) {
        const int rr = idx / W, xx = idx - rr * W;
        const int c = rr / P, ky = rr - c * P;
        ps[idx] = src[(int64_t)c * chan_stride + (int64_t)(py * P + ky) * W + xx];
    }
    __syncthreads();
    const int K = C * P * P;
    __nv_bfloat16* dst = out + ((int64_t)b * tokens_per_img + token_off + (int64_t)py * gw) * Kpad;
    if (py == 0) {   // leading rows of this image (the cls slot) and any trailing slack are zero
        __nv_bfloat16* lead = out + (int64_t)b * tokens_per_img * Kpad;
        for (int idx = threadIdx.x; idx < token_off * (Kpad / 2); idx += blockDim.x)
            reinterpret_cast<uint32_t*>(lead)[idx] = 0u;
        __nv_bfloat16* trail = lead + (int64_t)(token_off + gh * gw) * Kpad;
        for (int idx = threadIdx.x; idx < (tokens_per_img - token_off - gh * gw) * (Kpad / 2); idx += blockDim.x)
            reinterpret_cast<uint32_t*>(trail)[idx] = 0u;
    }
    for (int idx = threadIdx.x; idx < gw * (Kpad / 2); idx += blockDim.x) {
        const int px = idx / (Kpad / 2);
        const int k = (idx - px * (Kpad / 2)) * 2;
        float v0 = 0.f, v1 = 0.f;
        if (k < K) {
            const int rr = k / P, kx = k - rr * P;
            v0 = ps[rr * W + px * P + kx];
        }
        if (k + 1 < K) {
            const int rr = (k + 1) / P, kx = (k + 1) - rr * P;
            v1 = ps[rr * W + px * P + kx];
        }
        *reinterpret_cast<uint32_t*>(dst + (int64_t)px * Kpad + k) = pack_bf16x2(v0, v1);
    }
}

__global__ void cls_pos_row_kernel(const float* __restrict__ cls, const float* __restrict__ pos0, float* __restrict__ x,
                                   int64_t sample_stride, int B, int D) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B * D) return;
    const int b = i / D, d = i - b * D;
    x[(int64_t)b * sample_stride + d] = cls[d] + pos0[d];
}

__global__ void scale_cast_bf16_kernel(const float* __restrict__ x, int64_t ldx, const float* __restrict__ row_scale,
                                       int rows_per_group, __nv_bfloat16* __restrict__ y, int64_t ldy, int M, int D) {
    const int nvec = D >> 2;
    const int64_t total = (int64_t)M * nvec;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int row = (int)(i / nvec), c = (int)(i - (int64_t)row * nvec) * 4;
        const float s = row_scale ? row_scale[row / rows_per_group] : 1.0f;
        const float4 v = *reinterpret_cast<const float4*>(x + (int64_t)row * ldx + c);
        *reinterpret_cast<uint2*>(y + (int64_t)row * ldy + c) =
            make_uint2(pack_bf16x2(v.x * s, v.y * s), pack_bf16x2(v.z * s, v.w * s));
    }
}

// DropPath multipliers for a whole tower in one launch: out[l][j][b] = Bernoulli(1-p_l) / (1-p_l)
// (eva_vit_model.py:121-138, scale_by_keep=True); Philox counter = flat index, so results depend only on (seed, offset).
__global__ void drop_path_scales_kernel(const float* __restrict__ drop_prob, int L, int B, uint64_t seed, uint64_t offset,
                                        float* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= L * 2 * B) return;
    const float p = drop_prob[i / (2 * B)];
    float v = 1.0f;
    if (p > 0.0f) {
        curandStatePhilox4_32_10_t st;
        curand_init(seed, (uint64_t)i, offset, &st);
        const float keep = 1.0f - p;
        const float u = curand_uniform(&st);   // (0, 1]
        v = (u <= keep) ? (keep > 0.0f ? 1.0f / keep : 1.0f) : 0.0f;
    }
    out[i] = v;
}

int colsum_splits(int M, int N) {
    const int colblocks = ceil_div(N, kColsumCols);
    int s = ceil_div(num_sms() * 4, colblocks);
    const int maxs = ceil_div(M, 64);
    if (s > maxs) s = maxs;
    return s < 1 ? 1 : s;
}

}  // namespace
}  // namespace mico

extern "C" int mico_cast_f32_to_bf16(const float* src, void* dst, int64_t n, void* stream_) {
    using namespace mico;
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    ProfScope prof(kProfOther, 6.0 * (double)n, stream);
    MICO_CHECK_ARG(src && dst && n > 0);
    MICO_CHECK_ARG((reinterpret_cast<uintptr_t>(src) & 15) == 0 && (reinterpret_cast<uintptr_t>(dst) & 15) == 0);
    int64_t want = (n / 8 + 255) / 256;
    const int grid = (int)(want < 1 ? 1 : (want > num_sms() * 16 ? num_sms() * 16 : want));
    cast_f32_bf16_kernel<<<grid, 256, 0, stream>>>(src, reinterpret_cast<__nv_bfloat16*>(dst), n);
    MICO_CHECK_CUDA(cudaGetLastError());
    count_launch();
    return MICO_OK;
}

extern "C" int mico_rope(const void* x, int x_is_bf16, void* y, int y_is_bf16, const float* cosv, const float* sinv, int B, int T,
                         int H, int d, int inverse, void* stream_) {
    using namespace mico;
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    MICO_CHECK_ARG(x && y && cosv && sinv && B > 0 && T > 1 && H > 0 && d > 0 && d % 2 == 0);
    MICO_CHECK_ARG(x_is_bf16 != y_is_bf16);       // fp32 -> bf16 (forward: the attention operand) or bf16 -> fp32 (gradient)
    const int64_t n_pairs = (int64_t)B * T * H * (d / 2);
    ProfScope prof(kProfOther, 6.0 * 2.0 * (double)n_pairs, stream);
    int64_t want = (n_pairs + 255) / 256;
    const int grid = (int)(want > num_sms() * 16 ? num_sms() * 16 : want);
    if (x_is_bf16)
        rope_kernel<__nv_bfloat16, float><<<grid, 256, 0, stream>>>(reinterpret_cast<const __nv_bfloat16*>(x),
                                                                   reinterpret_cast<float*>(y), cosv, sinv, n_pairs, T, H, d, inverse);
    else
        rope_kernel<float, __nv_bfloat16><<<grid, 256, 0, stream>>>(reinterpret_cast<const float*>(x),
                                                                   reinterpret_cast<__nv_bfloat16*>(y), cosv, sinv, n_pairs, T, H, d, inverse);
    MICO_CHECK_CUDA(cudaGetLastError());
    count_launch();
    return MICO_OK;
}

extern "C" int mico_swiglu(const float* u1, const float* u2, const float* dg, float* out0, float* out1, int64_t n, void* stream_) {
    using namespace mico;
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    MICO_CHECK_ARG(u1 && u2 && out0 && n > 0 && (dg == nullptr || out1 != nullptr));
    ProfScope prof(kProfOther, (dg ? 20.0 : 12.0) * (double)n, stream);
    int64_t want = (n + 255) / 256;
    const int grid = (int)(want > num_sms() * 16 ? num_sms() * 16 : want);
    if (dg) swiglu_bwd_kernel<<<grid, 256, 0, stream>>>(u1, u2, dg, out0, out1, n);
    else swiglu_fwd_kernel<<<grid, 256, 0, stream>>>(u1, u2, out0, n);
    MICO_CHECK_CUDA(cudaGetLastError());
    count_launch();
    return MICO_OK;
}

extern "C" int mico_scale_f32(float* x, const float* scale_dev, float scale_host, int64_t n, void* stream_) {
    using namespace mico;
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    ProfScope prof(kProfOther, 8.0 * (double)n, stream);
    MICO_CHECK_ARG(x && n > 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0);
    int64_t want = (n / 4 + 255) / 256;
    const int grid = (int)(want < 1 ? 1 : (want > num_sms() * 16 ? num_sms() * 16 : want));
    scale_f32_kernel<<<grid, 256, 0, stream>>>(x, scale_dev, scale_host, n);
    MICO_CHECK_CUDA(cudaGetLastError());
    count_launch();
    return MICO_OK;
}

extern "C" int mico_cast_bf16_to_f32(const void* src, float* dst, int64_t n, void* stream_) {
    using namespace mico;
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    ProfScope prof(kProfOther, 6.0 * (double)n, stream);
    MICO_CHECK_ARG(src && dst && n > 0);
    MICO_CHECK_ARG((reinterpret_cast<uintptr_t>(src) & 15) == 0 && (reinterpret_cast<uintptr_t>(dst) & 15) == 0);
    int64_t want = (n / 8 + 255) / 256;
    const int grid = (int)(want < 1 ? 1 : (want > num_sms() * 16 ? num_sms() * 16 : want));
    cast_bf16_f32_kernel<<<grid, 256, 0, stream>>>(reinterpret_cast<const __nv_bfloat16*>(src), dst, n);
    MICO_CHECK_CUDA(cudaGetLastError());
    count_launch();
    return MICO_OK;
}

extern "C" int mico_cast_f32_to_bf16_2d(const float* src, int64_t lds, int rows, int cols, void* dst, int64_t ldd,
                                        void* stream_) {
    using namespace mico;
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    ProfScope prof(kProfOther, 6.0 * (double)rows * cols, stream);
    MICO_CHECK_ARG(src && dst && rows > 0 && cols > 0 && lds >= cols && ldd >= cols);
    const int64_t total = (int64_t)rows * ldd;
    int64_t want = (total + 255) / 256;
    const int grid = (int)(want > num_sms() * 16 ? num_sms() * 16 : want);
    cast_f32_bf16_2d_kernel<<<grid, 256, 0, stream>>>(src, lds, rows, cols, reinterpret_cast<__nv_bfloat16*>(dst), ldd);
    MICO_CHECK_CUDA(cudaGetLastError());
    count_launch();
    return MICO_OK;
}

extern "C" size_t mico_colsum_workspace(int M, int N) { return (size_t)mico::colsum_splits(M, N) * N * sizeof(float); }

namespace mico {
namespace {
int colsum_launch(const void* x, int64_t ldx, int M, int N, int gap_start, int gap_len, float* out0, float* out1,
                  int accumulate, void* workspace, size_t ws_bytes, cudaStream_t stream) {
    static std::atomic<unsigned> slot_counter{0};
    const int splits = colsum_splits(M, N);
    MICO_CHECK_ARG(ws_bytes >= (size_t)splits * N * sizeof(float));
    const int colblocks = ceil_div(N, kColsumCols);
    dim3 grid(colblocks, splits);
    unsigned int* tickets = nullptr;
    if (colblocks <= kColsumMaxBlocks) {
        unsigned int* base = nullptr;
        MICO_CHECK_CUDA(cudaGetSymbolAddress(reinterpret_cast<void**>(&base), g_colsum_tickets));
        tickets = base + (slot_counter.fetch_add(1) % kColsumSlots) * kColsumMaxBlocks;   // concurrent launches: own slice
    }
    colsum_bf16_kernel<<<grid, 256, 0, stream>>>(reinterpret_cast<const __nv_bfloat16*>(x), ldx, M, N, gap_start, gap_len,
                                                 reinterpret_cast<float*>(workspace), out0, out1, accumulate, tickets);
    MICO_CHECK_CUDA(cudaGetLastError());
    count_launch();
    if (!tickets) {
        MICO_CHECK_ARG(gap_len == 0);
        colsum_finalize_kernel<<<ceil_div(N, 256), 256, 0, stream>>>(reinterpret_cast<const float*>(workspace), splits, N,
                                                                    out0, accumulate);
        MICO_CHECK_CUDA(cudaGetLastError());
        count_launch();
    }
    return MICO_OK;
}
}  // namespace
}  // namespace mico

extern "C" int mico_colsum_bf16(const void* x, int64_t ldx, int M, int N, float* out, int accumulate, void* workspace,
                                size_t ws_bytes, void* stream_) {
    using namespace mico;
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    ProfScope prof(kProfOther, 2.0 * (double)M * N, stream);
    MICO_CHECK_ARG(x && out && workspace && M > 0 && N > 0);
    MICO_CHECK_ARG(ldx % 8 == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0);
    return colsum_launch(x, ldx, M, N, N, 0, out, out, accumulate, workspace, ws_bytes, stream);
}

extern "C" int mico_colsum2_bf16(const void* x, int64_t ldx, int M, int n0, int gap, int n1, float* out0, float* out1,
                                 void* workspace, size_t ws_bytes, void* stream_) {
    using namespace mico;
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    ProfScope prof(kProfOther, 2.0 * (double)M * (n0 + n1), stream);
    MICO_CHECK_ARG(x && out0 && out1 && workspace && M > 0 && n0 > 0 && n1 > 0 && gap >= 0);
    MICO_CHECK_ARG(ldx % 8 == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0);
    MICO_CHECK_ARG(n0 % 8 == 0 && gap % 8 == 0 && n1 % 8 == 0 && ceil_div(n0 + n1, kColsumCols) <= kColsumMaxBlocks);
    return colsum_launch(x, ldx, M, n0 + n1, n0, gap, out0, out1, 0, workspace, ws_bytes, stream);
}

extern "C" int mico_batch_sum_f32(const float* x, int B, int64_t R, float* out, int accumulate, void* stream_) {
    using namespace mico;
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    ProfScope prof(kProfOther, 4.0 * (double)B * R, stream);
    MICO_CHECK_ARG(x && out && B > 0 && R > 0 && R % 4 == 0);
    batch_sum_f32_kernel<<<(int)((R / 4 + 255) / 256), 256, 0, stream>>>(x, B, R, out, accumulate);
    MICO_CHECK_CUDA(cudaGetLastError());
    count_launch();
    return MICO_OK;
}

extern "C" int mico_patchify(const float* img, int64_t img_stride, int64_t chan_stride, int B, int C, int H, int W,
                             int P, int Kpad, int tokens_per_img, int token_off, void* out, void* stream_) {
    using namespace mico;
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    ProfScope prof(kProfOther, (double)B * C * H * W * 4.0 + (double)B * (H / P) * (W / P) * Kpad * 2.0, stream);
    MICO_CHECK_ARG(img && out && B > 0 && C > 0 && P > 0 && H % P == 0 && W % P == 0);
    MICO_CHECK_ARG(Kpad >= C * P * P && Kpad % 8 == 0);
    if (tokens_per_img <= 0) { tokens_per_img = (H / P) * (W / P); token_off = 0; }
    MICO_CHECK_ARG(token_off >= 0 && tokens_per_img >= token_off + (H / P) * (W / P));
    const size_t smem = (size_t)C * P * W * sizeof(float);
    MICO_CHECK_ARG(smem <= 200 * 1024);
    static bool attr = false;
    if (smem > 48 * 1024 && !attr) {
        MICO_CHECK_CUDA(cudaFuncSetAttribute(patchify_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        attr = true;
    }
    patchify_kernel<<<B * (H / P), 256, smem, stream>>>(img, img_stride, chan_stride, C, H, W, P, Kpad, tokens_per_img,
                                                        token_off, reinterpret_cast<__nv_bfloat16*>(out));
    MICO_CHECK_CUDA(cudaGetLastError());
    count_launch();
    return MICO_OK;
}

extern "C" int mico_cls_pos_row(const float* cls_token, const float* pos0, float* x, int64_t sample_stride, int B, int D,
                                void* stream_) {
    using namespace mico;
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    ProfScope prof(kProfOther, 4.0 * (double)B * D, stream);
    MICO_CHECK_ARG(cls_token && pos0 && x && B > 0 && D > 0);
    cls_pos_row_kernel<<<ceil_div(B * D, 256), 256, 0, stream>>>(cls_token, pos0, x, sample_stride, B, D);
    MICO_CHECK_CUDA(cudaGetLastError());
    count_launch();
    return MICO_OK;
}

extern "C" int mico_scale_cast_bf16(const float* x, int64_t ldx, const float* row_scale, int rows_per_group, void* y,
                                    int64_t ldy, int M, int D, void* stream_) {
    using namespace mico;
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    ProfScope prof(kProfOther, 6.0 * (double)M * D, stream);
    MICO_CHECK_ARG(x && y && M > 0 && D > 0 && D % 4 == 0 && ldx % 4 == 0 && ldy % 4 == 0);
    MICO_CHECK_ARG(!(row_scale && rows_per_group <= 0));
    const int64_t total = (int64_t)M * (D / 4);
    int64_t want = (total + 255) / 256;
    const int grid = (int)(want > num_sms() * 16 ? num_sms() * 16 : want);
    scale_cast_bf16_kernel<<<grid, 256, 0, stream>>>(x, ldx, row_scale, rows_per_group,
                                                     reinterpret_cast<__nv_bfloat16*>(y), ldy, M, D);
    MICO_CHECK_CUDA(cudaGetLastError());
    count_launch();
    return MICO_OK;
}

extern "C" int mico_drop_path_scales(const float* drop_prob, int L, int B, uint64_t seed, uint64_t offset, float* out,
                                     void* stream_) {
    using namespace mico;
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    ProfScope prof(kProfOther, 4.0 * (double)L * 2 * B, stream);
    MICO_CHECK_ARG(drop_prob && out && L > 0 && B > 0);
    drop_path_scales_kernel<<<ceil_div(L * 2 * B, 256), 256, 0, stream>>>(drop_prob, L, B, seed, offset, out);
    MICO_CHECK_CUDA(cudaGetLastError());
    count_launch();
    return MICO_OK;
}
