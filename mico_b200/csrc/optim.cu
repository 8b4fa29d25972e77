// mico_b200 -- SURVEY 8(f).1: fused multi-tensor AdamW (decoupled weight decay "fix") for the whole parameter set.
//
// Replaces the per-parameter Python loop of data/utils/build_optimizer.py:136-196 (AdamW.step: ~8 elementwise torch
// kernels per tensor x 900 tensors per step):
//     m = beta1*m + (1-beta1)*g            v = beta2*v + (1-beta2)*g*g           denom = sqrt(v) + eps
//     p -= step_size * m / denom           step_size = lr * sqrt(1-beta2^t) / (1-beta1^t)   (correct_bias, host side)
//     p -= lr * weight_decay * p           (after the Adam update, on the updated value; build_optimizer.py:193-194)
// ONE launch walks a device-resident table of (tensor, chunk) work items; every fp32 element is read once and written
// once with 128-bit accesses (HBM-bound: 28 B/element, +2 B when the bf16 GEMM-operand copy of the updated parameter
// is emitted -- which replaces the tower's per-step weight cast, mico_cast_f32_to_bf16).
#include "common.cuh"
#include "host_utils.h"

namespace mico {
namespace {

constexpr int kAdamThreads = 256;
constexpr int kAdamMaxGroups = 16;

struct AdamHyperTable {
    MicoAdamHyper g[kAdamMaxGroups];
};

__device__ __forceinline__ float adam_one(float& p, float g, float& m, float& v, const MicoAdamHyper& h, float gscale) {
    g *= gscale;
    m = m * h.beta1 + (1.0f - h.beta1) * g;
    v = v * h.beta2 + (1.0f - h.beta2) * g * g;
    const float denom = sqrtf(v) + h.eps;
    float x = p - h.step_size * (m / denom);
    if (h.weight_decay > 0.0f) x = x - (h.lr * h.weight_decay) * x;
    p = x;
    return x;
}

__global__ void __launch_bounds__(kAdamThreads)
adamw_multi_kernel(const MicoAdamTensor* __restrict__ tensors, const int32_t* __restrict__ chunk_tensor,
                   const int32_t* __restrict__ chunk_index, int chunk_elems, AdamHyperTable hyper, float gscale) {
    const MicoAdamTensor t = tensors[chunk_tensor[blockIdx.x]];
    const MicoAdamHyper h = hyper.g[t.group];
    const int64_t begin = (int64_t)chunk_index[blockIdx.x] * chunk_elems;
    const int64_t end = begin + chunk_elems < t.n ? begin + chunk_elems : t.n;
    float* p = t.p;
    const float* g = t.g;
    float* m = t.m;
    float* v = t.v;
    __nv_bfloat16* pb = reinterpret_cast<__nv_bfloat16*>(t.p_bf16);
    const bool vec = ((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(m) |
                       reinterpret_cast<uintptr_t>(v)) & 15) == 0 && (reinterpret_cast<uintptr_t>(pb) & 7) == 0;
    auto scalar = [&](int64_t j) {
        float pj = p[j], mj = m[j], vj = v[j];
        adam_one(pj, g[j], mj, vj, h, gscale);
        p[j] = pj; m[j] = mj; v[j] = vj;
        if (pb) pb[j] = __float2bfloat16(pj);
    };
    if (!vec) {     // unaligned tensor (a view at an odd offset): scalar
        for (int64_t j = begin + threadIdx.x; j < end; j += kAdamThreads) scalar(j);
        return;
    }
    const int64_t n4 = (end - begin) >> 2;
    for (int64_t q = threadIdx.x; q < n4; q += kAdamThreads) {
        const int64_t i = begin + 4 * q;
        float4 p4 = *reinterpret_cast<const float4*>(p + i);
        const float4 g4 = *reinterpret_cast<const float4*>(g + i);
        float4 m4 = *reinterpret_cast<const float4*>(m + i);
        float4 v4 = *reinterpret_cast<const float4*>(v + i);
        adam_one(p4.x, g4.x, m4.x, v4.x, h, gscale);
        adam_one(p4.y, g4.y, m4.y, v4.y, h, gscale);
        adam_one(p4.z, g4.z, m4.z, v4.z, h, gscale);
        adam_one(p4.w, g4.w, m4.w, v4.w, h, gscale);
        *reinterpret_cast<float4*>(p + i) = p4;
        *reinterpret_cast<float4*>(m + i) = m4;
        *reinterpret_cast<float4*>(v + i) = v4;
        if (pb) *reinterpret_cast<uint2*>(pb + i) = make_uint2(pack_bf16x2(p4.x, p4.y), pack_bf16x2(p4.z, p4.w));
    }
    for (int64_t j = begin + 4 * n4 + threadIdx.x; j < end; j += kAdamThreads) scalar(j);   // < 4 elements
}

}  // namespace
}  // namespace mico

extern "C" int mico_adamw_multi(const MicoAdamTensor* tensors_dev, const int32_t* chunk_tensor_dev,
                                const int32_t* chunk_index_dev, int n_chunks, int chunk_elems, const MicoAdamHyper* hyper,
                                int n_groups, float grad_scale, double total_elems, void* stream_) {
    using namespace mico;
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    MICO_CHECK_ARG(tensors_dev && chunk_tensor_dev && chunk_index_dev && hyper);
    MICO_CHECK_ARG(n_chunks > 0 && chunk_elems > 0 && chunk_elems % 4 == 0);
    MICO_CHECK_ARG(n_groups > 0 && n_groups <= kAdamMaxGroups);
    AdamHyperTable tab;
    for (int i = 0; i < kAdamMaxGroups; ++i) tab.g[i] = hyper[i < n_groups ? i : 0];
    ProfScope prof(kProfOther, 28.0 * total_elems, stream);
    adamw_multi_kernel<<<n_chunks, kAdamThreads, 0, stream>>>(tensors_dev, chunk_tensor_dev, chunk_index_dev, chunk_elems, tab,
                                                              grad_scale);
    MICO_CHECK_CUDA(cudaGetLastError());
    count_launch();
    return MICO_OK;
}
