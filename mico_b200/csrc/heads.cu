// mico_b200 -- small kernels around the text head and the losses (all fp32 SIMT; HBM- or latency-bound).
//
//   mico_embedding_gather / _scatter_add   K6  BertEmbeddings word + position + token-type sum (bert.py:139-146) and
//                                              its backward into the word table (LayerNorm is the K2 kernel)
//   mico_cross_entropy_fwd / _bwd          K7/K8  F.cross_entropy with ignore_index and label smoothing
//                                              (bert.py:1088-1090 MLM loss; vast.py:412-415 ITC; vast.py:456 ITM)
//   mico_l2norm_fwd / _bwd                 F.normalize(dim=-1) (vast.py:225 and every feat_* there)
//   mico_sgemm_strided                     K8 contrastive logits  sim = f . f_all^T / contra_temp (vast.py:405-408) and
//                                              the small fp32 head GEMMs (Contra_head mico.py:36-41, Match_head :44-52):
//                                              kept in fp32 -- logits are divided by 0.07, bf16 operands would cost 1e-2
//   mico_dot_f32                           scalar reductions (d contra_temp)
#include "common.cuh"
#include "host_utils.h"

namespace mico {
namespace {

// ---------------------------------------------------------------------------------------------- embeddings
// one warp per token row; D % 4 == 0
__global__ void __launch_bounds__(256)
embedding_gather_kernel(const int64_t* __restrict__ ids, const int64_t* __restrict__ type_ids,
                        const int64_t* __restrict__ pos_ids, int pos_offset, const float* __restrict__ word,
                        const float* __restrict__ pos, const float* __restrict__ type, float* __restrict__ out, int M, int S,
                        int D, int V, int P, int T) {
    const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (row >= M) return;
    int64_t id = ids[row];
    id = id < 0 ? 0 : (id >= V ? V - 1 : id);
    int64_t pi = pos_ids ? pos_ids[row] : (row % S) + pos_offset;
    pi = pi < 0 ? 0 : (pi >= P ? P - 1 : pi);
    int64_t ti = type_ids ? type_ids[row] : 0;
    ti = ti < 0 ? 0 : (ti >= T ? T - 1 : ti);
    const float4* w = reinterpret_cast<const float4*>(word + id * D);
    const float4* p = reinterpret_cast<const float4*>(pos + pi * D);
    const float4* t = reinterpret_cast<const float4*>(type + ti * D);
    float4* o = reinterpret_cast<float4*>(out + (int64_t)row * D);
    for (int i = threadIdx.x & 31; i < D / 4; i += 32) {
        const float4 a = __ldg(w + i), b = __ldg(p + i), c = __ldg(t + i);
        o[i] = make_float4(a.x + c.x + b.x, a.y + c.y + b.y, a.z + c.z + b.z, a.w + c.w + b.w);   // (word + type) + pos
    }
}

__global__ void __launch_bounds__(256)
embedding_scatter_add_kernel(const float* __restrict__ dx, const int64_t* __restrict__ ids, float* __restrict__ dtable,
                             int M, int D, int V) {
    const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (row >= M) return;
    const int64_t id = ids[row];
    if (id < 0 || id >= V) return;
    const float* src = dx + (int64_t)row * D;
    float* dst = dtable + id * D;
    for (int i = (threadIdx.x & 31); i < D; i += 32) atomicAdd(dst + i, src[i]);
}

// ---------------------------------------------------------------------------------------------- cross entropy
template <typename T> __device__ __forceinline__ float ldf(const T* p);
template <> __device__ __forceinline__ float ldf<float>(const float* p) { return *p; }
template <> __device__ __forceinline__ float ldf<__nv_bfloat16>(const __nv_bfloat16* p) { return __bfloat162float(*p); }

__device__ __forceinline__ float block_reduce_256(float v, float* red, bool is_max) {
    v = is_max ? warp_max(v) : warp_sum(v);
    __syncthreads();
    if (lane_id() == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    float r = red[0];
#pragma unroll
    for (int w = 1; w < 8; ++w) r = is_max ? fmaxf(r, red[w]) : r + red[w];
    return r;
}

// one 256-thread block per row.  loss_row = (1-ls) * (lse - x_y) + ls * (lse - mean_v x_v); 0 for ignored rows.
template <typename T>
__global__ void __launch_bounds__(256)
ce_fwd_kernel(const T* __restrict__ logits, int64_t ld, const int64_t* __restrict__ labels, int64_t ignore_index,
              float ls, float* __restrict__ row_loss, float* __restrict__ lse_out, int M, int V) {
    __shared__ float red[8];
    const int row = blockIdx.x;
    const T* x = logits + (int64_t)row * ld;
    float mx = -INFINITY;
    for (int i = threadIdx.x; i < V; i += 256) mx = fmaxf(mx, ldf(x + i));
    mx = block_reduce_256(mx, red, true);
    float se = 0.f, sx = 0.f;
    for (int i = threadIdx.x; i < V; i += 256) {
        const float v = ldf(x + i);
        se += __expf(v - mx);
        sx += v;
    }
    se = block_reduce_256(se, red, false);
    sx = block_reduce_256(sx, red, false);
    if (threadIdx.x == 0) {
        const float lse = mx + __logf(se);
        lse_out[row] = lse;
        const int64_t y = labels[row];
        float l = 0.f;
        if (y != ignore_index && y >= 0 && y < V) l = (1.0f - ls) * (lse - ldf(x + y)) + ls * (lse - sx / (float)V);
        row_loss[row] = l;
    }
}

// loss = sum(row_loss) / n_valid ; out[0] = loss, out[1] = n_valid   (single block, deterministic)
__global__ void __launch_bounds__(256)
ce_reduce_kernel(const float* __restrict__ row_loss, const int64_t* __restrict__ labels, int64_t ignore_index, int M,
                 int V, float* __restrict__ out) {
    __shared__ float red[8];
    float s = 0.f, n = 0.f;
    for (int i = threadIdx.x; i < M; i += 256) {
        s += row_loss[i];
        const int64_t y = labels[i];
        n += (y != ignore_index && y >= 0 && y < V) ? 1.f : 0.f;
    }
    s = block_reduce_256(s, red, false);
    n = block_reduce_256(n, red, false);
    if (threadIdx.x == 0) {
        out[0] = s / n;      // n == 0 -> nan, as torch
        out[1] = n;
    }
}

// dlogits = g / n_valid * (softmax - (1-ls) onehot - ls / V); zero rows for ignored labels
template <typename T, typename TO>
__global__ void __launch_bounds__(256)
ce_bwd_kernel(const T* __restrict__ logits, int64_t ld, const int64_t* __restrict__ labels, int64_t ignore_index,
              float ls, const float* __restrict__ lse, const float* __restrict__ grad, const float* __restrict__ stats,
              TO* __restrict__ dlogits, int64_t ldd, int M, int V) {
    const int row = blockIdx.x;
    const T* x = logits + (int64_t)row * ld;
    TO* d = dlogits + (int64_t)row * ldd;
    const int64_t y = labels[row];
    const bool valid = (y != ignore_index && y >= 0 && y < V);
    const float g = valid ? grad[0] / stats[1] : 0.f;
    const float l = lse[row];
    const float base = ls / (float)V;
    for (int i = threadIdx.x; i < V; i += 256) {
        float v = 0.f;
        if (valid) v = g * (__expf(ldf(x + i) - l) - base - (i == y ? 1.0f - ls : 0.f));
        if constexpr (sizeof(TO) == 2) d[i] = __float2bfloat16(v);
        else d[i] = v;
    }
}

// ---------------------------------------------------------------------------------------------- K7: chunked LM-head loss
// The LM head's [rows, 30522] logits are never materialised in training mode (bert.py:606-608 + 1084-1090 fused): the vocabulary
// is walked in chunks, each chunk's logits come out of the decoder GEMM into one reusable buffer, and these kernels keep an
// online log-sum-exp per row (running max m, running sum s of exp(x - m)) plus the label's logit; the backward pass recomputes
// a chunk's logits with the same GEMM and turns them into dlogits in place of a second full-size tensor.
// one 256-thread block per row
__global__ void __launch_bounds__(256)
ce_chunk_update_kernel(const float* __restrict__ logits, int64_t ld, int col0, int Vc, const int64_t* __restrict__ labels,
                       float* __restrict__ run_max, float* __restrict__ run_sum, float* __restrict__ label_logit, int first) {
    __shared__ float red[8];
    const int row = blockIdx.x;
    const float* x = logits + (int64_t)row * ld;
    float mx = -INFINITY;
    for (int i = threadIdx.x; i < Vc; i += 256) mx = fmaxf(mx, x[i]);
    mx = block_reduce_256(mx, red, true);
    float se = 0.f;
    for (int i = threadIdx.x; i < Vc; i += 256) se += __expf(x[i] - mx);
    se = block_reduce_256(se, red, false);
    if (threadIdx.x == 0) {
        const float m0 = first ? -INFINITY : run_max[row], s0 = first ? 0.f : run_sum[row];
        const float m1 = fmaxf(m0, mx);
        run_max[row] = m1;
        run_sum[row] = s0 * __expf(m0 - m1) + se * __expf(mx - m1);      // exp(-inf) = 0 on the first chunk
        const int64_t y = labels[row] - col0;
        if (y >= 0 && y < Vc) label_logit[row] = x[y];
    }
}
// lse = m + log s; row_loss = lse - x_label for rows whose label is a vocabulary index other than ignore_index, else 0
__global__ void ce_chunk_finalize_kernel(const float* __restrict__ run_max, const float* __restrict__ run_sum,
                                         const float* __restrict__ label_logit, const int64_t* __restrict__ labels,
                                         int64_t ignore_index, int V, float* __restrict__ row_loss, float* __restrict__ lse, int M) {
    const int row = blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= M) return;
    const float l = run_max[row] + __logf(run_sum[row]);
    lse[row] = l;
    const int64_t y = labels[row];
    row_loss[row] = (y != ignore_index && y >= 0 && y < V) ? l - label_logit[row] : 0.f;
}
// dlogits chunk (bf16) = g / n_valid * (exp(x - lse) - [label == col0 + j]); zero rows for ignored labels
__global__ void __launch_bounds__(256)
ce_chunk_grad_kernel(const float* __restrict__ logits, int64_t ld, int col0, int Vc, const int64_t* __restrict__ labels,
                     int64_t ignore_index, int V, const float* __restrict__ lse, const float* __restrict__ grad,
                     const float* __restrict__ stats, __nv_bfloat16* __restrict__ dlogits, int64_t ldd) {
    const int row = blockIdx.x;
    const float* x = logits + (int64_t)row * ld;
    __nv_bfloat16* d = dlogits + (int64_t)row * ldd;
    const int64_t y = labels[row];
    const bool valid = (y != ignore_index && y >= 0 && y < V);
    const float g = valid ? grad[0] / stats[1] : 0.f;
    const float l = lse[row];
    const int64_t yc = y - col0;
    for (int i = threadIdx.x; i < Vc; i += 256) {
        float v = 0.f;
        if (valid) v = g * (__expf(x[i] - l) - (i == yc ? 1.0f : 0.f));
        d[i] = __float2bfloat16(v);
    }
}

// ---------------------------------------------------------------------------------------------- L2 normalise
__global__ void __launch_bounds__(256)
l2norm_fwd_kernel(const float* __restrict__ x, float* __restrict__ y, float* __restrict__ norm, int M, int D, float eps) {
    const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (row >= M) return;
    const float* xr = x + (int64_t)row * D;
    float s = 0.f;
    for (int i = threadIdx.x & 31; i < D; i += 32) s += xr[i] * xr[i];
    const float n = fmaxf(sqrtf(warp_sum(s)), eps);
    if ((threadIdx.x & 31) == 0 && norm) norm[row] = n;
    const float inv = 1.0f / n;
    for (int i = threadIdx.x & 31; i < D; i += 32) y[(int64_t)row * D + i] = xr[i] * inv;
}

// dx = (dy - y * <y, dy>) / n
__global__ void __launch_bounds__(256)
l2norm_bwd_kernel(const float* __restrict__ y, const float* __restrict__ dy, const float* __restrict__ norm,
                  float* __restrict__ dx, int M, int D) {
    const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (row >= M) return;
    const float* yr = y + (int64_t)row * D;
    const float* dr = dy + (int64_t)row * D;
    float s = 0.f;
    for (int i = threadIdx.x & 31; i < D; i += 32) s += yr[i] * dr[i];
    s = warp_sum(s);
    const float inv = 1.0f / norm[row];
    for (int i = threadIdx.x & 31; i < D; i += 32) dx[(int64_t)row * D + i] = (dr[i] - yr[i] * s) * inv;
}

// ---------------------------------------------------------------------------------------------- small fp32 GEMM
// C[m,n] = alpha * (alpha_dev ? (alpha_recip ? 1 / *alpha_dev : *alpha_dev) : 1) * sum_k A(m,k) B(n,k) [+ bias[n]] [+ C]
// A(m,k) = a[m*a_sm + k*a_sk], B(n,k) = b[n*b_sn + k*b_sk]: any layout.  32x32 output tile, 16x16 threads, 2x2 each.
constexpr int kSgTile = 32, kSgK = 32;
__global__ void __launch_bounds__(256)
sgemm_strided_kernel(const float* __restrict__ a, int64_t a_sm, int64_t a_sk, const float* __restrict__ b, int64_t b_sn,
                     int64_t b_sk, float* __restrict__ c, int64_t ldc, const float* __restrict__ bias, int M, int N, int K,
                     float alpha, const float* __restrict__ alpha_dev, int alpha_recip, int accumulate) {
    __shared__ float sa[kSgK][kSgTile + 1], sb[kSgK][kSgTile + 1];
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    const int m0 = blockIdx.y * kSgTile, n0 = blockIdx.x * kSgTile;
    float acc[2][2] = {{0, 0}, {0, 0}};
    for (int k0 = 0; k0 < K; k0 += kSgK) {
        for (int i = threadIdx.x; i < kSgTile * kSgK; i += 256) {
            // pick the faster-varying index so that the dominant stride-1 dimension is read coalesced
            int r, kk;
            if (a_sk == 1) { kk = i % kSgK; r = i / kSgK; } else { r = i % kSgTile; kk = i / kSgTile; }
            sa[kk][r] = (m0 + r < M && k0 + kk < K) ? a[(int64_t)(m0 + r) * a_sm + (int64_t)(k0 + kk) * a_sk] : 0.f;
            if (b_sk == 1) { kk = i % kSgK; r = i / kSgK; } else { r = i % kSgTile; kk = i / kSgTile; }
            sb[kk][r] = (n0 + r < N && k0 + kk < K) ? b[(int64_t)(n0 + r) * b_sn + (int64_t)(k0 + kk) * b_sk] : 0.f;
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < kSgK; ++kk) {
            const float a0 = sa[kk][ty], a1 = sa[kk][ty + 16], b0 = sb[kk][tx], b1 = sb[kk][tx + 16];
            acc[0][0] += a0 * b0; acc[0][1] += a0 * b1; acc[1][0] += a1 * b0; acc[1][1] += a1 * b1;
        }
        __syncthreads();
    }
    float s = alpha;
    if (alpha_dev) s *= alpha_recip ? 1.0f / alpha_dev[0] : alpha_dev[0];
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            const int m = m0 + ty + 16 * i, n = n0 + tx + 16 * j;
            if (m < M && n < N) {
                float v = acc[i][j] * s + (bias ? bias[n] : 0.f);
                float* dst = c + (int64_t)m * ldc + n;
                *dst = accumulate ? *dst + v : v;
            }
        }
}

// exact-erf GELU on a small fp32 tensor (Match_head, mico.py:44-52) and its derivative
__global__ void gelu_f32_kernel(const float* __restrict__ x, const float* __restrict__ dy, float* __restrict__ out, int64_t n) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    out[i] = dy ? dy[i] * gelu_erf_grad(x[i]) : gelu_erf(x[i]);
}

// Hidden-state dropout (bert.py:148, 294, 372): out = [res +] x * m,  m = counter-based Bernoulli(1-p)/(1-p) of element
// (site_offset + flat index); forward (fp32 x, optional fp32 residual, fp32 and/or bf16 outputs) and backward
// (bf16 or fp32 gradient times the same mask) share the kernel.
template <typename TIn>
__global__ void dropout_kernel(const TIn* __restrict__ x, const float* __restrict__ res, float* __restrict__ out_f32,
                               __nv_bfloat16* __restrict__ out_bf16, int64_t n, DropCfg d, uint64_t site_offset) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float v = ldf(x + i) * drop_mult(d, site_offset + (uint64_t)i);
    if (res) v += res[i];
    if (out_f32) out_f32[i] = v;
    if (out_bf16) out_bf16[i] = __float2bfloat16(v);
}

// four elements per thread, 128-bit (bf16: 64-bit) accesses: the scalar kernel above moved 2.4 TB/s in the omni step
// (54 GB of hidden-dropout traffic in 22 ms); same mask (one splitmix64 per element counter)
template <typename TIn>
__global__ void __launch_bounds__(256)
dropout_vec4_kernel(const TIn* __restrict__ x, const float* __restrict__ res, float* __restrict__ out_f32,
                    __nv_bfloat16* __restrict__ out_bf16, int64_t n4, DropCfg d, uint64_t site_offset) {
    const int64_t i4 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i4 >= n4) return;
    const int64_t i = i4 * 4;
    float v[4];
    if constexpr (sizeof(TIn) == 4) {
        const float4 t = __ldg(reinterpret_cast<const float4*>(x) + i4);
        v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
    } else {
        const uint2 t = __ldg(reinterpret_cast<const uint2*>(x) + i4);
        v[0] = __uint_as_float(t.x << 16); v[1] = __uint_as_float(t.x & 0xFFFF0000u);
        v[2] = __uint_as_float(t.y << 16); v[3] = __uint_as_float(t.y & 0xFFFF0000u);
    }
    float4 r = make_float4(0.f, 0.f, 0.f, 0.f);
    if (res) r = __ldg(reinterpret_cast<const float4*>(res) + i4);
    const float rr[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
    for (int e = 0; e < 4; ++e) v[e] = v[e] * drop_mult(d, site_offset + (uint64_t)(i + e)) + rr[e];
    if (out_f32) reinterpret_cast<float4*>(out_f32)[i4] = make_float4(v[0], v[1], v[2], v[3]);
    if (out_bf16) reinterpret_cast<uint2*>(out_bf16)[i4] = make_uint2(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]));
}

// out[0] (+)= alpha * sum_i a[i] * b[i]   (single block, deterministic)
__global__ void __launch_bounds__(256)
dot_kernel(const float* __restrict__ a, const float* __restrict__ b, int64_t n, float alpha, float* __restrict__ out,
           int accumulate) {
    __shared__ float red[8];
    float s = 0.f;
    for (int64_t i = threadIdx.x; i < n; i += 256) s += a[i] * b[i];
    s = block_reduce_256(s, red, false);
    if (threadIdx.x == 0) out[0] = accumulate ? out[0] + alpha * s : alpha * s;
}

}  // namespace
}  // namespace mico

using namespace mico;

extern "C" int mico_embedding_gather(const int64_t* ids, const int64_t* type_ids, const int64_t* pos_ids, int pos_offset,
                                     const float* word, const float* pos, const float* type, float* out, int M, int S, int D,
                                     int V, int P, int T, void* stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    MICO_CHECK_ARG(ids && word && pos && type && out && M > 0 && S > 0 && D > 0 && D % 4 == 0 && V > 0 && P > 0 && T > 0);
    ProfScope prof(kProfOther, 16.0 * (double)M * D, stream);
    embedding_gather_kernel<<<ceil_div(M, 8), 256, 0, stream>>>(ids, type_ids, pos_ids, pos_offset, word, pos, type, out, M,
                                                              S, D, V, P, T);
    MICO_CHECK_CUDA(cudaGetLastError());
    count_launch();
    return MICO_OK;
}

extern "C" int mico_embedding_scatter_add(const float* dx, const int64_t* ids, float* dtable, int M, int D, int V,
                                          void* stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    MICO_CHECK_ARG(dx && ids && dtable && M > 0 && D > 0 && V > 0);
    ProfScope prof(kProfOther, 12.0 * (double)M * D, stream);
    embedding_scatter_add_kernel<<<ceil_div(M, 8), 256, 0, stream>>>(dx, ids, dtable, M, D, V);
    MICO_CHECK_CUDA(cudaGetLastError());
    count_launch();
    return MICO_OK;
}

extern "C" int mico_cross_entropy_fwd(const void* logits, int logits_bf16, int64_t ld, const int64_t* labels,
                                      int64_t ignore_index, float label_smoothing, float* row_loss, float* lse,
                                      float* loss_and_count, int M, int V, void* stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    MICO_CHECK_ARG(logits && labels && row_loss && lse && loss_and_count && M > 0 && V > 0 && ld >= V);
    ProfScope prof(kProfOther, (double)M * V * (logits_bf16 ? 2 : 4) * 2, stream);
    if (logits_bf16)
        ce_fwd_kernel<__nv_bfloat16><<<M, 256, 0, stream>>>(reinterpret_cast<const __nv_bfloat16*>(logits), ld, labels,
                                                           ignore_index, label_smoothing, row_loss, lse, M, V);
    else
        ce_fwd_kernel<float><<<M, 256, 0, stream>>>(reinterpret_cast<const float*>(logits), ld, labels, ignore_index,
                                                   label_smoothing, row_loss, lse, M, V);
    MICO_CHECK_CUDA(cudaGetLastError());
    ce_reduce_kernel<<<1, 256, 0, stream>>>(row_loss, labels, ignore_index, M, V, loss_and_count);
    MICO_CHECK_CUDA(cudaGetLastError());
    count_launch(2);
    return MICO_OK;
}

extern "C" int mico_cross_entropy_bwd(const void* logits, int logits_bf16, int64_t ld, const int64_t* labels,
                                      int64_t ignore_index, float label_smoothing, const float* lse, const float* grad,
                                      const float* loss_and_count, void* dlogits, int dlogits_bf16, int64_t ldd, int M, int V,
                                      void* stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    MICO_CHECK_ARG(logits && labels && lse && grad && loss_and_count && dlogits && M > 0 && V > 0 && ld >= V && ldd >= V);
    ProfScope prof(kProfOther, (double)M * V * ((logits_bf16 ? 2 : 4) + (dlogits_bf16 ? 2 : 4)), stream);
#define MICO_CE_BWD(TI, TO)                                                                                            \
    ce_bwd_kernel<TI, TO><<<M, 256, 0, stream>>>(reinterpret_cast<const TI*>(logits), ld, labels, ignore_index,         \
                                                 label_smoothing, lse, grad, loss_and_count,                           \
                                                 reinterpret_cast<TO*>(dlogits), ldd, M, V)
    if (logits_bf16 && dlogits_bf16) MICO_CE_BWD(__nv_bfloat16, __nv_bfloat16);
    else if (logits_bf16) MICO_CE_BWD(__nv_bfloat16, float);
    else if (dlogits_bf16) MICO_CE_BWD(float, __nv_bfloat16);
    else MICO_CE_BWD(float, float);
#undef MICO_CE_BWD
    MICO_CHECK_CUDA(cudaGetLastError());
    count_launch();
    return MICO_OK;
}

extern "C" int mico_ce_chunk_update(const float* logits, int64_t ld, int col0, int Vc, const int64_t* labels, float* run_max,
                                    float* run_sum, float* label_logit, int M, int first, void* stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    MICO_CHECK_ARG(logits && labels && run_max && run_sum && label_logit && M > 0 && Vc > 0 && ld >= Vc && col0 >= 0);
    ProfScope prof(kProfOther, (double)M * Vc * 4 * 2, stream);
    ce_chunk_update_kernel<<<M, 256, 0, stream>>>(logits, ld, col0, Vc, labels, run_max, run_sum, label_logit, first);
    MICO_CHECK_CUDA(cudaGetLastError());
    count_launch();
    return MICO_OK;
}

extern "C" int mico_ce_chunk_finalize(const float* run_max, const float* run_sum, const float* label_logit, const int64_t* labels,
                                      int64_t ignore_index, int V, float* row_loss, float* lse, float* loss_and_count, int M,
                                      void* stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    MICO_CHECK_ARG(run_max && run_sum && label_logit && labels && row_loss && lse && loss_and_count && M > 0 && V > 0);
    ce_chunk_finalize_kernel<<<ceil_div(M, 256), 256, 0, stream>>>(run_max, run_sum, label_logit, labels, ignore_index, V, row_loss,
                                                                 lse, M);
    MICO_CHECK_CUDA(cudaGetLastError());
    ce_reduce_kernel<<<1, 256, 0, stream>>>(row_loss, labels, ignore_index, M, V, loss_and_count);
    MICO_CHECK_CUDA(cudaGetLastError());
    count_launch(2);
    return MICO_OK;
}

extern "C" int mico_ce_chunk_grad(const float* logits, int64_t ld, int col0, int Vc, const int64_t* labels, int64_t ignore_index,
                                  int V, const float* lse, const float* grad, const float* loss_and_count, void* dlogits_bf16,
                                  int64_t ldd, int M, void* stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    MICO_CHECK_ARG(logits && labels && lse && grad && loss_and_count && dlogits_bf16 && M > 0 && Vc > 0 && ld >= Vc && ldd >= Vc);
    ProfScope prof(kProfOther, (double)M * Vc * (4 + 2), stream);
    ce_chunk_grad_kernel<<<M, 256, 0, stream>>>(logits, ld, col0, Vc, labels, ignore_index, V, lse, grad, loss_and_count,
                                              reinterpret_cast<__nv_bfloat16*>(dlogits_bf16), ldd);
    MICO_CHECK_CUDA(cudaGetLastError());
    count_launch();
    return MICO_OK;
}

extern "C" int mico_l2norm_fwd(const float* x, float* y, float* norm, int M, int D, float eps, void* stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    MICO_CHECK_ARG(x && y && M > 0 && D > 0);
    ProfScope prof(kProfOther, 8.0 * (double)M * D, stream);
    l2norm_fwd_kernel<<<ceil_div(M, 8), 256, 0, stream>>>(x, y, norm, M, D, eps);
    MICO_CHECK_CUDA(cudaGetLastError());
    count_launch();
    return MICO_OK;
}

extern "C" int mico_l2norm_bwd(const float* y, const float* dy, const float* norm, float* dx, int M, int D, void* stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    MICO_CHECK_ARG(y && dy && norm && dx && M > 0 && D > 0);
    ProfScope prof(kProfOther, 12.0 * (double)M * D, stream);
    l2norm_bwd_kernel<<<ceil_div(M, 8), 256, 0, stream>>>(y, dy, norm, dx, M, D);
    MICO_CHECK_CUDA(cudaGetLastError());
    count_launch();
    return MICO_OK;
}

extern "C" int mico_sgemm_strided(const float* a, int64_t a_sm, int64_t a_sk, const float* b, int64_t b_sn, int64_t b_sk,
                                  float* c, int64_t ldc, const float* bias, int M, int N, int K, float alpha,
                                  const float* alpha_dev, int alpha_recip, int accumulate, void* stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    MICO_CHECK_ARG(a && b && c && M > 0 && N > 0 && K > 0 && ldc >= N);
    ProfScope prof(kProfOther, 4.0 * ((double)M * K + (double)N * K + (double)M * N), stream);
    dim3 grid(ceil_div(N, kSgTile), ceil_div(M, kSgTile));
    sgemm_strided_kernel<<<grid, 256, 0, stream>>>(a, a_sm, a_sk, b, b_sn, b_sk, c, ldc, bias, M, N, K, alpha, alpha_dev,
                                                   alpha_recip, accumulate);
    MICO_CHECK_CUDA(cudaGetLastError());
    count_launch();
    return MICO_OK;
}

extern "C" int mico_dot_f32(const float* a, const float* b, int64_t n, float alpha, float* out, int accumulate,
                            void* stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    MICO_CHECK_ARG(a && b && out && n > 0);
    ProfScope prof(kProfOther, 8.0 * (double)n, stream);
    dot_kernel<<<1, 256, 0, stream>>>(a, b, n, alpha, out, accumulate);
    MICO_CHECK_CUDA(cudaGetLastError());
    count_launch();
    return MICO_OK;
}

extern "C" int mico_gelu_f32(const float* x, const float* dy, float* out, int64_t n, void* stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    MICO_CHECK_ARG(x && out && n > 0);
    ProfScope prof(kProfOther, (dy ? 12.0 : 8.0) * (double)n, stream);
    gelu_f32_kernel<<<(int)((n + 255) / 256), 256, 0, stream>>>(x, dy, out, n);
    MICO_CHECK_CUDA(cudaGetLastError());
    count_launch();
    return MICO_OK;
}

extern "C" int mico_dropout(const void* x, int x_is_bf16, const float* res, float* out_f32, void* out_bf16, int64_t n, float p,
                            uint64_t seed, uint64_t site_offset, void* stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    MICO_CHECK_ARG(x && (out_f32 || out_bf16) && n > 0 && p >= 0.0f && p < 1.0f);
    ProfScope prof(kProfOther, (double)n * ((x_is_bf16 ? 2 : 4) + (res ? 4 : 0) + (out_f32 ? 4 : 0) + (out_bf16 ? 2 : 0)), stream);
    DropCfg d;
    d.p = p; d.inv_keep = 1.0f / (1.0f - p); d.seed = seed;
    auto al16 = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; };
    if (n % 4 == 0 && al16(x) && al16(res) && al16(out_f32) && al16(out_bf16)) {
        const int64_t n4 = n / 4;
        const int grid4 = (int)((n4 + 255) / 256);
        if (x_is_bf16)
            dropout_vec4_kernel<__nv_bfloat16><<<grid4, 256, 0, stream>>>(reinterpret_cast<const __nv_bfloat16*>(x), res, out_f32,
                                                                         reinterpret_cast<__nv_bfloat16*>(out_bf16), n4, d, site_offset);
        else
            dropout_vec4_kernel<float><<<grid4, 256, 0, stream>>>(reinterpret_cast<const float*>(x), res, out_f32,
                                                                 reinterpret_cast<__nv_bfloat16*>(out_bf16), n4, d, site_offset);
        MICO_CHECK_CUDA(cudaGetLastError());
        count_launch();
        return MICO_OK;
    }
    const int grid = (int)((n + 255) / 256);
    if (x_is_bf16)
        dropout_kernel<__nv_bfloat16><<<grid, 256, 0, stream>>>(reinterpret_cast<const __nv_bfloat16*>(x), res, out_f32,
                                                               reinterpret_cast<__nv_bfloat16*>(out_bf16), n, d, site_offset);
    else
        dropout_kernel<float><<<grid, 256, 0, stream>>>(reinterpret_cast<const float*>(x), res, out_f32,
                                                       reinterpret_cast<__nv_bfloat16*>(out_bf16), n, d, site_offset);
    MICO_CHECK_CUDA(cudaGetLastError());
    count_launch();
    return MICO_OK;
}
