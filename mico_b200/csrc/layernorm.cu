// mico_b200 -- K2: LayerNorm forward / backward (HBM-bound; one warp per row, 128-bit accesses).
//
// Replaces torch LayerNorm at eva_vit_model.py:375,382,542 (eps 1e-6), bert.py:92,290,368,583 and
// mico.py:49,400-403 (eps 1e-12), swin.py:212,218,329,565 (eps 1e-5).
//
// forward : y = (x - mean) * rstd * gamma + beta, statistics in fp32, two-pass (mean, then centred
//           variance) with the row held in registers; writes bf16 (next GEMM operand) and/or fp32.
// backward: dx = [dres +] rstd * (g - mean(g) - xhat * mean(g*xhat)),  g = dy * gamma
//           optionally also emits bf16(dx * row_scale[row / rows_per_group]) -- the operand of the
//           upstream linear's dgrad/wgrad (DropPath scale folded in);
//           dgamma/dbeta: per-warp shared-memory accumulators -> per-block partials in a caller
//           workspace -> deterministic finalize kernel (no atomics).
#include "common.cuh"
#include "host_utils.h"

namespace mico {
namespace {

constexpr int kLnWarps = 4;
constexpr int kLnThreads = kLnWarps * 32;
constexpr int kMaxVec = 12;   // float4 per lane kept in registers: D <= 12*32*4 = 1536

__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ float4 ld4(const __nv_bfloat16* p) {
    const uint2 u = *reinterpret_cast<const uint2*>(p);
    return make_float4(bf16_lo(u.x), bf16_hi(u.x), bf16_lo(u.y), bf16_hi(u.y));
}
__device__ __forceinline__ void st4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
__device__ __forceinline__ void st4(__nv_bfloat16* p, float4 v) {
    *reinterpret_cast<uint2*>(p) = make_uint2(pack_bf16x2(v.x, v.y), pack_bf16x2(v.z, v.w));
}

template <typename TIn>
__global__ void __launch_bounds__(kLnThreads)
ln_fwd_kernel(const TIn* __restrict__ x, int64_t ldx, const float* __restrict__ gamma,
              const float* __restrict__ beta, __nv_bfloat16* __restrict__ y_bf16, float* __restrict__ y_f32,
              int64_t ldy, float* __restrict__ mean_out, float* __restrict__ rstd_out, int M, int D, float eps) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nvec = D >> 2;
    for (int row = blockIdx.x * kLnWarps + warp; row < M; row += gridDim.x * kLnWarps) {
        const TIn* xr = x + (int64_t)row * ldx;
        float4 v[kMaxVec];
        float s = 0.f;
#pragma unroll
        for (int j = 0; j < kMaxVec; ++j) {
            const int i = lane + 32 * j;
            if (i < nvec) {
                v[j] = ld4(xr + 4 * i);
                s += (v[j].x + v[j].y) + (v[j].z + v[j].w);
            }
        }
        const float mean = warp_sum(s) / (float)D;
        float q = 0.f;
#pragma unroll
        for (int j = 0; j < kMaxVec; ++j) {
            const int i = lane + 32 * j;
            if (i < nvec) {
                const float a = v[j].x - mean, b = v[j].y - mean, c = v[j].z - mean, d = v[j].w - mean;
                q += (a * a + b * b) + (c * c + d * d);
            }
        }
        const float rstd = rsqrtf(warp_sum(q) / (float)D + eps);
        if (lane == 0) {
            if (mean_out) mean_out[row] = mean;
            if (rstd_out) rstd_out[row] = rstd;
        }
#pragma unroll
        for (int j = 0; j < kMaxVec; ++j) {
            const int i = lane + 32 * j;
            if (i < nvec) {
                const float4 g = ld4(gamma + 4 * i), b = ld4(beta + 4 * i);
                float4 o;
                o.x = (v[j].x - mean) * rstd * g.x + b.x;
                o.y = (v[j].y - mean) * rstd * g.y + b.y;
                o.z = (v[j].z - mean) * rstd * g.z + b.z;
                o.w = (v[j].w - mean) * rstd * g.w + b.w;
                if (y_bf16) st4(y_bf16 + (int64_t)row * ldy + 4 * i, o);
                if (y_f32) st4(y_f32 + (int64_t)row * ldy + 4 * i, o);
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// Forward with TMA-staged rows.  The register-staged kernel above keeps ~100 KB of loads in flight per SM only while its
// warps sit in the load phase (ncu round 1: 40 % of DRAM peak, 62 % long-scoreboard stalls at 28 % occupancy).  Here each
// warp owns a ring of kLnStages row buffers in shared memory that the TMA engine fills with 1-D bulk copies; the row after
// next is already travelling while the current one is normalised, independent of what the warp is doing:
// 8 warps x 3 rows x 5.6 KB = 135 KB in flight per SM all the time.  One persistent block per SM.
constexpr int kLnBulkWarps = 8;
constexpr int kLnStages = 4;

template <typename TIn>
__global__ void __launch_bounds__(kLnBulkWarps * 32, 1)
ln_fwd_bulk_kernel(const TIn* __restrict__ x, int64_t ldx, const float* __restrict__ gamma,
                   const float* __restrict__ beta, __nv_bfloat16* __restrict__ y_bf16, float* __restrict__ y_f32,
                   int64_t ldy, float* __restrict__ mean_out, float* __restrict__ rstd_out, int M, int D, float eps,
                   int row_bytes_padded) {
    extern __shared__ __align__(128) uint8_t ln_bulk_smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nvec = D >> 2;
    const uint32_t row_bytes = (uint32_t)D * sizeof(TIn);
    uint8_t* ring = ln_bulk_smem + (size_t)warp * kLnStages * row_bytes_padded;
    uint64_t* bars = reinterpret_cast<uint64_t*>(ln_bulk_smem + (size_t)kLnBulkWarps * kLnStages * row_bytes_padded) +
                     warp * kLnStages;
    const int row0 = blockIdx.x * kLnBulkWarps + warp;
    const int stride = gridDim.x * kLnBulkWarps;
    if (lane == 0) {
        for (int s = 0; s < kLnStages; ++s) mbar_init(&bars[s], 1);
        fence_mbar_init();
        for (int s = 0; s < kLnStages; ++s) {
            const int r = row0 + s * stride;
            if (r < M) {
                mbar_arrive_expect_tx(&bars[s], row_bytes);
                bulk_load_1d(smem_u32(ring + (size_t)s * row_bytes_padded), x + (int64_t)r * ldx, row_bytes, &bars[s]);
            }
        }
    }
    __syncwarp();
    // gamma / beta of this lane's columns stay in registers for all of the warp's rows (with eight warps per SM the
    // per-row parameter loads were the exposed latency: 45 % long-scoreboard stalls in the first version)
    float4 gm[kMaxVec], bt[kMaxVec];
#pragma unroll
    for (int j = 0; j < kMaxVec; ++j) {
        const int i = lane + 32 * j;
        gm[j] = i < nvec ? ld4(gamma + 4 * i) : make_float4(0, 0, 0, 0);
        bt[j] = i < nvec ? ld4(beta + 4 * i) : make_float4(0, 0, 0, 0);
    }
    const float inv_d = 1.0f / (float)D;
    int it = 0;
    for (int row = row0; row < M; row += stride, ++it) {
        const int s = it % kLnStages;
        mbar_wait(&bars[s], (it / kLnStages) & 1);
        const TIn* xr = reinterpret_cast<const TIn*>(ring + (size_t)s * row_bytes_padded);
        float4 v[kMaxVec];
        float sum = 0.f;
#pragma unroll
        for (int j = 0; j < kMaxVec; ++j) {
            const int i = lane + 32 * j;
            if (i < nvec) {
                v[j] = ld4(xr + 4 * i);
                sum += (v[j].x + v[j].y) + (v[j].z + v[j].w);
            }
        }
        __syncwarp();                       // every lane has read the buffer: refill it with the row kLnStages ahead
        if (lane == 0) {
            const int r = row + kLnStages * stride;
            if (r < M) {
                mbar_arrive_expect_tx(&bars[s], row_bytes);
                bulk_load_1d(smem_u32(ring + (size_t)s * row_bytes_padded), x + (int64_t)r * ldx, row_bytes, &bars[s]);
            }
        }
        const float mean = warp_sum(sum) * inv_d;
        float q = 0.f;
#pragma unroll
        for (int j = 0; j < kMaxVec; ++j) {
            const int i = lane + 32 * j;
            if (i < nvec) {
                const float a = v[j].x - mean, b = v[j].y - mean, c = v[j].z - mean, d = v[j].w - mean;
                q += (a * a + b * b) + (c * c + d * d);
            }
        }
        const float rstd = rsqrtf(warp_sum(q) * inv_d + eps);
        if (lane == 0) {
            if (mean_out) mean_out[row] = mean;
            if (rstd_out) rstd_out[row] = rstd;
        }
#pragma unroll
        for (int j = 0; j < kMaxVec; ++j) {
            const int i = lane + 32 * j;
            if (i < nvec) {
                float4 o;
                o.x = (v[j].x - mean) * rstd * gm[j].x + bt[j].x;
                o.y = (v[j].y - mean) * rstd * gm[j].y + bt[j].y;
                o.z = (v[j].z - mean) * rstd * gm[j].z + bt[j].z;
                o.w = (v[j].w - mean) * rstd * gm[j].w + bt[j].w;
                if (y_bf16) st4(y_bf16 + (int64_t)row * ldy + 4 * i, o);
                if (y_f32) st4(y_f32 + (int64_t)row * ldy + 4 * i, o);
            }
        }
    }
}

// Each lane owns the same columns of every row it visits, so the dgamma / dbeta partial sums live in registers
// for the whole kernel; all global loads of a row are issued before the first use (the kernel is latency-bound
// otherwise: ncu round 1 showed 90 % long-scoreboard stalls at 15 % occupancy with shared-memory accumulators).
// dynamic smem: [kLnWarps][2][D] floats, used once at the end for the cross-warp reduction.
template <typename TDy>
__global__ void __launch_bounds__(kLnThreads, 2)
ln_bwd_kernel(const TDy* __restrict__ dy, int64_t lddy, const __nv_bfloat16* __restrict__ dy2, int64_t lddy2,
              const float* __restrict__ x, int64_t ldx,
              const float* __restrict__ mean, const float* __restrict__ rstd, const float* __restrict__ gamma,
              const float* __restrict__ dres, int64_t lddres, float* __restrict__ dx, int64_t lddx,
              __nv_bfloat16* __restrict__ dx_bf16, int64_t lddxb, const float* __restrict__ row_scale,
              int rows_per_group, float* __restrict__ partials, int M, int D) {
    extern __shared__ float ln_smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nvec = D >> 2;
    float4 ag[kMaxVec], ab[kMaxVec];
#pragma unroll
    for (int j = 0; j < kMaxVec; ++j) { ag[j] = make_float4(0, 0, 0, 0); ab[j] = make_float4(0, 0, 0, 0); }
    for (int row = blockIdx.x * kLnWarps + warp; row < M; row += gridDim.x * kLnWarps) {
        const TDy* dyr = dy + (int64_t)row * lddy;
        const float* xr = x + (int64_t)row * ldx;
        float4 d[kMaxVec], xh[kMaxVec];
#pragma unroll
        for (int j = 0; j < kMaxVec; ++j) {     // issue every load of the row first
            const int i = lane + 32 * j;
            if (i < nvec) {
                d[j] = ld4(dyr + 4 * i);
                xh[j] = ld4(xr + 4 * i);
            }
        }
        if (dy2) {   // second consumer of the LayerNorm output (post-LN BERT: next GEMM's dgrad + next residual)
            const __nv_bfloat16* d2r = dy2 + (int64_t)row * lddy2;
#pragma unroll
            for (int j = 0; j < kMaxVec; ++j) {
                const int i = lane + 32 * j;
                if (i < nvec) {
                    const float4 e = ld4(d2r + 4 * i);
                    d[j].x += e.x; d[j].y += e.y; d[j].z += e.z; d[j].w += e.w;
                }
            }
        }
        const float mu = mean[row], rs = rstd[row];
        float s1 = 0.f, s2 = 0.f;
#pragma unroll
        for (int j = 0; j < kMaxVec; ++j) {
            const int i = lane + 32 * j;
            if (i < nvec) {
                const float4 gm = ld4(gamma + 4 * i);
                xh[j] = make_float4((xh[j].x - mu) * rs, (xh[j].y - mu) * rs, (xh[j].z - mu) * rs, (xh[j].w - mu) * rs);
                ag[j].x += d[j].x * xh[j].x; ag[j].y += d[j].y * xh[j].y; ag[j].z += d[j].z * xh[j].z; ag[j].w += d[j].w * xh[j].w;
                ab[j].x += d[j].x; ab[j].y += d[j].y; ab[j].z += d[j].z; ab[j].w += d[j].w;
                d[j] = make_float4(d[j].x * gm.x, d[j].y * gm.y, d[j].z * gm.z, d[j].w * gm.w);   // g = dy * gamma
                s1 += (d[j].x + d[j].y) + (d[j].z + d[j].w);
                s2 += (d[j].x * xh[j].x + d[j].y * xh[j].y) + (d[j].z * xh[j].z + d[j].w * xh[j].w);
            }
        }
        const float m1 = warp_sum(s1) / (float)D;
        const float m2 = warp_sum(s2) / (float)D;
        const float sc = row_scale ? row_scale[row / rows_per_group] : 1.0f;
#pragma unroll
        for (int jb = 0; jb < kMaxVec; jb += 4) {      // residual loads in batches of four, then finish four vectors
            float4 r[4];
#pragma unroll
            for (int t = 0; t < 4; ++t) {
                const int i = lane + 32 * (jb + t);
                r[t] = (dres && i < nvec) ? ld4(dres + (int64_t)row * lddres + 4 * i) : make_float4(0, 0, 0, 0);
            }
#pragma unroll
            for (int t = 0; t < 4; ++t) {
                const int j = jb + t;
                const int i = lane + 32 * j;
                if (i < nvec) {
                    float4 o;
                    o.x = rs * (d[j].x - m1 - xh[j].x * m2) + r[t].x;
                    o.y = rs * (d[j].y - m1 - xh[j].y * m2) + r[t].y;
                    o.z = rs * (d[j].z - m1 - xh[j].z * m2) + r[t].z;
                    o.w = rs * (d[j].w - m1 - xh[j].w * m2) + r[t].w;
                    if (dx) st4(dx + (int64_t)row * lddx + 4 * i, o);
                    if (dx_bf16)
                        st4(dx_bf16 + (int64_t)row * lddxb + 4 * i, make_float4(o.x * sc, o.y * sc, o.z * sc, o.w * sc));
                }
            }
        }
    }
    // cross-warp reduction of the register partials -> partials[block][2][D]
    float* sg = ln_smem + (size_t)warp * 2 * D;
    float* sb = sg + D;
#pragma unroll
    for (int j = 0; j < kMaxVec; ++j) {
        const int i = lane + 32 * j;
        if (i < nvec) { st4(sg + 4 * i, ag[j]); st4(sb + 4 * i, ab[j]); }
    }
    __syncthreads();
    float* out = partials + (size_t)blockIdx.x * 2 * D;
    for (int i = threadIdx.x; i < 2 * D; i += kLnThreads) {
        float a = 0.f;
#pragma unroll
        for (int w = 0; w < kLnWarps; ++w) a += ln_smem[(size_t)w * 2 * D + i];
        out[i] = a;
    }
}


// ---------------------------------------------------------------------------------------------------------------
// Block-per-row backward (128 <= D/4 <= 384 vectors, i.e. every tower on the path): the 128 threads of a block share a
// row, so a thread holds 3 float4 of each operand instead of 12 -- ~100 registers, five resident blocks per SM, and
// ALL loads of a row (dy, x, residual gradient) are issued together: one exposed memory latency per row instead of
// two, 70 KB in flight per SM (ncu round 1: the warp-per-row kernel reached 3.7-5.1 of 6.5 TB/s).  Each thread owns the
// same columns of every row, so dgamma, dbeta and the column sum of the scaled bf16 output (the bias gradient of the
// upstream linear layer: nn.Linear bias backward fused here instead of re-reading the tensor) stay in registers and go
// straight to the per-block partials -- no cross-warp reduction.
constexpr int kV2 = 3;
// kDrop: the bf16 copy (and its column sums) is the gradient w.r.t. the INPUT of a hidden-state dropout that sat between the
// upstream dense layer and this LayerNorm's residual add (post-LN BERT, bert.py:293-296): element (row, col) is multiplied by
// the same counter-based mask the forward pass used (drop_mult(seed, site_offset + row * D + col)); the fp32 dx (the residual
// gradient) is not masked.  Replaces a separate pass over the bf16 gradient plus a column-sum pass.
template <typename TDy, bool kColsum, bool kDrop = false>
__global__ void __launch_bounds__(kLnThreads, 5)
ln_bwd_row_kernel(const TDy* __restrict__ dy, int64_t lddy, const __nv_bfloat16* __restrict__ dy2, int64_t lddy2,
                  const float* __restrict__ x, int64_t ldx, const float* __restrict__ mean,
                  const float* __restrict__ rstd, const float* __restrict__ gamma, const float* __restrict__ dres,
                  int64_t lddres, float* __restrict__ dx, int64_t lddx, __nv_bfloat16* __restrict__ dx_bf16,
                  int64_t lddxb, const float* __restrict__ row_scale, int rows_per_group, float* __restrict__ partials,
                  int M, int D, DropCfg drop = DropCfg{}, uint64_t drop_site = 0) {
    __shared__ float red[2][kLnWarps][2];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nvec = D >> 2;
    float4 ag[kV2], ab[kV2], ac[kV2], gm[kV2];
#pragma unroll
    for (int j = 0; j < kV2; ++j) {
        ag[j] = make_float4(0, 0, 0, 0); ab[j] = make_float4(0, 0, 0, 0); ac[j] = make_float4(0, 0, 0, 0);
        const int i = threadIdx.x + kLnThreads * j;
        gm[j] = i < nvec ? ld4(gamma + 4 * i) : make_float4(0, 0, 0, 0);
    }
    const float inv_d = 1.0f / (float)D;
    int it = 0;
    for (int row = blockIdx.x; row < M; row += gridDim.x, ++it) {
        const TDy* dyr = dy + (int64_t)row * lddy;
        const float* xr = x + (int64_t)row * ldx;
        float4 d[kV2], xh[kV2], r[kV2];
#pragma unroll
        for (int j = 0; j < kV2; ++j) {     // every load of the row first
            const int i = threadIdx.x + kLnThreads * j;
            const bool ok = i < nvec;
            d[j] = ok ? ld4(dyr + 4 * i) : make_float4(0, 0, 0, 0);
            xh[j] = ok ? ld4(xr + 4 * i) : make_float4(0, 0, 0, 0);
            r[j] = (ok && dres) ? ld4(dres + (int64_t)row * lddres + 4 * i) : make_float4(0, 0, 0, 0);
        }
        if (dy2) {
            const __nv_bfloat16* d2r = dy2 + (int64_t)row * lddy2;
#pragma unroll
            for (int j = 0; j < kV2; ++j) {
                const int i = threadIdx.x + kLnThreads * j;
                if (i < nvec) {
                    const float4 e = ld4(d2r + 4 * i);
                    d[j].x += e.x; d[j].y += e.y; d[j].z += e.z; d[j].w += e.w;
                }
            }
        }
        const float mu = __ldg(mean + row), rs = __ldg(rstd + row);
        const float sc = row_scale ? __ldg(row_scale + row / rows_per_group) : 1.0f;
        float s1 = 0.f, s2 = 0.f;
#pragma unroll
        for (int j = 0; j < kV2; ++j) {
            const int i = threadIdx.x + kLnThreads * j;
            if (i < nvec) {
                xh[j] = make_float4((xh[j].x - mu) * rs, (xh[j].y - mu) * rs, (xh[j].z - mu) * rs, (xh[j].w - mu) * rs);
                ag[j].x += d[j].x * xh[j].x; ag[j].y += d[j].y * xh[j].y; ag[j].z += d[j].z * xh[j].z; ag[j].w += d[j].w * xh[j].w;
                ab[j].x += d[j].x; ab[j].y += d[j].y; ab[j].z += d[j].z; ab[j].w += d[j].w;
                d[j] = make_float4(d[j].x * gm[j].x, d[j].y * gm[j].y, d[j].z * gm[j].z, d[j].w * gm[j].w);
                s1 += (d[j].x + d[j].y) + (d[j].z + d[j].w);
                s2 += (d[j].x * xh[j].x + d[j].y * xh[j].y) + (d[j].z * xh[j].z + d[j].w * xh[j].w);
            }
        }
        s1 = warp_sum(s1);
        s2 = warp_sum(s2);
        float (*rb)[2] = red[it & 1];      // double-buffered: one __syncthreads per row
        if (lane == 0) { rb[warp][0] = s1; rb[warp][1] = s2; }
        __syncthreads();
        const float m1 = ((rb[0][0] + rb[1][0]) + (rb[2][0] + rb[3][0])) * inv_d;
        const float m2 = ((rb[0][1] + rb[1][1]) + (rb[2][1] + rb[3][1])) * inv_d;
#pragma unroll
        for (int j = 0; j < kV2; ++j) {
            const int i = threadIdx.x + kLnThreads * j;
            if (i < nvec) {
                float4 o;
                o.x = rs * (d[j].x - m1 - xh[j].x * m2) + r[j].x;
                o.y = rs * (d[j].y - m1 - xh[j].y * m2) + r[j].y;
                o.z = rs * (d[j].z - m1 - xh[j].z * m2) + r[j].z;
                o.w = rs * (d[j].w - m1 - xh[j].w * m2) + r[j].w;
                if (dx) st4(dx + (int64_t)row * lddx + 4 * i, o);
                float4 os = make_float4(o.x * sc, o.y * sc, o.z * sc, o.w * sc);
                if constexpr (kDrop) {
                    const uint64_t ctr = drop_site + (uint64_t)row * (uint64_t)D + (uint64_t)(4 * i);
                    os.x *= drop_mult(drop, ctr); os.y *= drop_mult(drop, ctr + 1);
                    os.z *= drop_mult(drop, ctr + 2); os.w *= drop_mult(drop, ctr + 3);
                }
                if (dx_bf16) st4(dx_bf16 + (int64_t)row * lddxb + 4 * i, os);
                if constexpr (kColsum) { ac[j].x += os.x; ac[j].y += os.y; ac[j].z += os.z; ac[j].w += os.w; }
            }
        }
    }
    float* out = partials + (size_t)blockIdx.x * (kColsum ? 3 : 2) * D;
#pragma unroll
    for (int j = 0; j < kV2; ++j) {
        const int i = threadIdx.x + kLnThreads * j;
        if (i < nvec) {
            st4(out + 4 * i, ag[j]);
            st4(out + D + 4 * i, ab[j]);
            if constexpr (kColsum) st4(out + 2 * D + 4 * i, ac[j]);
        }
    }
}

// partials[nblocks][nsets*D] -> dgamma, dbeta [, colsum].  Block = 32 columns x 32 row-slices; coalesced 128-byte reads.
__global__ void __launch_bounds__(1024)
ln_bwd_finalize_kernel(const float* __restrict__ partials, int nblocks, int D, int nsets, float* __restrict__ dgamma,
                       float* __restrict__ dbeta, float* __restrict__ colsum, int accumulate) {
    __shared__ float red[32][33];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int i = blockIdx.x * 32 + tx;   // over nsets*D
    const int n = nsets * D;
    float a = 0.f;
    if (i < n)
        for (int b = ty; b < nblocks; b += 32) a += partials[(size_t)b * n + i];
    red[ty][tx] = a;
    __syncthreads();
    if (ty == 0 && i < n) {
        float t = 0.f;
#pragma unroll
        for (int k = 0; k < 32; ++k) t += red[k][tx];
        if (i >= 2 * D) { colsum[i - 2 * D] = t; return; }
        float* dst = (i < D) ? (dgamma + i) : (dbeta + (i - D));
        *dst = accumulate ? (*dst + t) : t;
    }
}

int ln_bwd_grid(int M) {
    const int want = ceil_div(M, kLnWarps);
    const int cap = num_sms() * 2;   // two resident blocks per SM (register-limited): exactly one wave
    return want < cap ? want : cap;
}
// ---------------------------------------------------------------------------------------------- any width
// Rows wider than the register-resident kernels hold (D > 1536: Swin-B's patch-merging LayerNorm over 4C = 2048,
// swin.py:329) or not a multiple of 4 (EVA02-L's SwiGLU LayerNorm over 2730, eva_vit_model.py:213): one 256-thread block
// per row, scalar accesses, two-pass statistics.  Off the hot path of every measured configuration.
constexpr int kLnGenThreads = 256;

__device__ __forceinline__ float block_sum_256(float v, float* red) {
    v = warp_sum(v);
    __syncthreads();
    if (lane_id() == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < kLnGenThreads / 32; ++w) t += red[w];
    return t;
}

template <typename TX>
__global__ void __launch_bounds__(kLnGenThreads)
ln_fwd_generic_kernel(const TX* __restrict__ x, int64_t ldx, const float* __restrict__ gamma, const float* __restrict__ beta,
                      __nv_bfloat16* __restrict__ yb, float* __restrict__ yf, int64_t ldy, float* __restrict__ mean,
                      float* __restrict__ rstd, int D, float eps) {
    __shared__ float red[kLnGenThreads / 32];
    const int64_t row = blockIdx.x;
    const TX* xr = x + row * ldx;
    float s = 0.f;
    for (int c = threadIdx.x; c < D; c += kLnGenThreads) s += (float)xr[c];
    const float mu = block_sum_256(s, red) / D;
    float q = 0.f;
    for (int c = threadIdx.x; c < D; c += kLnGenThreads) { const float d = (float)xr[c] - mu; q += d * d; }
    const float rs = rsqrtf(block_sum_256(q, red) / D + eps);
    for (int c = threadIdx.x; c < D; c += kLnGenThreads) {
        const float y = ((float)xr[c] - mu) * rs * gamma[c] + beta[c];
        if (yb) yb[row * ldy + c] = __float2bfloat16(y);
        if (yf) yf[row * ldy + c] = y;
    }
    if (threadIdx.x == 0) {
        if (mean) mean[row] = mu;
        if (rstd) rstd[row] = rs;
    }
}

// dx = [dres +] rstd * (g - mean(g) - xhat * mean(g * xhat)), g = (dy [+ dy2]) * gamma; dgamma += sum dy*xhat, dbeta += sum dy
// (fp32 atomics into zeroed / accumulated buffers: the summation order over rows is not fixed)
template <typename TDY>
__global__ void __launch_bounds__(kLnGenThreads)
ln_bwd_generic_kernel(const TDY* __restrict__ dy, int64_t lddy, const __nv_bfloat16* __restrict__ dy2, int64_t lddy2,
                      const float* __restrict__ x, int64_t ldx, const float* __restrict__ mean, const float* __restrict__ rstd,
                      const float* __restrict__ gamma, const float* __restrict__ dres, int64_t lddres, float* __restrict__ dx,
                      int64_t lddx, __nv_bfloat16* __restrict__ dxb, int64_t lddxb, const float* __restrict__ row_scale,
                      int rows_per_group, float* __restrict__ dgamma, float* __restrict__ dbeta, int D) {
    __shared__ float red[kLnGenThreads / 32];
    const int64_t row = blockIdx.x;
    const float mu = mean[row], rs = rstd[row];
    auto up = [&](int c) { return (float)dy[row * lddy + c] + (dy2 ? __bfloat162float(dy2[row * lddy2 + c]) : 0.f); };
    float s1 = 0.f, s2 = 0.f;
    for (int c = threadIdx.x; c < D; c += kLnGenThreads) {
        const float u = up(c), xh = (x[row * ldx + c] - mu) * rs, g = u * gamma[c];
        s1 += g;
        s2 += g * xh;
        atomicAdd(dgamma + c, u * xh);
        atomicAdd(dbeta + c, u);
    }
    const float m1 = block_sum_256(s1, red) / D;
    const float m2 = block_sum_256(s2, red) / D;
    const float sc = row_scale ? row_scale[row / rows_per_group] : 1.0f;
    for (int c = threadIdx.x; c < D; c += kLnGenThreads) {
        const float xh = (x[row * ldx + c] - mu) * rs;
        float v = rs * (up(c) * gamma[c] - m1 - xh * m2);
        if (dres) v += dres[row * lddres + c];
        if (dx) dx[row * lddx + c] = v;
        if (dxb) dxb[row * lddxb + c] = __float2bfloat16(v * sc);
    }
}

bool ln_needs_generic(int D, int64_t a, int64_t b, int64_t c, int64_t d, int64_t e, int64_t f) {
    return D > kMaxVec * 128 || D % 4 != 0 || ((a | b | c | d | e | f) & 3) != 0;
}

bool ln_bwd_use_rows(int D) { return (D >> 2) >= kLnThreads && (D >> 2) <= kV2 * kLnThreads; }
int ln_bwd_row_grid(int M) {
    const int cap = num_sms() * 5;   // five resident blocks per SM: one wave
    return M < cap ? M : cap;
}

}  // namespace
}  // namespace mico

extern "C" int mico_layernorm_fwd(const void* x, int x_is_bf16, int64_t ldx, const float* gamma, const float* beta,
                                  void* y_bf16, float* y_f32, int64_t ldy, float* mean, float* rstd, int M, int D,
                                  float eps, void* stream_) {
    using namespace mico;
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    MICO_CHECK_ARG(x && gamma && beta && (y_bf16 || y_f32));
    MICO_CHECK_ARG(M > 0 && D > 0);
    ProfScope prof(kProfLnFwd, (double)M * D * ((x_is_bf16 ? 2 : 4) + (y_bf16 ? 2 : 0) + (y_f32 ? 4 : 0)), stream);
    if (ln_needs_generic(D, ldx, ldy, 0, 0, 0, 0)) {
        if (x_is_bf16)
            ln_fwd_generic_kernel<__nv_bfloat16><<<M, kLnGenThreads, 0, stream>>>(
                reinterpret_cast<const __nv_bfloat16*>(x), ldx, gamma, beta, reinterpret_cast<__nv_bfloat16*>(y_bf16), y_f32, ldy,
                mean, rstd, D, eps);
        else
            ln_fwd_generic_kernel<float><<<M, kLnGenThreads, 0, stream>>>(reinterpret_cast<const float*>(x), ldx, gamma, beta,
                                                                          reinterpret_cast<__nv_bfloat16*>(y_bf16), y_f32, ldy,
                                                                          mean, rstd, D, eps);
        MICO_CHECK_CUDA(cudaGetLastError());
        count_launch();
        return MICO_OK;
    }
    {   // TMA-staged rows when the ring fits in shared memory and the rows can be bulk-copied (16-byte aligned)
        const int row_bytes = D * (x_is_bf16 ? 2 : 4);
        const int padded = (row_bytes + 127) & ~127;
        const size_t smem = (size_t)kLnBulkWarps * kLnStages * padded + kLnBulkWarps * kLnStages * sizeof(uint64_t);
        const bool aligned = row_bytes % 16 == 0 && (ldx * (x_is_bf16 ? 2 : 4)) % 16 == 0 &&
                             (reinterpret_cast<uintptr_t>(x) & 15) == 0;
        if (aligned && smem <= 220 * 1024 && M >= 4 * kLnBulkWarps * kLnStages) {
            const int want = ceil_div(M, kLnBulkWarps);
            const int grid = want < num_sms() ? want : num_sms();
            if (x_is_bf16) {
                auto k = ln_fwd_bulk_kernel<__nv_bfloat16>;
                MICO_CHECK_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                k<<<grid, kLnBulkWarps * 32, smem, stream>>>(reinterpret_cast<const __nv_bfloat16*>(x), ldx, gamma, beta,
                                                           reinterpret_cast<__nv_bfloat16*>(y_bf16), y_f32, ldy, mean, rstd, M, D,
                                                           eps, padded);
            } else {
                auto k = ln_fwd_bulk_kernel<float>;
                MICO_CHECK_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                k<<<grid, kLnBulkWarps * 32, smem, stream>>>(reinterpret_cast<const float*>(x), ldx, gamma, beta,
                                                           reinterpret_cast<__nv_bfloat16*>(y_bf16), y_f32, ldy, mean, rstd, M, D,
                                                           eps, padded);
            }
            MICO_CHECK_CUDA(cudaGetLastError());
            count_launch();
            return MICO_OK;
        }
    }
    // one row per warp, no grid cap: with ~2 rows per warp a capped grid ends in a half-empty second pass; the block
    // scheduler balances 4k small blocks better than a strided loop does
    const int grid = ceil_div(M, kLnWarps);
    if (x_is_bf16)
        ln_fwd_kernel<__nv_bfloat16><<<grid, kLnThreads, 0, stream>>>(
            reinterpret_cast<const __nv_bfloat16*>(x), ldx, gamma, beta, reinterpret_cast<__nv_bfloat16*>(y_bf16), y_f32,
            ldy, mean, rstd, M, D, eps);
    else
        ln_fwd_kernel<float><<<grid, kLnThreads, 0, stream>>>(reinterpret_cast<const float*>(x), ldx, gamma, beta,
                                                              reinterpret_cast<__nv_bfloat16*>(y_bf16), y_f32, ldy,
                                                              mean, rstd, M, D, eps);
    MICO_CHECK_CUDA(cudaGetLastError());
    count_launch();
    return MICO_OK;
}

extern "C" size_t mico_layernorm_bwd_workspace(int M, int D) {
    using namespace mico;
    if (D > kMaxVec * 128 || D % 4 != 0) return 16;      // the any-width kernels need no workspace
    if (ln_bwd_use_rows(D)) return (size_t)ln_bwd_row_grid(M) * 3 * (size_t)D * sizeof(float);
    return (size_t)ln_bwd_grid(M) * 2 * (size_t)D * sizeof(float);
}

static int layernorm_bwd_impl(const void* dy, int dy_is_bf16, int64_t lddy, const void* dy2_bf16, int64_t lddy2,
                              const float* x, int64_t ldx,
                              const float* mean, const float* rstd, const float* gamma, const float* dres,
                              int64_t lddres, float* dx, int64_t lddx, void* dx_bf16, int64_t lddxb,
                              const float* row_scale, int rows_per_group, float* dgamma, float* dbeta,
                              int accumulate_param_grads, float* dxb_colsum, int M, int D, void* workspace,
                              size_t ws_bytes, void* stream_, float drop_p, uint64_t drop_seed, uint64_t drop_site) {
    using namespace mico;
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    MICO_CHECK_ARG(dy && x && mean && rstd && gamma && dgamma && dbeta && workspace);
    MICO_CHECK_ARG(drop_p >= 0.0f && drop_p < 1.0f);
    const bool dropping = drop_p > 0.0f;
    if (dropping && (!dx_bf16 || !ln_bwd_use_rows(D) || ln_needs_generic(D, lddy, ldx, lddx, lddxb, lddres, lddy2))) {
        set_last_error(__FILE__, __LINE__, "layernorm_bwd: the fused dropout mask needs the block-per-row kernel (512 <= D <= 1536, "
                                           "aligned rows) and a bf16 output");
        return MICO_ERR_UNSUPPORTED;
    }
    DropCfg dcfg;
    dcfg.p = drop_p; dcfg.inv_keep = 1.0f / (1.0f - drop_p); dcfg.seed = drop_seed;
    MICO_CHECK_ARG(dx || dx_bf16);
    MICO_CHECK_ARG(M > 0 && D > 0);
    const __nv_bfloat16* dy2 = reinterpret_cast<const __nv_bfloat16*>(dy2_bf16);
    MICO_CHECK_ARG(!(row_scale && rows_per_group <= 0));
    if (ln_needs_generic(D, lddy, ldx, lddx, lddxb, lddres, lddy2)) {
        MICO_CHECK_ARG(!dxb_colsum);
        ProfScope prof(kProfLnBwd, (double)M * D * ((dy_is_bf16 ? 2 : 4) + 4 + (dres ? 4 : 0) + (dx ? 4 : 0) + (dx_bf16 ? 2 : 0)), stream);
        if (!accumulate_param_grads) {
            MICO_CHECK_CUDA(cudaMemsetAsync(dgamma, 0, (size_t)D * sizeof(float), stream));
            MICO_CHECK_CUDA(cudaMemsetAsync(dbeta, 0, (size_t)D * sizeof(float), stream));
        }
        auto go = [&](auto k, auto dyp) {
            k<<<M, kLnGenThreads, 0, stream>>>(dyp, lddy, dy2, lddy2, x, ldx, mean, rstd, gamma, dres, lddres, dx, lddx,
                                              reinterpret_cast<__nv_bfloat16*>(dx_bf16), lddxb, row_scale, rows_per_group, dgamma,
                                              dbeta, D);
        };
        if (dy_is_bf16) go(ln_bwd_generic_kernel<__nv_bfloat16>, reinterpret_cast<const __nv_bfloat16*>(dy));
        else go(ln_bwd_generic_kernel<float>, reinterpret_cast<const float*>(dy));
        MICO_CHECK_CUDA(cudaGetLastError());
        count_launch();
        return MICO_OK;
    }
    MICO_CHECK_ARG(ws_bytes >= mico_layernorm_bwd_workspace(M, D));
    const bool rows = ln_bwd_use_rows(D);
    // the fused column sum lives in the block-per-row kernel only (every tower width on the path)
    MICO_CHECK_ARG(!(dxb_colsum && !rows));
    ProfScope prof(kProfLnBwd, (double)M * D * ((dy_is_bf16 ? 2 : 4) + 4 + (dres ? 4 : 0) + (dx ? 4 : 0) + (dx_bf16 ? 2 : 0)),
                   stream);
    float* partials = reinterpret_cast<float*>(workspace);
    __nv_bfloat16* dxb = reinterpret_cast<__nv_bfloat16*>(dx_bf16);
    int grid;
    if (rows) {
        grid = ln_bwd_row_grid(M);
        auto go = [&](auto k, auto dyp) {
            k<<<grid, kLnThreads, 0, stream>>>(dyp, lddy, dy2, lddy2, x, ldx, mean, rstd, gamma, dres, lddres, dx, lddx, dxb,
                                               lddxb, row_scale, rows_per_group, partials, M, D, dcfg, drop_site);
        };
        if (dropping) {      // post-LN BERT sub-layers: fp32 upstream gradient, masked bf16 copy, optional column sums
            MICO_CHECK_ARG(!dy_is_bf16);
            const float* p = reinterpret_cast<const float*>(dy);
            if (dxb_colsum) go(ln_bwd_row_kernel<float, true, true>, p); else go(ln_bwd_row_kernel<float, false, true>, p);
        } else if (dy_is_bf16) {
            const __nv_bfloat16* p = reinterpret_cast<const __nv_bfloat16*>(dy);
            if (dxb_colsum) go(ln_bwd_row_kernel<__nv_bfloat16, true>, p); else go(ln_bwd_row_kernel<__nv_bfloat16, false>, p);
        } else {
            const float* p = reinterpret_cast<const float*>(dy);
            if (dxb_colsum) go(ln_bwd_row_kernel<float, true>, p); else go(ln_bwd_row_kernel<float, false>, p);
        }
    } else {
        grid = ln_bwd_grid(M);
        const size_t smem = (size_t)kLnWarps * 2 * D * sizeof(float);
        if (dy_is_bf16) {
            auto k = ln_bwd_kernel<__nv_bfloat16>;
            if (smem > 48 * 1024) MICO_CHECK_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            k<<<grid, kLnThreads, smem, stream>>>(reinterpret_cast<const __nv_bfloat16*>(dy), lddy, dy2, lddy2, x, ldx, mean, rstd,
                                                  gamma, dres, lddres, dx, lddx, dxb, lddxb, row_scale, rows_per_group, partials,
                                                  M, D);
        } else {
            auto k = ln_bwd_kernel<float>;
            if (smem > 48 * 1024) MICO_CHECK_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            k<<<grid, kLnThreads, smem, stream>>>(reinterpret_cast<const float*>(dy), lddy, dy2, lddy2, x, ldx, mean, rstd, gamma,
                                                  dres, lddres, dx, lddx, dxb, lddxb, row_scale, rows_per_group, partials, M, D);
        }
    }
    MICO_CHECK_CUDA(cudaGetLastError());
    const int nsets = (rows && dxb_colsum) ? 3 : 2;
    ln_bwd_finalize_kernel<<<ceil_div(nsets * D, 32), 1024, 0, stream>>>(partials, grid, D, nsets, dgamma, dbeta, dxb_colsum,
                                                                        accumulate_param_grads);
    MICO_CHECK_CUDA(cudaGetLastError());
    count_launch(2);
    return MICO_OK;
}

extern "C" int mico_layernorm_bwd(const void* dy, int dy_is_bf16, int64_t lddy, const void* dy2_bf16, int64_t lddy2,
                                  const float* x, int64_t ldx,
                                  const float* mean, const float* rstd, const float* gamma, const float* dres,
                                  int64_t lddres, float* dx, int64_t lddx, void* dx_bf16, int64_t lddxb,
                                  const float* row_scale, int rows_per_group, float* dgamma, float* dbeta,
                                  int accumulate_param_grads, float* dxb_colsum, int M, int D, void* workspace,
                                  size_t ws_bytes, void* stream_) {
    return layernorm_bwd_impl(dy, dy_is_bf16, lddy, dy2_bf16, lddy2, x, ldx, mean, rstd, gamma, dres, lddres, dx, lddx, dx_bf16,
                              lddxb, row_scale, rows_per_group, dgamma, dbeta, accumulate_param_grads, dxb_colsum, M, D,
                              workspace, ws_bytes, stream_, 0.0f, 0, 0);
}

extern "C" int mico_layernorm_bwd_dropout(const void* dy, int dy_is_bf16, int64_t lddy, const void* dy2_bf16, int64_t lddy2,
                                          const float* x, int64_t ldx,
                                          const float* mean, const float* rstd, const float* gamma, const float* dres,
                                          int64_t lddres, float* dx, int64_t lddx, void* dx_bf16, int64_t lddxb,
                                          const float* row_scale, int rows_per_group, float* dgamma, float* dbeta,
                                          int accumulate_param_grads, float* dxb_colsum, int M, int D, void* workspace,
                                          size_t ws_bytes, float drop_p, uint64_t drop_seed, uint64_t drop_site,
                                          void* stream_) {
    return layernorm_bwd_impl(dy, dy_is_bf16, lddy, dy2_bf16, lddy2, x, ldx, mean, rstd, gamma, dres, lddres, dx, lddx, dx_bf16,
                              lddxb, row_scale, rows_per_group, dgamma, dbeta, accumulate_param_grads, dxb_colsum, M, D,
                              workspace, ws_bytes, stream_, drop_p, drop_seed, drop_site);
}

