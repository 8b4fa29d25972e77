// mico_b200 -- K3: persistent warp-specialised bf16 GEMM for sm_100a.
//
//   out[M,N] = epilogue(alpha * A[M,K] . B[N,K]^T)      fp32 accumulation in TMEM
//
// Structure (one CTA per SM, 320 threads; CTAs paired in clusters of two along M):
//   warp 0      TMA producer: cp.async.bulk.tensor 128B-swizzled tiles -> STAGES-deep smem ring
//   warp 1      MMA issuer:   one elected lane issues tcgen05.mma (K = 16) x4 per stage; tcgen05.commit releases smem
//                             slots and publishes the accumulator.  Pair variants:
//                               CL = 3  ONE tcgen05.mma.cta_group::2 (M = 256) per pair, issued by the leader CTA; each CTA
//                                       stages its own 128 A rows and its half of the B tile; barriers live in the leader
//                                       (default for 256- and 128-wide tiles, N % 128 == 0)
//                               CL = 2  each CTA issues its own M = 128 MMA; the B tile is shared by TMA multicast
//                               CL = 1  single CTA (odd shapes, one M tile)
//   warps 2..9  epilogue:     tcgen05.ld TMEM -> registers -> fused bias / GELU (+ GELU' store) / x saved GELU' / DropPath row
//                             scale / + fp32 residual -> TMA tile stores (specialised kernels, template EPI) or the generic
//                             run-time epilogue (two warps per TMEM lane quarter, alternating 32-column chunks)
// Two TMEM accumulator stages (2 x 256 columns) let the epilogue of tile i overlap the MMAs of tile i+1.  The last N tile of
// a row issues a narrower MMA, so N = 1408 runs on 256-wide tiles (5 full + 1 half).
// Work units are dealt to the CTA pairs by a balanced schedule table (GemmPlan below: per round, longest-first to the least
// loaded pair; split-K of a weight gradient by planned makespan); a weight gradient can also return its layer's bias gradient
// from 32 ones columns behind the last N tile (asum_out).
// Either operand may be "MN-major" (the contraction index is the slow dimension in HBM); that is how
// wgrad (dW = dY^T X) and dgrad (dX = dY W) run on the same kernel with no transposes in HBM.
//
// Replaces F.linear/matmul call sites listed in include/mico_b200.h (reference: eva_vit_model.py:191,
// 197, 310, 363, 446; bert.py:196-209, 293, 357, 370, 601, 607).
#include "common.cuh"
#include <stdlib.h>
#include <algorithm>
#include <array>
#include <map>
#include <mutex>
#include <vector>

#include "host_utils.h"

namespace mico {

namespace {

constexpr int BM = 128;
constexpr int BK = 64;
constexpr int kGemmThreads = 320;
constexpr int kAccStride = 256;   // TMEM columns between the two accumulator stages

struct GemmEpi {
    void* out;
    int64_t ldo;
    int out_fp32;
    const float* bias;
    const float* residual;
    int64_t ldr;
    const float* row_scale;
    int rows_per_group;
    int act;
    __nv_bfloat16* aux_out;
    int64_t ld_aux_out;
    const __nv_bfloat16* aux_in;
    int64_t ld_aux_in;
    int accumulate;
    float alpha;
    int remap_gin, remap_gout, remap_off, residual_bcast;
    int vec_ok;   // all pitches / bases allow 16-byte vector access
    int ksplit;   // > 1: the K loop of every output tile is split over `ksplit` work units whose fp32 partial tiles are
                  // ADDED into the (zeroed) output by TMA reduce (EPI_F32 only: weight gradients with few output tiles)
    float* asum_out;    // weight gradients: asum_out[m] += sum_k A(m,k) (the layer's bias gradient) from 32 extra MMA columns of
                        // the last N tile, whose B operand is a tile of ones (tmAux); null = off
    const int* sched;   // balanced unit schedule (plan_schedule below): slot s runs sched[s * sched_rounds + it], -1 ends its
    int sched_rounds;   // list; null = the strided default (slot s runs units s, s + slots, ...)
};

template <int BN, int STAGES, int EPI = 0, bool B_MN = false, bool P2 = false>
struct GemmCfg {
    static constexpr int A_BYTES = BM * BK * 2;
    // an MN-major B tile is loaded as [64 k x 64 n] boxes: a 176-wide tile takes three (the last box reaches 16 columns
    // into the neighbouring tile, or is zero-filled at the edge; the N = 176 MMA never reads them)
    static constexpr int B_CHUNKS = (BN + 63) / 64;
    // P2 (one cta_group::2 MMA over the pair): each CTA holds only ITS half of the B tile (BN/2 rows)
    static constexpr int B_BYTES = (B_MN ? B_CHUNKS * 8192 : BN * BK * 2) / (P2 ? 2 : 1);
    static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
    // per epilogue warp: generic = one [32 x 64 B] transpose tile; specialised = two such tiles, 64B-swizzled, that
    // TMA stores read (512-byte aligned: they sit right behind the 1024-aligned operand stages)
    static constexpr int STAGE_TILE = EPI != 0 ? 4096 : 2048;
    static constexpr int STAGING_BYTES = 8 * STAGE_TILE;
    static constexpr int BAR_BYTES = 256;            // (2 * STAGES + 4) mbarriers + the TMEM slot: <= 20 x 8 + 8
    static constexpr int BIAS_BYTES = 2 * 256 * 4;   // bias slice of the tile, one copy per accumulator stage
    // specialised kernels rely on the 1024-byte alignment of the dynamic shared window (checked at kernel entry):
    // 4 x 48 KB stages + 32 KB of store tiles leave no room for alignment slack
    static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + STAGING_BYTES + BAR_BYTES + BIAS_BYTES + (EPI != 0 ? 0 : 1024);
};
static_assert(GemmCfg<256, 6, 1, false, true>::SMEM_BYTES <= 232448 && GemmCfg<128, 8, 1, true, true>::SMEM_BYTES <= 232448 &&
              GemmCfg<256, 4, 1>::SMEM_BYTES <= 232448 && GemmCfg<176, 5, 1>::SMEM_BYTES <= 232448 &&
              GemmCfg<176, 4, 1, true>::SMEM_BYTES <= 232448, "shared memory budget");

// Global operands of one epilogue chunk, requested BEFORE the TMEM load is waited for so that their latency
// overlaps it (ncu round 1: the epilogue warps sat in long-scoreboard stalls on exactly these loads).
// Holds the fp32 residual (8 x 16 B) or, when there is no residual, the bf16 pre-activation of *_BWD (4 x 16 B).
struct EpiOperands {
    uint4 pre[8];
    float rs;
    int64_t orow, rrow;
    bool full;
};

template <int NC>
__device__ __forceinline__ void epilogue_prefetch(const GemmEpi& e, EpiOperands& o, int row, int col0, int M, int N) {
    o.orow = row;
    o.rrow = row;
    if (e.remap_gin > 0) {
        const int g = row / e.remap_gin, r = row - g * e.remap_gin;
        o.orow = (int64_t)g * e.remap_gout + r + e.remap_off;
        o.rrow = e.residual_bcast ? (int64_t)(r + e.remap_off) : o.orow;
    }
    o.full = (col0 + NC <= N) && e.vec_ok;
    o.rs = 1.0f;
    if (row >= M) return;
    if (e.row_scale) o.rs = __ldg(e.row_scale + row / e.rows_per_group);
    if (!o.full) return;
    if (e.residual) {
        const uint4* r4 = reinterpret_cast<const uint4*>(e.residual + o.rrow * e.ldr + col0);
#pragma unroll
        for (int i = 0; i < NC / 4; ++i) o.pre[i] = __ldg(r4 + i);
    } else if (e.act == MICO_ACT_GELU_BWD || e.act == MICO_ACT_QUICK_GELU_BWD || e.act == MICO_ACT_MUL_AUX) {
        const uint4* u4 = reinterpret_cast<const uint4*>(e.aux_in + o.orow * e.ld_aux_in + col0);
#pragma unroll
        for (int i = 0; i < NC / 8; ++i) o.pre[i] = __ldg(u4 + i);
    }
}

template <int NC>
__device__ __forceinline__ void epilogue_chunk(const GemmEpi& e, const EpiOperands& o, const float* sbias,
                                               const uint32_t (&acc)[NC], int row, int col0, int M, int N) {
    if (row >= M) return;
    float v[NC];
#pragma unroll
    for (int i = 0; i < NC; ++i) v[i] = __uint_as_float(acc[i]) * e.alpha;
    const int64_t orow = o.orow, rrow = o.rrow;
    const float rs = o.rs;

    if (o.full) {
        if (e.bias) {
#pragma unroll
            for (int i = 0; i < NC / 4; ++i) {
                const float4 b = *reinterpret_cast<const float4*>(sbias + 4 * i);
                v[4 * i + 0] += b.x; v[4 * i + 1] += b.y; v[4 * i + 2] += b.z; v[4 * i + 3] += b.w;
            }
        }
        if (e.act == MICO_ACT_GELU_SAVE_GRAD || e.act == MICO_ACT_QUICK_GELU_SAVE_GRAD) {
            float gr[NC];
#pragma unroll
            for (int i = 0; i < NC; ++i) {
                gr[i] = e.act == MICO_ACT_GELU_SAVE_GRAD ? gelu_erf_grad(v[i]) : quick_gelu_grad(v[i]);
                v[i] = e.act == MICO_ACT_GELU_SAVE_GRAD ? gelu_erf(v[i]) : quick_gelu(v[i]);
            }
            if (e.aux_out) {
                uint4* a4 = reinterpret_cast<uint4*>(e.aux_out + orow * e.ld_aux_out + col0);
#pragma unroll
                for (int i = 0; i < NC / 8; ++i)
                    a4[i] = make_uint4(pack_bf16x2(gr[8 * i], gr[8 * i + 1]), pack_bf16x2(gr[8 * i + 2], gr[8 * i + 3]),
                                       pack_bf16x2(gr[8 * i + 4], gr[8 * i + 5]), pack_bf16x2(gr[8 * i + 6], gr[8 * i + 7]));
            }
        } else if (e.act == MICO_ACT_MUL_AUX) {
            const uint4* u4 = reinterpret_cast<const uint4*>(e.aux_in + orow * e.ld_aux_in + col0);
#pragma unroll
            for (int i = 0; i < NC / 8; ++i) {
                const uint4 u = e.residual ? __ldg(u4 + i) : o.pre[i];
                const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    v[8 * i + 2 * j] *= bf16_lo(w[j]);
                    v[8 * i + 2 * j + 1] *= bf16_hi(w[j]);
                }
            }
        } else if (e.act == MICO_ACT_GELU || e.act == MICO_ACT_QUICK_GELU) {
            if (e.aux_out) {
                uint4* a4 = reinterpret_cast<uint4*>(e.aux_out + orow * e.ld_aux_out + col0);
#pragma unroll
                for (int i = 0; i < NC / 8; ++i)
                    a4[i] = make_uint4(pack_bf16x2(v[8 * i], v[8 * i + 1]), pack_bf16x2(v[8 * i + 2], v[8 * i + 3]),
                                       pack_bf16x2(v[8 * i + 4], v[8 * i + 5]), pack_bf16x2(v[8 * i + 6], v[8 * i + 7]));
            }
            if (e.act == MICO_ACT_GELU) {
#pragma unroll
                for (int i = 0; i < NC; ++i) v[i] = gelu_erf(v[i]);
            } else {
#pragma unroll
                for (int i = 0; i < NC; ++i) v[i] = quick_gelu(v[i]);
            }
        } else if (e.act == MICO_ACT_GELU_BWD || e.act == MICO_ACT_QUICK_GELU_BWD) {
            const uint4* u4 = reinterpret_cast<const uint4*>(e.aux_in + orow * e.ld_aux_in + col0);
#pragma unroll
            for (int i = 0; i < NC / 8; ++i) {
                const uint4 u = e.residual ? __ldg(u4 + i) : o.pre[i];
                const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float lo = bf16_lo(w[j]), hi = bf16_hi(w[j]);
                    if (e.act == MICO_ACT_GELU_BWD) {
                        v[8 * i + 2 * j] *= gelu_erf_grad(lo);
                        v[8 * i + 2 * j + 1] *= gelu_erf_grad(hi);
                    } else {
                        v[8 * i + 2 * j] *= quick_gelu_grad(lo);
                        v[8 * i + 2 * j + 1] *= quick_gelu_grad(hi);
                    }
                }
            }
        }
        if (e.row_scale) {
#pragma unroll
            for (int i = 0; i < NC; ++i) v[i] *= rs;
        }
        if (e.residual) {
#pragma unroll
            for (int i = 0; i < NC / 4; ++i) {
                const uint4 r = o.pre[i];
                v[4 * i + 0] += __uint_as_float(r.x); v[4 * i + 1] += __uint_as_float(r.y);
                v[4 * i + 2] += __uint_as_float(r.z); v[4 * i + 3] += __uint_as_float(r.w);
            }
        }
        if (e.out_fp32) {
            float4* o4 = reinterpret_cast<float4*>(reinterpret_cast<float*>(e.out) + orow * e.ldo + col0);
            if (e.accumulate) {
#pragma unroll
                for (int i = 0; i < NC / 4; ++i) {
                    const float4 p = o4[i];
                    v[4 * i + 0] += p.x; v[4 * i + 1] += p.y; v[4 * i + 2] += p.z; v[4 * i + 3] += p.w;
                }
            }
#pragma unroll
            for (int i = 0; i < NC / 4; ++i) o4[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
        } else {
            uint4* o4 = reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(e.out) + orow * e.ldo + col0);
#pragma unroll
            for (int i = 0; i < NC / 8; ++i)
                o4[i] = make_uint4(pack_bf16x2(v[8 * i], v[8 * i + 1]), pack_bf16x2(v[8 * i + 2], v[8 * i + 3]),
                                   pack_bf16x2(v[8 * i + 4], v[8 * i + 5]), pack_bf16x2(v[8 * i + 6], v[8 * i + 7]));
        }
        return;
    }
    // ragged / unaligned tail: scalar path
#pragma unroll
    for (int i = 0; i < NC; ++i) {
        const int col = col0 + i;
        if (col >= N) break;
        float x = v[i];
        if (e.bias) x += sbias[i];
        if (e.act == MICO_ACT_GELU_SAVE_GRAD || e.act == MICO_ACT_QUICK_GELU_SAVE_GRAD) {
            const bool erf_gelu = e.act == MICO_ACT_GELU_SAVE_GRAD;
            if (e.aux_out)
                e.aux_out[orow * e.ld_aux_out + col] = __float2bfloat16(erf_gelu ? gelu_erf_grad(x) : quick_gelu_grad(x));
            x = erf_gelu ? gelu_erf(x) : quick_gelu(x);
        } else if (e.act == MICO_ACT_MUL_AUX) {
            x *= __bfloat162float(e.aux_in[orow * e.ld_aux_in + col]);
        } else if (e.act == MICO_ACT_GELU || e.act == MICO_ACT_QUICK_GELU) {
            if (e.aux_out) e.aux_out[orow * e.ld_aux_out + col] = __float2bfloat16(x);
            x = (e.act == MICO_ACT_GELU) ? gelu_erf(x) : quick_gelu(x);
        } else if (e.act == MICO_ACT_GELU_BWD) {
            x *= gelu_erf_grad(__bfloat162float(e.aux_in[orow * e.ld_aux_in + col]));
        } else if (e.act == MICO_ACT_QUICK_GELU_BWD) {
            x *= quick_gelu_grad(__bfloat162float(e.aux_in[orow * e.ld_aux_in + col]));
        }
        x *= rs;
        if (e.residual) x += e.residual[rrow * e.ldr + col];
        if (e.out_fp32) {
            float* po = reinterpret_cast<float*>(e.out) + orow * e.ldo + col;
            if (e.accumulate) x += *po;
            *po = x;
        } else {
            reinterpret_cast<__nv_bfloat16*>(e.out)[orow * e.ldo + col] = __float2bfloat16(x);
        }
    }
}

// ------------------------------------------------------------------------------------------------------------
// Coalesced epilogue I/O.  A TMEM lane is an output ROW, so a thread naturally owns 32 consecutive columns of one
// row and a warp-wide 16-byte access touches 32 different rows = 32 cache lines: the L1/LSU serialises them (one
// line per cycle), which made every epilogue with extra operands LSU-bound (ncu round 1: fc1+GELU 49 % and
// fc2-dgrad+GELU' 44 % tensor-active vs 83 % for a plain store).  Each epilogue warp therefore owns a 2 KB staging
// tile [32 rows x 64 B] in shared memory (16-byte groups XOR-swizzled by (row >> 1) & 3: conflict-free both ways)
// and moves 64-byte row segments with the mapping  lane l <-> (row 8i + l/4, group l%4), i = 0..3:
// 8 rows x 64 contiguous bytes per instruction instead of 32 rows x 16 bytes.
__device__ __forceinline__ uint32_t stage_at(uint32_t st, int row, int g) {
    return st + row * 64 + ((g ^ ((row >> 1) & 3)) << 4);
}
__device__ __forceinline__ int64_t map_out_row(const GemmEpi& e, int row) {
    if (e.remap_gin <= 0) return row;
    const int g = row / e.remap_gin, r = row - g * e.remap_gin;
    return (int64_t)g * e.remap_gout + r + e.remap_off;
}
__device__ __forceinline__ int64_t map_res_row(const GemmEpi& e, int row) {
    if (e.remap_gin <= 0) return row;
    const int g = row / e.remap_gin, r = row - g * e.remap_gin;
    return e.residual_bcast ? (int64_t)(r + e.remap_off) : (int64_t)g * e.remap_gout + r + e.remap_off;
}
// own row (4 x 16 B in `own`) -> staging -> global rows base + maprow(row)*pitch_bytes + byte_off
template <bool kResRows>
__device__ __forceinline__ void staged_store(const GemmEpi& e, uint32_t st, const uint4 (&own)[4], uint8_t* base,
                                             int64_t pitch_bytes, int64_t byte_off, int row0, int M) {
    const int l = (int)lane_id();
#pragma unroll
    for (int g = 0; g < 4; ++g) sts128(stage_at(st, l, g), own[g]);
    __syncwarp();
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int rr = 8 * i + (l >> 2), g = l & 3, row = row0 + rr;
        const uint4 v = lds128(stage_at(st, rr, g));
        if (row < M) {
            const int64_t mr = kResRows ? map_res_row(e, row) : map_out_row(e, row);
            *reinterpret_cast<uint4*>(base + mr * pitch_bytes + byte_off + g * 16) = v;
        }
    }
    __syncwarp();
}
// coalesced registers (pre[i] <-> row 8i + l/4, group l%4) -> staging -> own row
__device__ __forceinline__ void staged_to_own(uint32_t st, const uint4* pre, uint4 (&own)[4]) {
    const int l = (int)lane_id();
#pragma unroll
    for (int i = 0; i < 4; ++i) sts128(stage_at(st, 8 * i + (l >> 2), l & 3), pre[i]);
    __syncwarp();
#pragma unroll
    for (int g = 0; g < 4; ++g) own[g] = lds128(stage_at(st, l, g));
    __syncwarp();
}
template <bool kResRows>
__device__ __forceinline__ void coalesced_load(const GemmEpi& e, uint4* pre, const uint8_t* base, int64_t pitch_bytes,
                                               int64_t byte_off, int row0, int M) {
    const int l = (int)lane_id();
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int row = row0 + 8 * i + (l >> 2);
        uint4 v = make_uint4(0, 0, 0, 0);
        if (row < M) {
            const int64_t mr = kResRows ? map_res_row(e, row) : map_out_row(e, row);
            v = __ldg(reinterpret_cast<const uint4*>(base + mr * pitch_bytes + byte_off + (l & 3) * 16));
        }
        pre[i] = v;
    }
}

// operands of a full 32-column chunk, requested before the accumulator wait (warp-cooperative, coalesced)
__device__ __forceinline__ void staged_prefetch(const GemmEpi& e, uint4 (&pre)[8], int row0, int col0, int M) {
    if (e.residual) {
        const uint8_t* b = reinterpret_cast<const uint8_t*>(e.residual);
        coalesced_load<true>(e, pre, b, e.ldr * 4, (int64_t)col0 * 4, row0, M);
        coalesced_load<true>(e, pre + 4, b, e.ldr * 4, (int64_t)col0 * 4 + 64, row0, M);
    } else if (e.act == MICO_ACT_GELU_BWD || e.act == MICO_ACT_QUICK_GELU_BWD || e.act == MICO_ACT_MUL_AUX) {
        coalesced_load<false>(e, pre, reinterpret_cast<const uint8_t*>(e.aux_in), e.ld_aux_in * 2, (int64_t)col0 * 2, row0, M);
    }
}

// full 32-column chunk, all 32 lanes participate (rows >= M are predicated off at the global accesses)
__device__ __forceinline__ void epilogue_chunk32_staged(const GemmEpi& e, uint32_t st, const uint4 (&pre)[8],
                                                        const float* sbias, const uint32_t (&acc)[32], int row0, int col0,
                                                        int M) {
    const int row = row0 + (int)lane_id();
    float v[32];
    if (e.alpha != 1.0f) {
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(acc[i]) * e.alpha;
    } else {
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(acc[i]);
    }
    if (e.bias) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const float4 b = *reinterpret_cast<const float4*>(sbias + 4 * i);
            v[4 * i + 0] += b.x; v[4 * i + 1] += b.y; v[4 * i + 2] += b.z; v[4 * i + 3] += b.w;
        }
    }
    uint4 own[4];
    if (e.act == MICO_ACT_GELU_SAVE_GRAD || e.act == MICO_ACT_QUICK_GELU_SAVE_GRAD) {
        // out = act(v); aux_out = act'(v): the backward pass then needs one multiply per element (MICO_ACT_MUL_AUX).
        // Derivatives are packed eight at a time so that only v[32] + 4 words stay live (168-register budget).
        const bool erf_gelu = e.act == MICO_ACT_GELU_SAVE_GRAD;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            float gr[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const float x = v[8 * i + j];
                if (erf_gelu) {
                    float cdf, g;
                    gelu_parts(x, cdf, g);
                    gr[j] = fmaf(x * 0.3989422804014327f, g, cdf);
                    v[8 * i + j] = x * cdf;
                } else {
                    const float sg = rcp_fast(1.0f + ex2_raw(x * (-1.702f * 1.4426950408889634f)));
                    gr[j] = sg * fmaf(1.702f * x, 1.0f - sg, 1.0f);
                    v[8 * i + j] = x * sg;
                }
            }
            own[i] = make_uint4(pack_bf16x2(gr[0], gr[1]), pack_bf16x2(gr[2], gr[3]), pack_bf16x2(gr[4], gr[5]),
                                pack_bf16x2(gr[6], gr[7]));
        }
        if (e.aux_out)
            staged_store<false>(e, st, own, reinterpret_cast<uint8_t*>(e.aux_out), e.ld_aux_out * 2, (int64_t)col0 * 2, row0, M);
    } else if (e.act == MICO_ACT_GELU || e.act == MICO_ACT_QUICK_GELU) {
        if (e.aux_out) {
#pragma unroll
            for (int i = 0; i < 4; ++i)
                own[i] = make_uint4(pack_bf16x2(v[8 * i], v[8 * i + 1]), pack_bf16x2(v[8 * i + 2], v[8 * i + 3]),
                                    pack_bf16x2(v[8 * i + 4], v[8 * i + 5]), pack_bf16x2(v[8 * i + 6], v[8 * i + 7]));
            staged_store<false>(e, st, own, reinterpret_cast<uint8_t*>(e.aux_out), e.ld_aux_out * 2, (int64_t)col0 * 2, row0, M);
        }
        if (e.act == MICO_ACT_GELU) {
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = gelu_erf(v[i]);
        } else {
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = quick_gelu(v[i]);
        }
    } else if (e.act == MICO_ACT_MUL_AUX) {
        if (e.residual) {
            uint4 tmp[4];
            coalesced_load<false>(e, tmp, reinterpret_cast<const uint8_t*>(e.aux_in), e.ld_aux_in * 2, (int64_t)col0 * 2, row0, M);
            staged_to_own(st, tmp, own);
        } else {
            staged_to_own(st, pre, own);
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const uint32_t w[4] = {own[i].x, own[i].y, own[i].z, own[i].w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                v[8 * i + 2 * j] *= bf16_lo(w[j]);
                v[8 * i + 2 * j + 1] *= bf16_hi(w[j]);
            }
        }
    } else if (e.act == MICO_ACT_GELU_BWD || e.act == MICO_ACT_QUICK_GELU_BWD) {
        if (e.residual) {   // rare combination: the pre-activation was not prefetched
            uint4 tmp[4];
            coalesced_load<false>(e, tmp, reinterpret_cast<const uint8_t*>(e.aux_in), e.ld_aux_in * 2, (int64_t)col0 * 2, row0, M);
            staged_to_own(st, tmp, own);
        } else {
            staged_to_own(st, pre, own);
        }
        const bool erf_gelu = e.act == MICO_ACT_GELU_BWD;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const uint32_t w[4] = {own[i].x, own[i].y, own[i].z, own[i].w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float lo = bf16_lo(w[j]), hi = bf16_hi(w[j]);
                v[8 * i + 2 * j] *= erf_gelu ? gelu_erf_grad(lo) : quick_gelu_grad(lo);
                v[8 * i + 2 * j + 1] *= erf_gelu ? gelu_erf_grad(hi) : quick_gelu_grad(hi);
            }
        }
    }
    if (e.row_scale) {
        const float rs = row < M ? __ldg(e.row_scale + row / e.rows_per_group) : 0.f;
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] *= rs;
    }
    if (e.residual) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            staged_to_own(st, pre + 4 * h, own);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                v[16 * h + 4 * i + 0] += __uint_as_float(own[i].x); v[16 * h + 4 * i + 1] += __uint_as_float(own[i].y);
                v[16 * h + 4 * i + 2] += __uint_as_float(own[i].z); v[16 * h + 4 * i + 3] += __uint_as_float(own[i].w);
            }
        }
    }
    if (e.out_fp32) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
#pragma unroll
            for (int i = 0; i < 4; ++i)
                own[i] = make_uint4(__float_as_uint(v[16 * h + 4 * i]), __float_as_uint(v[16 * h + 4 * i + 1]),
                                    __float_as_uint(v[16 * h + 4 * i + 2]), __float_as_uint(v[16 * h + 4 * i + 3]));
            staged_store<false>(e, st, own, reinterpret_cast<uint8_t*>(e.out), e.ldo * 4, (int64_t)col0 * 4 + 64 * h, row0, M);
        }
    } else {
#pragma unroll
        for (int i = 0; i < 4; ++i)
            own[i] = make_uint4(pack_bf16x2(v[8 * i], v[8 * i + 1]), pack_bf16x2(v[8 * i + 2], v[8 * i + 3]),
                                pack_bf16x2(v[8 * i + 4], v[8 * i + 5]), pack_bf16x2(v[8 * i + 6], v[8 * i + 7]));
        staged_store<false>(e, st, own, reinterpret_cast<uint8_t*>(e.out), e.ldo * 2, (int64_t)col0 * 2, row0, M);
    }
}

// ------------------------------------------------------------------------------------------------------------
// Specialised epilogues.  The generic epilogue above decides everything at run time; inlined into the kernel it is
// ~20k SASS instructions, spends a third of its issue slots on branches / predicates / address arithmetic and stalls
// on instruction fetch (ncu round 1, fc1 forward: 38 warp instructions per output element, `no_inst` on every path).
// The hot launches of the towers fall into five shapes, each compiled as its own kernel (template parameter EPI):
//   EPI_BF16       out(bf16) = acc [+ bias]                                  qkv forward, every dgrad
//   EPI_GELU_SAVE  out(bf16) = act(acc + bias), aux_out(bf16) = act'(..)     fc1 forward
//   EPI_MUL_AUX    out(bf16) = acc * aux_in                                  fc2 dgrad
//   EPI_RES32      out(f32)  = rs[row] * (acc [+ bias]) + residual           proj / fc2 forward
//   EPI_F32        out(f32)  = acc                                           every wgrad
// Preconditions (checked by the dispatcher; anything else runs EPI_GENERIC): 16-byte aligned operands, N % 32 == 0,
// alpha == 1, no accumulate, no row remap.
enum { EPI_GENERIC = 0, EPI_BF16 = 1, EPI_GELU_SAVE = 2, EPI_MUL_AUX = 3, EPI_RES32 = 4, EPI_F32 = 5 };

template <int EPI>
__device__ __forceinline__ void fast_prefetch(const GemmEpi& e, uint4 (&pre)[8], int row0, int col0, int M) {
    if constexpr (EPI == EPI_RES32) {
        const uint8_t* b = reinterpret_cast<const uint8_t*>(e.residual);
        coalesced_load<false>(e, pre, b, e.ldr * 4, (int64_t)col0 * 4, row0, M);
        coalesced_load<false>(e, pre + 4, b, e.ldr * 4, (int64_t)col0 * 4 + 64, row0, M);
    } else if constexpr (EPI == EPI_MUL_AUX) {
        coalesced_load<false>(e, pre, reinterpret_cast<const uint8_t*>(e.aux_in), e.ld_aux_in * 2, (int64_t)col0 * 2, row0, M);
    }
}

// L2 prefetch of the NEXT chunk's epilogue operand (same lane -> address mapping as coalesced_load, no registers held): the
// register prefetch above is issued only one chunk ahead of its use, which leaves an HBM round trip exposed per chunk; with the
// line already in L2 the load costs an L2 hit (proj forward, K = 1408, is bound by exactly this latency).
template <int EPI>
__device__ __forceinline__ void fast_prefetch_l2(const GemmEpi& e, int row0, int col0, int M) {
    const int l = (int)lane_id();
    if constexpr (EPI == EPI_RES32) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int row = row0 + 8 * i + (l >> 2);
            if (row < M) {
                const float* p = e.residual + (int64_t)row * e.ldr + col0 + (l & 3) * 8;      // 4 lanes x 32 B = one 128-byte line
                asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
            }
        }
    } else if constexpr (EPI == EPI_MUL_AUX) {
        const int row = row0 + l;
        if (row < M) asm volatile("prefetch.global.L2 [%0];" ::"l"(e.aux_in + (int64_t)row * e.ld_aux_in + col0));   // 64 B of one line
    }
}

__device__ __forceinline__ void pack32_bf16(const float (&v)[32], uint4 (&own)[4]) {
#pragma unroll
    for (int i = 0; i < 4; ++i)
        own[i] = make_uint4(pack_bf16x2(v[8 * i], v[8 * i + 1]), pack_bf16x2(v[8 * i + 2], v[8 * i + 3]),
                            pack_bf16x2(v[8 * i + 4], v[8 * i + 5]), pack_bf16x2(v[8 * i + 6], v[8 * i + 7]));
}

// Specialised kernels write through TMA: the warp fills one of its two [32 rows x 64 B] tiles (own row, 64B swizzle ==
// stage_at) and lane 0 issues one bulk tensor store for the whole tile; rows >= M are clipped by the tensor map.
// The tiles alternate, so a tile is rewritten only after all but the newest store have finished reading (ncu round 1:
// the LDS -> STG halves of staged_store held a third of the epilogue's stall samples).
__device__ __forceinline__ uint32_t tile_acquire(uint32_t st, uint32_t& sidx) {
    if (lane_id() == 0) tma_store_wait_read<1>();
    __syncwarp();
    const uint32_t tile = st + (sidx & 1u) * 2048u;
    ++sidx;
    return tile;
}
__device__ __forceinline__ void tile_store(const CUtensorMap* tm, uint32_t tile, const uint4 (&own)[4], int c0, int row0,
                                           bool reduce_add = false) {
    const int l = (int)lane_id();
#pragma unroll
    for (int g = 0; g < 4; ++g) sts128(stage_at(tile, l, g), own[g]);
    fence_proxy_async_smem();
    __syncwarp();
    if (l == 0) {
        if (reduce_add) tma_reduce_add_2d(tm, tile, c0, row0);
        else tma_store_2d(tm, tile, c0, row0);
        tma_store_commit();
    }
}

// one full 32-column chunk; rs = DropPath row scale of this thread's row (EPI_RES32 only)
template <int EPI>
__device__ __forceinline__ void fast_chunk32(const GemmEpi& e, const CUtensorMap* tmOut, const CUtensorMap* tmAux, uint32_t st,
                                             uint32_t& sidx, const uint4 (&pre)[8], const float* sbias,
                                             const uint32_t (&acc)[32], int row0, int col0, int M, float rs) {
    float v[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(acc[i]);
    if constexpr (EPI == EPI_BF16 || EPI == EPI_GELU_SAVE || EPI == EPI_RES32) {
        if (e.bias) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const float4 b = *reinterpret_cast<const float4*>(sbias + 4 * i);
                v[4 * i + 0] += b.x; v[4 * i + 1] += b.y; v[4 * i + 2] += b.z; v[4 * i + 3] += b.w;
            }
        }
    }
    uint4 own[4];
    if constexpr (EPI == EPI_GELU_SAVE) {
        if (!e.aux_out) {
            // activation only (a checkpointed block's first forward pass keeps no derivative): same value as below
            if (e.act == MICO_ACT_GELU_SAVE_GRAD) {
#pragma unroll
                for (int i = 0; i < 32; ++i) {
                    float cdf, g;
                    gelu_parts(v[i], cdf, g);
                    v[i] *= cdf;
                }
            } else {
#pragma unroll
                for (int i = 0; i < 32; ++i)
                    v[i] *= rcp_fast(1.0f + ex2_raw(v[i] * (-1.702f * 1.4426950408889634f)));
            }
            pack32_bf16(v, own);
            tile_store(tmOut, tile_acquire(st, sidx), own, col0, row0);
            return;
        }
        if (e.act == MICO_ACT_GELU_SAVE_GRAD) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                float gr[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const float x = v[8 * i + j];
                    float cdf, g;
                    gelu_parts(x, cdf, g);
                    gr[j] = fmaf(x * 0.3989422804014327f, g, cdf);
                    v[8 * i + j] = x * cdf;
                }
                own[i] = make_uint4(pack_bf16x2(gr[0], gr[1]), pack_bf16x2(gr[2], gr[3]), pack_bf16x2(gr[4], gr[5]),
                                    pack_bf16x2(gr[6], gr[7]));
            }
        } else {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                float gr[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const float x = v[8 * i + j];
                    const float sg = rcp_fast(1.0f + ex2_raw(x * (-1.702f * 1.4426950408889634f)));
                    gr[j] = sg * fmaf(1.702f * x, 1.0f - sg, 1.0f);
                    v[8 * i + j] = x * sg;
                }
                own[i] = make_uint4(pack_bf16x2(gr[0], gr[1]), pack_bf16x2(gr[2], gr[3]), pack_bf16x2(gr[4], gr[5]),
                                    pack_bf16x2(gr[6], gr[7]));
            }
        }
        tile_store(tmAux, tile_acquire(st, sidx), own, col0, row0);
    } else if constexpr (EPI == EPI_MUL_AUX) {
        const uint32_t tile = tile_acquire(st, sidx);      // scratch for the transpose now, output tile below
        --sidx;
        staged_to_own(tile, pre, own);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const uint32_t w[4] = {own[i].x, own[i].y, own[i].z, own[i].w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                v[8 * i + 2 * j] *= bf16_lo(w[j]);
                v[8 * i + 2 * j + 1] *= bf16_hi(w[j]);
            }
        }
    } else if constexpr (EPI == EPI_RES32) {
        if (e.row_scale) {
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] *= rs;
        }
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const uint32_t tile = tile_acquire(st, sidx);   // transpose scratch, then this half's output tile
            staged_to_own(tile, pre + 4 * h, own);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                own[i].x = __float_as_uint(v[16 * h + 4 * i + 0] + __uint_as_float(own[i].x));
                own[i].y = __float_as_uint(v[16 * h + 4 * i + 1] + __uint_as_float(own[i].y));
                own[i].z = __float_as_uint(v[16 * h + 4 * i + 2] + __uint_as_float(own[i].z));
                own[i].w = __float_as_uint(v[16 * h + 4 * i + 3] + __uint_as_float(own[i].w));
            }
            tile_store(tmOut, tile, own, col0 + 16 * h, row0);
        }
        return;
    }
    if constexpr (EPI == EPI_F32) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
#pragma unroll
            for (int i = 0; i < 4; ++i)
                own[i] = make_uint4(__float_as_uint(v[16 * h + 4 * i]), __float_as_uint(v[16 * h + 4 * i + 1]),
                                    __float_as_uint(v[16 * h + 4 * i + 2]), __float_as_uint(v[16 * h + 4 * i + 3]));
            tile_store(tmOut, tile_acquire(st, sidx), own, col0 + 16 * h, row0, e.ksplit > 1);
        }
    } else {
        pack32_bf16(v, own);
        tile_store(tmOut, tile_acquire(st, sidx), own, col0, row0);
    }
}

// the 16-column remainder chunk of a 176-wide tile: per-thread vector accesses (1/11 of the columns)
template <int EPI>
__device__ __forceinline__ void fast_chunk16(const GemmEpi& e, const float* sbias, const uint32_t (&acc)[16], int row,
                                             int col0, int M, float rs) {
    if (row >= M) return;
    float v[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(acc[i]);
    if constexpr (EPI == EPI_BF16 || EPI == EPI_GELU_SAVE || EPI == EPI_RES32) {
        if (e.bias) {
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] += sbias[i];
        }
    }
    if constexpr (EPI == EPI_GELU_SAVE) {
        float gr[16];
        const bool erf_gelu = e.act == MICO_ACT_GELU_SAVE_GRAD;
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            gr[i] = erf_gelu ? gelu_erf_grad(v[i]) : quick_gelu_grad(v[i]);
            v[i] = erf_gelu ? gelu_erf(v[i]) : quick_gelu(v[i]);
        }
        uint4* a4 = reinterpret_cast<uint4*>(e.aux_out + (int64_t)row * e.ld_aux_out + col0);
#pragma unroll
        for (int i = 0; i < 2; ++i)
            if (e.aux_out) a4[i] = make_uint4(pack_bf16x2(gr[8 * i], gr[8 * i + 1]), pack_bf16x2(gr[8 * i + 2], gr[8 * i + 3]),
                               pack_bf16x2(gr[8 * i + 4], gr[8 * i + 5]), pack_bf16x2(gr[8 * i + 6], gr[8 * i + 7]));
    } else if constexpr (EPI == EPI_MUL_AUX) {
        const uint4* u4 = reinterpret_cast<const uint4*>(e.aux_in + (int64_t)row * e.ld_aux_in + col0);
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            const uint4 u = __ldg(u4 + i);
            const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                v[8 * i + 2 * j] *= bf16_lo(w[j]);
                v[8 * i + 2 * j + 1] *= bf16_hi(w[j]);
            }
        }
    } else if constexpr (EPI == EPI_RES32) {
        const float4* r4 = reinterpret_cast<const float4*>(e.residual + (int64_t)row * e.ldr + col0);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const float4 r = __ldg(r4 + i);
            v[4 * i + 0] = fmaf(v[4 * i + 0], rs, r.x); v[4 * i + 1] = fmaf(v[4 * i + 1], rs, r.y);
            v[4 * i + 2] = fmaf(v[4 * i + 2], rs, r.z); v[4 * i + 3] = fmaf(v[4 * i + 3], rs, r.w);
        }
    }
    if constexpr (EPI == EPI_RES32 || EPI == EPI_F32) {
        float4* o4 = reinterpret_cast<float4*>(reinterpret_cast<float*>(e.out) + (int64_t)row * e.ldo + col0);
#pragma unroll
        for (int i = 0; i < 4; ++i) o4[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
    } else {
        uint4* o4 = reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(e.out) + (int64_t)row * e.ldo + col0);
#pragma unroll
        for (int i = 0; i < 2; ++i)
            o4[i] = make_uint4(pack_bf16x2(v[8 * i], v[8 * i + 1]), pack_bf16x2(v[8 * i + 2], v[8 * i + 3]),
                               pack_bf16x2(v[8 * i + 4], v[8 * i + 5]), pack_bf16x2(v[8 * i + 6], v[8 * i + 7]));
    }
}

// CL = 2: two CTAs of a cluster (one SM pair) work on two vertically adjacent M tiles of the same N tile.  Each CTA
// loads only HALF of the B tile and multicasts it into both CTAs' shared memory (cp.async.bulk.tensor ...
// .multicast::cluster); smem slots are recycled when BOTH CTAs' MMAs have consumed them (multicast tcgen05.commit).
// The mainloop of this GEMM is bound by L2 -> SM operand traffic (15 TB/s for 128x256 tiles at 1.3 PFLOP/s, which is
// why any extra epilogue traffic used to ADD to the run time instead of overlapping); sharing B across the pair removes
// a third of it.
template <int BN, bool A_MN, bool B_MN, int STAGES, int CL, int EPI>
__global__ void __launch_bounds__(kGemmThreads, 1)   // 10 warps -> 3 on one SM sub-partition: 168 registers max
gemm_bf16_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                 const __grid_constant__ CUtensorMap tmOut, const __grid_constant__ CUtensorMap tmAux, int M, int N,
                 int K, GemmEpi epi) {
    using Cfg = GemmCfg<BN, STAGES, EPI, B_MN, CL == 3>;
    static_assert((2 * STAGES + 4) * 8 + 8 <= Cfg::BAR_BYTES, "barrier block");
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    if constexpr (EPI != EPI_GENERIC) {     // no alignment slack was allocated (see GemmCfg)
        if (smem != smem_raw) __trap();
    }
    uint8_t* staging_all = smem + STAGES * Cfg::STAGE_BYTES;
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(staging_all + Cfg::STAGING_BYTES);
    uint64_t* empty_bar = full_bar + STAGES;
    uint64_t* tfull_bar = empty_bar + STAGES;
    uint64_t* tempty_bar = tfull_bar + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);
    float* sbias_all = reinterpret_cast<float*>(staging_all + Cfg::STAGING_BYTES + Cfg::BAR_BYTES);

    const int warp = threadIdx.x >> 5;
    const int num_m = (M + BM - 1) / BM;
    const int num_n = (N + BN - 1) / BN;
    const int num_kb = (K + BK - 1) / BK;
    // persistent schedule over work units = (group of CL adjacent m tiles, n tile); both CTAs of a cluster walk the
    // same unit list in lock step
    // CL = 3 selects the cta_group::2 variant of the pair (P2): the leader CTA issues ONE M = 256 MMA for both CTAs, each CTA
    // stages its own 128 A rows and its own half of the B tile (no multicast, a third less shared-memory operand traffic,
    // six 32 KB stages instead of four 48 KB ones); CLN = CTAs per cluster.
    constexpr bool P2 = CL == 3;
    constexpr int CLN = CL == 1 ? 1 : 2;
    const uint32_t cta_rank = CLN > 1 ? cluster_ctarank() : 0u;
    const int num_tiles = ((num_m + CLN - 1) / CLN) * num_n;
    // Work units in m-group-major order: the CTAs running concurrently share A panels through L2.  (Tried: all full-width
    // N tiles first and the narrower last-N tiles at the end, to fill the last wave -- the final sweep re-reads every A
    // panel from HBM and was 5-8 % slower on the N = 1408 GEMMs.)
    auto unit_mg = [&](int u) { return u / num_n; };
    auto unit_nt = [&](int u) { return u % num_n; };
    const int unit0 = blockIdx.x / CLN, unit_stride = gridDim.x / CLN;
    constexpr uint16_t kMask = (1u << CLN) - 1;
    // split-K (epi.ksplit > 1, EPI_F32 only): work unit u = (tile u % num_tiles, K part u / num_tiles) -- the parts of one
    // tile run concurrently on different CTAs and add their fp32 partial tiles into the zeroed output (TMA reduce).  A weight
    // gradient with a 200k-row K loop and 36 output tiles otherwise leaves half of the SM pairs idle.
    const int ksplit = EPI == EPI_F32 ? epi.ksplit : 1;
    const int kb_per = (num_kb + ksplit - 1) / ksplit;
    const int num_units = num_tiles * ksplit;
    // it-th work unit of this CTA's slot (-1: none left).  With a schedule table the units of one round (`slots` consecutive
    // units) are dealt so that every slot accumulates the same MMA time although the last N tile is narrower.
    auto unit_at = [&](int it) -> int {
        if (epi.sched) return it < epi.sched_rounds ? __ldg(epi.sched + (size_t)unit0 * epi.sched_rounds + it) : -1;
        const int u = unit0 + it * unit_stride;
        return u < num_units ? u : -1;
    };

    if (warp == 0) {
        if (elect_one()) {
            tma_prefetch_desc(&tmA);
            tma_prefetch_desc(&tmB);
            if constexpr (EPI != EPI_GENERIC) tma_prefetch_desc(&tmOut);
            if constexpr (EPI == EPI_GELU_SAVE) tma_prefetch_desc(&tmAux);
        }
    } else if (warp == 1) {
        if (elect_one()) {
            for (int s = 0; s < STAGES; ++s) {
                mbar_init(&full_bar[s], 1);
                mbar_init(&empty_bar[s], P2 ? 1 : CLN);
            }
            for (int a = 0; a < 2; ++a) {
                mbar_init(&tfull_bar[a], 1);
                mbar_init(&tempty_bar[a], P2 ? 16 : 8);      // P2: the leader also waits for the peer's eight epilogue warps
            }
            fence_mbar_init();
        }
        __syncwarp();
        if constexpr (P2) tmem_alloc_pair<512>(tmem_slot); else tmem_alloc<512>(tmem_slot);
    }
    tc_fence_before();
    __syncthreads();
    if constexpr (CLN > 1) cluster_sync_all();    // peer barriers are initialised before any multicast reaches them
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ------------------------------------------------------------- TMA producer
        if (elect_one()) {
            int stage = 0;
            uint32_t phase = 0;
            int unit = unit_at(0);
            for (int it = 0; unit >= 0; ++it) {
                const int unit_next = unit_at(it + 1);      // the table read of the next unit overlaps this one's K loop
                const int tile = unit % num_tiles, kb0 = (unit / num_tiles) * kb_per;
                const int kb1 = kb0 + kb_per < num_kb ? kb0 + kb_per : num_kb;
                const int m0 = (unit_mg(tile) * CLN + (int)cta_rank) * BM;
                const int n0 = unit_nt(tile) * BN;
                for (int kb = kb0; kb < kb1; ++kb) {
                    mbar_wait(&empty_bar[stage], phase ^ 1);
                    uint8_t* sa = smem + stage * Cfg::STAGE_BYTES;
                    uint8_t* sb = sa + Cfg::A_BYTES;
                    if constexpr (P2) {
                        // both CTAs' loads are counted on the LEADER's full barrier (it issues the MMA for the pair)
                        const uint32_t fb = mapa_u32(smem_u32(&full_bar[stage]), 0);
                        if (cta_rank == 0) mbar_arrive_expect_tx(&full_bar[stage], 2 * Cfg::STAGE_BYTES);
                        const int n_left = N - n0;
                        // asum_out: the last (narrow) N tile carries 32 more MMA columns; the peer CTA's second 64-column box
                        // (tile columns n_eff/2 + 64 ..., all past N) is loaded from the ones tensor instead of B
                        const bool ones_tile = B_MN && epi.asum_out != nullptr && n_left < BN;
                        const int n_eff = n_left >= BN ? BN : ((n_left + 31) & ~31) + (ones_tile ? 32 : 0);
                        const int nb = n0 + (int)cta_rank * (n_eff / 2);        // this CTA's half of the (possibly narrow) tile
                        if constexpr (!A_MN) {
                            tma_load_2d_pair(sa, &tmA, fb, kb * BK, m0);
                        } else {
#pragma unroll
                            for (int j = 0; j < BM / 64; ++j) tma_load_2d_pair(sa + j * 8192, &tmA, fb, m0 + 64 * j, kb * BK);
                        }
                        if constexpr (!B_MN) {
                            tma_load_2d_pair(sb, &tmB, fb, kb * BK, nb);
                        } else {
#pragma unroll
                            for (int j = 0; j < Cfg::B_CHUNKS / 2; ++j) {
                                if (ones_tile && cta_rank == 1 && j == 1) tma_load_2d_pair(sb + j * 8192, &tmAux, fb, 0, kb * BK);
                                else tma_load_2d_pair(sb + j * 8192, &tmB, fb, nb + 64 * j, kb * BK);
                            }
                        }
                        if (++stage == STAGES) { stage = 0; phase ^= 1; }
                        continue;
                    }
                    mbar_arrive_expect_tx(&full_bar[stage], Cfg::STAGE_BYTES);
                    if constexpr (!A_MN) {
                        tma_load_2d(sa, &tmA, &full_bar[stage], kb * BK, m0);
                    } else {
#pragma unroll
                        for (int j = 0; j < BM / 64; ++j)
                            tma_load_2d(sa + j * 8192, &tmA, &full_bar[stage], m0 + 64 * j, kb * BK);
                    }
                    if constexpr (CLN == 1) {
                        if constexpr (!B_MN) {
                            tma_load_2d(sb, &tmB, &full_bar[stage], kb * BK, n0);
                        } else {
#pragma unroll
                            for (int j = 0; j < Cfg::B_CHUNKS; ++j)
                                tma_load_2d(sb + j * 8192, &tmB, &full_bar[stage], n0 + 64 * j, kb * BK);
                        }
                    } else {     // this CTA's half of B, multicast to both CTAs of the pair
                        if constexpr (!B_MN) {
                            constexpr int HALF = BN / CLN;
                            tma_load_2d_mc(sb + cta_rank * (HALF * 128), &tmB, &full_bar[stage], kb * BK,
                                           n0 + (int)cta_rank * HALF, kMask);
                        } else {
#pragma unroll
                            for (int j = 0; j < Cfg::B_CHUNKS; ++j) {     // chunks are dealt out to the CTAs of the pair
                                if ((j * CLN) / Cfg::B_CHUNKS == (int)cta_rank)
                                    tma_load_2d_mc(sb + j * 8192, &tmB, &full_bar[stage], n0 + 64 * j, kb * BK, kMask);
                            }
                        }
                    }
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
                unit = unit_next;
            }
        }
    } else if (warp == 1) {
        // ------------------------------------------------------------- MMA issuer (P2: the leader CTA only)
        if ((!P2 || cta_rank == 0) && elect_one()) {
            int stage = 0;
            uint32_t phase = 0;
            int unit = unit_at(0);
            for (int it = 0; unit >= 0; ++it) {
                const int unit_next = unit_at(it + 1);
                const int tile = unit % num_tiles, kb0 = (unit / num_tiles) * kb_per;
                const int kb1 = kb0 + kb_per < num_kb ? kb0 + kb_per : num_kb;
                const int acc = it & 1;
                const uint32_t acc_phase = (it >> 1) & 1;
                // the last N tile issues a narrower MMA: no tensor-pipe time is spent on the columns past N (the smem rows
                // behind them are zero-filled by TMA), which lets N = 1408 run on 256-wide tiles as 5 full tiles + 1 half
                const int n_left = N - unit_nt(tile) * BN;
                const bool ones_tile = P2 && B_MN && epi.asum_out != nullptr && n_left < BN;
                const uint32_t idesc = P2 ? umma_idesc_bf16(n_left >= BN ? BN : ((n_left + 31) & ~31) + (ones_tile ? 32 : 0), A_MN, B_MN, 256)
                                          : umma_idesc_bf16(n_left >= BN ? BN : ((n_left + 15) & ~15), A_MN, B_MN);
                mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + acc * kAccStride;
                for (int kb = kb0; kb < kb1; ++kb) {
                    mbar_wait(&full_bar[stage], phase);
                    tc_fence_after();
                    const uint32_t sa = smem_u32(smem + stage * Cfg::STAGE_BYTES);
                    const uint32_t sb = sa + Cfg::A_BYTES;
#pragma unroll
                    for (int k = 0; k < BK / 16; ++k) {
                        const uint64_t adesc = A_MN ? umma_smem_desc_sw128(sa + k * 2048, 8192, 1024)
                                                    : umma_smem_desc_sw128(sa + k * 32, 16, 1024);
                        const uint64_t bdesc = B_MN ? umma_smem_desc_sw128(sb + k * 2048, 8192, 1024)
                                                    : umma_smem_desc_sw128(sb + k * 32, 16, 1024);
                        if constexpr (P2) umma_bf16_ss_pair(d_tmem, adesc, bdesc, idesc, ((kb - kb0) | k) != 0);
                        else umma_bf16_ss(d_tmem, adesc, bdesc, idesc, ((kb - kb0) | k) != 0);
                    }
                    if constexpr (CLN == 1) umma_commit(&empty_bar[stage]);
                    else if constexpr (P2) umma_commit_pair(&empty_bar[stage], kMask);
                    else umma_commit_mc(&empty_bar[stage], kMask);      // frees the slot in both CTAs
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
                if constexpr (P2) umma_commit_pair(&tfull_bar[acc], kMask);     // both CTAs' epilogues read their half
                else umma_commit(&tfull_bar[acc]);
                unit = unit_next;
            }
        }
    } else {
        // ------------------------------------------------------------- epilogue (warps 2..9)
        const int q = warp & 3;            // TMEM lane quarter this warp may access
        const int half = (warp - 2) >> 2;  // which of the two warps of that quarter: owns chunks c with (c & 1) == half
        uint32_t sidx = 0;                 // TMA store tiles of this warp alternate (specialised epilogues)
        int unit = unit_at(0);
        for (int it = 0; unit >= 0; ++it) {
            const int unit_cur = unit;
            unit = unit_at(it + 1);
            const int tile = unit_cur % num_tiles;
            const int acc = it & 1;
            const uint32_t acc_phase = (it >> 1) & 1;
            const int m0 = (unit_mg(tile) * CLN + (int)cta_rank) * BM;
            const int n0 = unit_nt(tile) * BN;
            // bias slice of this tile -> smem (one coalesced load per tile instead of 8 x 16 B per thread and chunk).
            // Stage `acc` of the buffer was last read two tiles ago; every epilogue thread has passed the named
            // barrier of the previous tile since then.
            float* sbias = sbias_all + acc * 256;
            if (epi.bias) {
                const int t = threadIdx.x - 64;
                if (t < BN) sbias[t] = (n0 + t < N) ? __ldg(epi.bias + n0 + t) : 0.f;
            }
            asm volatile("bar.sync 1, 256;" ::: "memory");
            const int row = m0 + q * 32 + (int)lane_id();
            const uint32_t t0 = tmem_base + acc * kAccStride + ((uint32_t)(q * 32) << 16);
            bool waited = false;
            const uint32_t st = smem_u32(staging_all + (warp - 2) * Cfg::STAGE_TILE);
            const int row0 = m0 + q * 32;
            // the staged (coalesced) path needs 16-byte-aligned pitches and no read-modify-write of the output
            const bool staged_ok = epi.vec_ok && !epi.accumulate;
            if constexpr (EPI != EPI_GENERIC) {
                float rs = 1.0f;
                if constexpr (EPI == EPI_RES32) {
                    if (epi.row_scale) rs = row < M ? __ldg(epi.row_scale + row / epi.rows_per_group) : 0.f;
                }
                if constexpr (EPI == EPI_RES32 || EPI == EPI_MUL_AUX) {
                    // pull this warp's share of the tile's epilogue operand into L2 while the accumulator is still being produced
#pragma unroll 1
                    for (int c = half; c < BN / 32; c += 2)
                        if (n0 + c * 32 < N) fast_prefetch_l2<EPI>(epi, row0, n0 + c * 32, M);
                }
#pragma unroll 1
                for (int c = half; c < BN / 32; c += 2) {
                    if (n0 + c * 32 >= N) break;   // warp-uniform; N % 32 == 0: a chunk is full or absent
                    uint4 pre[8];
                    fast_prefetch<EPI>(epi, pre, row0, n0 + c * 32, M);
                    if (!waited) { mbar_wait(&tfull_bar[acc], acc_phase); tc_fence_after(); waited = true; }
                    uint32_t v[32];
                    tmem_ld_x32(t0 + c * 32, v);
                    tmem_ld_wait();
                    fast_chunk32<EPI>(epi, &tmOut, &tmAux, st, sidx, pre, sbias + c * 32, v, row0, n0 + c * 32, M, rs);
                }
                if (!waited) { mbar_wait(&tfull_bar[acc], acc_phase); tc_fence_after(); }
                if constexpr (EPI == EPI_F32 && P2 && B_MN) {
                    // asum_out: the ones columns of the last N tile start at tile column n_eff/2 + 64 (see the producer);
                    // every row adds its sum over this unit's K range
                    if (epi.asum_out != nullptr && N - n0 < BN && half == 0) {
                        const int n_eff = ((N - n0 + 31) & ~31) + 32;
                        uint32_t v[16];
                        tmem_ld_x16(t0 + n_eff / 2 + 64, v);
                        tmem_ld_wait();
                        if (row < M) atomicAdd(epi.asum_out + row, __uint_as_float(v[0]));
                    }
                }
                if constexpr (BN % 32 != 0) {
                    constexpr int c0 = (BN / 32) * 32;
                    if (half == ((BN / 32) & 1) && n0 + c0 < N) {
                        uint32_t v[16];
                        tmem_ld_x16(t0 + c0, v);
                        tmem_ld_wait();
                        fast_chunk16<EPI>(epi, sbias + c0, v, row, n0 + c0, M, rs);
                    }
                }
                tc_fence_before();
                __syncwarp();
                if (lane_id() == 0) { if constexpr (P2) mbar_arrive_cluster(mapa_u32(smem_u32(&tempty_bar[acc]), 0)); else mbar_arrive(&tempty_bar[acc]); }
            } else {
#pragma unroll 1
            for (int c = half; c < BN / 32; c += 2) {
                if (n0 + c * 32 >= N) break;   // warp-uniform
                if (staged_ok && n0 + c * 32 + 32 <= N) {
                    uint4 pre[8];
                    staged_prefetch(epi, pre, row0, n0 + c * 32, M);
                    if (!waited) { mbar_wait(&tfull_bar[acc], acc_phase); tc_fence_after(); waited = true; }
                    uint32_t v[32];
                    tmem_ld_x32(t0 + c * 32, v);
                    tmem_ld_wait();
                    epilogue_chunk32_staged(epi, st, pre, sbias + c * 32, v, row0, n0 + c * 32, M);
                    continue;
                }
                EpiOperands ops;
                epilogue_prefetch<32>(epi, ops, row, n0 + c * 32, M, N);
                if (!waited) { mbar_wait(&tfull_bar[acc], acc_phase); tc_fence_after(); waited = true; }
                uint32_t v[32];
                tmem_ld_x32(t0 + c * 32, v);
                tmem_ld_wait();
                epilogue_chunk<32>(epi, ops, sbias + c * 32, v, row, n0 + c * 32, M, N);
            }
            if (!waited) { mbar_wait(&tfull_bar[acc], acc_phase); tc_fence_after(); }
            if constexpr (BN % 32 != 0) {
                constexpr int c0 = (BN / 32) * 32;
                if (half == ((BN / 32) & 1) && n0 + c0 < N) {
                    EpiOperands ops;
                    epilogue_prefetch<16>(epi, ops, row, n0 + c0, M, N);
                    uint32_t v[16];
                    tmem_ld_x16(t0 + c0, v);
                    tmem_ld_wait();
                    epilogue_chunk<16>(epi, ops, sbias + c0, v, row, n0 + c0, M, N);
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane_id() == 0) { if constexpr (P2) mbar_arrive_cluster(mapa_u32(smem_u32(&tempty_bar[acc]), 0)); else mbar_arrive(&tempty_bar[acc]); }
            }   // EPI_GENERIC
        }
        if constexpr (EPI != EPI_GENERIC) {
            if (lane_id() == 0) tma_store_wait<0>();     // shared memory stays valid until every store has read it
        }
    }
    tc_fence_before();
    __syncthreads();
    if constexpr (CLN > 1) cluster_sync_all();    // no CTA leaves while its peer may still multicast into it
    if (warp == 1) {
        tc_fence_after();
        if constexpr (P2) tmem_dealloc_pair<512>(tmem_base); else tmem_dealloc<512>(tmem_base);
    }
}

// ---------------------------------------------------------------------------------------------------------------------------
// Work-unit plan of one launch: the split-K factor (weight gradients) and a balanced unit schedule.
//
// The persistent kernel used to deal units to its slots (CTAs / CTA pairs) with a fixed stride.  With N = 1408 on 256-wide tiles
// the sixth N tile issues a half-width MMA, and a stride of 74 slots over 6 N tiles leaves every slot on n tiles of ONE parity:
// odd slots only ever saw n = 1, 3, 5 -- 5/6 of the even slots' MMA time -- so the launch ran 62.5 full-tile times where 57.3
// were needed, and the faster half of the grid drifted ahead of the A panels the others were still reading (fc2 forward at
// 197 376 rows: 8.4 GB of DRAM reads for 3.5 GB of operands).  The plan keeps the round structure (round r = the r-th group of
// `slots` consecutive units: concurrent CTAs share A panels through L2, and the K parts of a split weight gradient stay in lock
// step) but deals each round's units longest-first to the least-loaded slots, so every slot's accumulated MMA time stays
// within half a tile of every other's.  Unit cost = MMA columns x K blocks (+ a fixed per-unit drain).  The split-K factor of a
// weight gradient is the one with the smallest planned makespan.  MICO_GEMM_SCHED=0 restores the strided schedule.
struct GemmPlan {
    int ks = 1;
    int rounds = 0;
    double makespan = 0;            // planned time of the slowest slot, in (MMA column x K block) units
    bool uniform = true;            // every unit costs the same: the strided default is already balanced
    std::vector<int> table;         // [slots x rounds], -1 padded
    const int* dev = nullptr;       // device copy, uploaded on first use
};

static constexpr int kUnitDrainKb = 4;      // per-unit fixed cost (pipeline fill + accumulator hand-over), in K blocks

static void plan_units(GemmPlan& p, int num_tiles, int num_n, int bn, int n_last, int num_kb, int ks, int slots) {
    const int kb_per = ceil_div(num_kb, ks);
    const int num_units = num_tiles * ks;
    auto cost = [&](int u) {
        const int tile = u % num_tiles, part = u / num_tiles;
        const int w = (tile % num_n == num_n - 1) ? n_last : bn;
        int kbc = num_kb - part * kb_per;
        if (kbc > kb_per) kbc = kb_per;
        if (kbc < 0) kbc = 0;
        return (double)w * (kbc + kUnitDrainKb);
    };
    p.ks = ks;
    p.rounds = ceil_div(num_units, slots);
    p.uniform = true;
    const double c0 = cost(0);
    for (int u = 1; u < num_units && p.uniform; ++u) p.uniform = cost(u) == c0;
    std::vector<double> load(slots, 0.0);
    std::vector<int> count(slots, 0), order(slots), units;
    p.table.assign((size_t)slots * p.rounds, -1);
    for (int r = 0; r < p.rounds; ++r) {
        const int u0 = r * slots, u1 = u0 + slots < num_units ? u0 + slots : num_units;
        units.resize(u1 - u0);
        for (int u = u0; u < u1; ++u) units[u - u0] = u;
        if (!p.uniform) {
            std::stable_sort(units.begin(), units.end(), [&](int a, int b) { return cost(a) > cost(b); });
            for (int s = 0; s < slots; ++s) order[s] = s;
            std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return load[a] < load[b]; });
        }
        for (int i = 0; i < (int)units.size(); ++i) {
            const int s = p.uniform ? i : order[i];
            p.table[(size_t)s * p.rounds + count[s]++] = units[i];
            load[s] += cost(units[i]);
        }
    }
    p.makespan = *std::max_element(load.begin(), load.end());
}

static bool sched_enabled() {
    static const bool on = [] { const char* e = getenv("MICO_GEMM_SCHED"); return !(e && e[0] == '0'); }();
    return on;
}

// cached per (shape, device); returned reference stays valid (node-based map)
static GemmPlan& get_plan(int num_tiles, int num_n, int bn, int n_last, int num_kb, int slots, bool allow_split) {
    static std::mutex mu;
    static std::map<std::array<int, 8>, GemmPlan> cache;
    int dev = 0;
    cudaGetDevice(&dev);
    const std::array<int, 8> key = {num_tiles, num_n, bn, n_last, num_kb, slots, allow_split ? 1 : 0, dev};
    std::lock_guard<std::mutex> lock(mu);
    auto it = cache.find(key);
    if (it != cache.end()) return it->second;
    static const int ks_max = [] { const char* e = getenv("MICO_GEMM_SPLITK_MAX"); return e ? atoi(e) : 4; }();
    GemmPlan best;
    plan_units(best, num_tiles, num_n, bn, n_last, num_kb, 1, slots);
    if (allow_split) {
        // keep >= 16 K blocks per part; a finer split has to win by 4 % (it zeroes the output first and adds partial tiles:
        // two parts add commutatively (bit-reproducible), more than two arrive in any order -- last-bit run-to-run
        // differences like cuBLAS split-K; MICO_GEMM_SPLITK_MAX=2 / =1 restricts)
        for (int cand = 2; cand <= ks_max; cand *= 2) {
            if (num_kb / cand < 16) break;
            GemmPlan p;
            plan_units(p, num_tiles, num_n, bn, n_last, num_kb, cand, slots);
            if (p.makespan < best.makespan * 0.96) best = std::move(p);
        }
    }
    if (!sched_enabled()) {          // strided schedule: its makespan is what the estimate should see
        const int ks = best.ks, num_units = num_tiles * ks, kb_per = ceil_div(num_kb, ks);
        std::vector<double> load(slots, 0.0);
        for (int u = 0; u < num_units; ++u) {
            const int tile = u % num_tiles, part = u / num_tiles;
            int kbc = num_kb - part * kb_per;
            kbc = kbc > kb_per ? kb_per : (kbc < 0 ? 0 : kbc);
            load[u % slots] += (double)((tile % num_n == num_n - 1) ? n_last : bn) * (kbc + kUnitDrainKb);
        }
        best.makespan = *std::max_element(load.begin(), load.end());
        best.uniform = true;
    }
    return cache.emplace(key, std::move(best)).first->second;
}

// device copy of a plan's schedule table (null when the strided default is as good)
static int plan_device_table(GemmPlan& p, const int** out) {
    *out = nullptr;
    if (p.uniform || p.table.empty()) return MICO_OK;
    static std::mutex mu;
    std::lock_guard<std::mutex> lock(mu);
    if (!p.dev) {
        int* d = nullptr;
        MICO_CHECK_CUDA(cudaMalloc(&d, p.table.size() * sizeof(int)));
        MICO_CHECK_CUDA(cudaMemcpy(d, p.table.data(), p.table.size() * sizeof(int), cudaMemcpyHostToDevice));
        p.dev = d;
    }
    *out = p.dev;
    return MICO_OK;
}

// MMA columns of the last N tile (the narrower MMA the issuer picks, see the kernel)
static int last_tile_cols(int N, int bn, bool p2) {
    const int n_left = N - (ceil_div(N, bn) - 1) * bn;
    const int g = p2 ? 32 : 16;
    return n_left >= bn ? bn : ((n_left + g - 1) / g) * g;
}

template <int BN, bool A_MN, bool B_MN, int STAGES, int CL, int EPI>
int launch_gemm(const MicoGemmArgs& g, const GemmEpi& epi_in, cudaStream_t stream) {
    using Cfg = GemmCfg<BN, STAGES, EPI, B_MN, CL == 3>;
    constexpr int CLN = CL == 1 ? 1 : 2;
    CUtensorMap tmA, tmB;
    int rc;
    if (!A_MN) {
        const uint64_t dims[2] = {(uint64_t)g.K, (uint64_t)g.M};
        const uint64_t strides[2] = {2, (uint64_t)g.lda * 2};
        const uint32_t box[2] = {BK, BM};
        rc = make_tmap_bf16(&tmA, g.a, 2, dims, strides, box);
    } else {
        const uint64_t dims[2] = {(uint64_t)g.M, (uint64_t)g.K};
        const uint64_t strides[2] = {2, (uint64_t)g.lda * 2};
        const uint32_t box[2] = {64, BK};
        rc = make_tmap_bf16(&tmA, g.a, 2, dims, strides, box);
    }
    if (rc) return rc;
    if (!B_MN) {
        const uint64_t dims[2] = {(uint64_t)g.K, (uint64_t)g.N};
        const uint64_t strides[2] = {2, (uint64_t)g.ldb * 2};
        const uint32_t box[2] = {BK, BN / CLN};      // pairs: each CTA loads half of the B tile (CL = 2: and multicasts it)
        rc = make_tmap_bf16(&tmB, g.b, 2, dims, strides, box);
    } else {
        const uint64_t dims[2] = {(uint64_t)g.N, (uint64_t)g.K};
        const uint64_t strides[2] = {2, (uint64_t)g.ldb * 2};
        const uint32_t box[2] = {64, BK};
        rc = make_tmap_bf16(&tmB, g.b, 2, dims, strides, box);
    }
    if (rc) return rc;

    CUtensorMap tmOut = tmA, tmAux = tmA;     // placeholders unless the specialised epilogue stores through them
    if constexpr (EPI != EPI_GENERIC) {
        const bool f32 = epi_in.out_fp32 != 0;
        if ((rc = make_tmap_tile64(&tmOut, epi_in.out, f32, (uint64_t)g.N, (uint64_t)g.M, (uint64_t)epi_in.ldo * (f32 ? 4 : 2)))) return rc;
    }
    if constexpr (EPI == EPI_GELU_SAVE) {
        if (epi_in.aux_out)
        if ((rc = make_tmap_tile64(&tmAux, epi_in.aux_out, false, (uint64_t)g.N, (uint64_t)g.M, (uint64_t)epi_in.ld_aux_out * 2))) return rc;
    }
    int asum_cols = 0;      // extra MMA columns of the last N tile (asum_out)
    if (epi_in.asum_out) {
        if constexpr (EPI == EPI_F32 && CL == 3 && A_MN && B_MN && BN == 256) {
            const uint64_t dims[2] = {64, (uint64_t)g.K};
            const uint64_t strides[2] = {2, 128};
            const uint32_t box[2] = {64, BK};
            if ((rc = make_tmap_bf16(&tmAux, g.ones, 2, dims, strides, box))) return rc;
            MICO_CHECK_CUDA(cudaMemsetAsync(epi_in.asum_out, 0, (size_t)g.M * sizeof(float), stream));
            asum_cols = 32;
        } else {
            set_last_error(__FILE__, __LINE__, "asum_out: unsupported kernel variant");
            return MICO_ERR_UNSUPPORTED;
        }
    }

    auto kern = gemm_bf16_kernel<BN, A_MN, B_MN, STAGES, CL, EPI>;
    static bool attr_set = false;   // benign race: idempotent
    if (!attr_set) {
        MICO_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
        attr_set = true;
    }
    const int num_n = ceil_div(g.N, BN);
    int units = ceil_div(ceil_div(g.M, BM), CLN) * num_n;
    const int max_clusters = num_sms() / CLN;
    GemmEpi epi = epi_in;
    epi.ksplit = 1;
    epi.sched = nullptr;
    epi.sched_rounds = 0;
    {
        // Split-K for weight gradients (EPI_F32) with fewer output tiles than SM (pair) slots, and the balanced unit schedule
        // (GemmPlan above).
        GemmPlan& plan = get_plan(units, num_n, BN, last_tile_cols(g.N, BN, CL == 3) + asum_cols, ceil_div(g.K, BK), max_clusters,
                                  EPI == EPI_F32);
        if (plan.ks > 1) {
            MICO_CHECK_CUDA(cudaMemset2DAsync(epi.out, (size_t)epi.ldo * 4, 0, (size_t)g.N * 4, (size_t)g.M, stream));
            epi.ksplit = plan.ks;
            units *= plan.ks;
        }
        if ((rc = plan_device_table(plan, &epi.sched))) return rc;
        epi.sched_rounds = plan.rounds;
    }
    const int grid = (units < max_clusters ? units : max_clusters) * CLN;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(kGemmThreads);
    cfg.dynamicSmemBytes = Cfg::SMEM_BYTES;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CLN;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    int Mi = g.M, Ni = g.N, Ki = g.K;
    MICO_CHECK_CUDA(cudaLaunchKernelEx(&cfg, kern, tmA, tmB, tmOut, tmAux, Mi, Ni, Ki, epi));
    count_launch();
    return MICO_OK;
}

static bool g_force_single_cta = false;     // MICO_GEMM_SINGLE_CTA=1: A/B switch for measurements
static bool g_force_generic = false;        // MICO_GEMM_GENERIC_EPI=1: A/B switch for measurements
static bool g_pair_mma = true;              // MICO_GEMM_PAIR_MMA=0: A/B switch for measurements
static bool g_pair_mma_128 = true;          // MICO_GEMM_PAIR_MMA_128=0: A/B switch for measurements
static bool g_pair_mma_wgrad = true;        // MICO_GEMM_PAIR_MMA_WGRAD=0: A/B switch for measurements

// which specialised epilogue (if any) computes exactly what `e` asks for
static int classify_epilogue(const MicoGemmArgs& g, const GemmEpi& e, int bn) {
    if (g_force_generic || !e.vec_ok || e.accumulate || e.remap_gin > 0 || e.alpha != 1.0f) return EPI_GENERIC;
    if (g.N % 32 != 0 || bn == 64 || (bn == 176 && g.N % 176 != 0)) return EPI_GENERIC;
    const bool plain = !e.residual && !e.row_scale;
    if (e.act == MICO_ACT_NONE && plain && !e.aux_out) {
        if (!e.out_fp32) return EPI_BF16;
        return e.bias ? EPI_GENERIC : EPI_F32;
    }
    if ((e.act == MICO_ACT_GELU_SAVE_GRAD || e.act == MICO_ACT_QUICK_GELU_SAVE_GRAD) && plain && !e.out_fp32)
        return EPI_GELU_SAVE;      // aux_out may be null: activation only
    if (e.act == MICO_ACT_MUL_AUX && plain && !e.out_fp32 && !e.bias && !e.aux_out) return EPI_MUL_AUX;
    if (e.act == MICO_ACT_NONE && e.residual && e.out_fp32 && !e.aux_out && !e.residual_bcast) return EPI_RES32;
    return EPI_GENERIC;
}

template <int BN, bool A_MN, bool B_MN, int STAGES, int CL>
int dispatch_epi(const MicoGemmArgs& g, const GemmEpi& epi, cudaStream_t stream) {
    const int kind = classify_epilogue(g, epi, BN);
    // only the (layout, epilogue) pairs a tower actually launches are specialised (compile time)
    if constexpr (BN != 64) {
        if constexpr (!A_MN && !B_MN) {          // forward: y = x W^T
            if (kind == EPI_BF16) return launch_gemm<BN, A_MN, B_MN, STAGES, CL, EPI_BF16>(g, epi, stream);
            if (kind == EPI_GELU_SAVE) return launch_gemm<BN, A_MN, B_MN, STAGES, CL, EPI_GELU_SAVE>(g, epi, stream);
            if (kind == EPI_RES32) return launch_gemm<BN, A_MN, B_MN, STAGES, CL, EPI_RES32>(g, epi, stream);
        } else if constexpr (!A_MN && B_MN) {    // dgrad: dx = dy W
            if (kind == EPI_BF16) return launch_gemm<BN, A_MN, B_MN, STAGES, CL, EPI_BF16>(g, epi, stream);
            if (kind == EPI_MUL_AUX) return launch_gemm<BN, A_MN, B_MN, STAGES, CL, EPI_MUL_AUX>(g, epi, stream);
        } else if constexpr (A_MN && B_MN) {     // wgrad: dW = dy^T x
            if (kind == EPI_F32) return launch_gemm<BN, A_MN, B_MN, STAGES, CL, EPI_F32>(g, epi, stream);
        }
    }
    return launch_gemm<BN, A_MN, B_MN, STAGES, CL, EPI_GENERIC>(g, epi, stream);
}

template <bool A_MN, bool B_MN>
int dispatch_bn(const MicoGemmArgs& g, const GemmEpi& epi, cudaStream_t stream) {
    // Pick the N tile by estimated time: whole waves of (fractional: the last N tile issues a narrower MMA) tile units
    // over the SMs, times the tile width over its measured tensor-pipe efficiency (round 1, ViT-g shapes: 256-wide tiles
    // run at ~1.39 PFLOP/s, 176-wide at ~1.29, 128-wide at ~1.15 -- operand reads per MMA flop grow as the tile narrows).
    const int m_tiles = ceil_div(g.M, BM);
    const double n16 = (double)((g.N + 15) & ~15);
    auto est = [&](int bn, double eff) {
        const double units = (double)m_tiles * (n16 / bn);
        const double waves = (double)(int64_t)((units + num_sms() - 1e-9) / num_sms());
        return (waves < 1.0 ? 1.0 : waves) * bn / eff;
    };
    int best = 256;
    double best_t = est(256, 1.0);
    // 176-wide tiles only with a K-major B: with an MN-major B the tile needs three 64-column boxes and drops to four
    // stages -- measured slower than 128-wide tiles (fc1 dgrad 253 vs 238 us)
    if (!B_MN && g.N % 176 == 0 && est(176, 0.93) < best_t) { best = 176; best_t = est(176, 0.93); }
    if (est(128, 0.85) < best_t) { best = 128; best_t = est(128, 0.85); }
    if constexpr (A_MN && B_MN) {
        // Weight gradients run as CTA pairs and may split K (launch_gemm): estimate with THAT unit count.  Round 2: the plain
        // estimate above gave the qkv / proj gradients of the 197 376-token tower pass 128-wide tiles (187 / 66 pair units, no
        // split, ~1.0 PFLOP/s); 256-wide tiles with K split in two fill the 74 SM pairs as well (204 / 72 units) at the wider
        // tile's efficiency.
        static const bool splitk_aware = [] { const char* e = getenv("MICO_GEMM_WGRAD_TILE_SPLITK"); return !(e && e[0] == '0'); }();
        if (splitk_aware && m_tiles >= 2 && g.N % 128 == 0 && g_pair_mma && g_pair_mma_wgrad && g_pair_mma_128 && !g_force_single_cta) {
            const int slots = num_sms() / 2, num_kb = ceil_div(g.K, BK);
            auto est_w = [&](int bn, double eff) {
                const int nn = ceil_div(g.N, bn);
                return get_plan(ceil_div(m_tiles, 2) * nn, nn, bn, last_tile_cols(g.N, bn, true), num_kb, slots, true).makespan / eff;
            };
            best = est_w(256, 1.0) <= est_w(128, 0.85) ? 256 : 128;
        }
    }
    if (g.N <= 64) best = 64;
    // CTA pairs (B multicast) whenever there are at least two M tiles to pair up
    // (not for wgrad, A and B both MN-major with a 16k-long K loop: measured 4-15 % slower in lock step)
    const bool pair = ceil_div(g.M, BM) >= 2 && !g_force_single_cta && !(A_MN && B_MN);
    // cta_group::2 variant of the 256-wide pair: ONE M = 256 MMA per pair, each CTA staging its own A rows and its half of B
    // (six 32 KB stages).  Measured on the ViT-g shapes: plain fc1 1.38 -> 1.47 PFLOP/s (cuBLAS 1.43-1.45), qkv fwd 1.36 -> 1.46,
    // fc2 fwd + residual 1.16 -> 1.25, fc2 dgrad x GELU' 1.21 -> 1.29, fc2 / fc1 wgrad 1.28 -> 1.35 / 1.35 -> 1.39; bench step
    // 123.8 -> 119.3 ms on one box.  MICO_GEMM_PAIR_MMA=0 falls back to the multicast pair, MICO_GEMM_PAIR_MMA_WGRAD=0 keeps
    // single CTAs for wgrad (A and B both MN-major).
    if constexpr (A_MN == B_MN || !A_MN) {
        const bool wgrad = A_MN && B_MN;
        if (best == 256 && ceil_div(g.M, BM) >= 2 && !g_force_single_cta && g_pair_mma && g.N % 128 == 0 &&
            (!wgrad || g_pair_mma_wgrad))
            return dispatch_epi<256, A_MN, B_MN, 6, 3>(g, epi, stream);
        if (best == 128 && ceil_div(g.M, BM) >= 2 && !g_force_single_cta && g_pair_mma && g_pair_mma_128 && g.N % 128 == 0 &&
            (!wgrad || g_pair_mma_wgrad))
            return dispatch_epi<128, A_MN, B_MN, 8, 3>(g, epi, stream);     // 256 x 128 pair tile, eight 24 KB stages: qkv wgrad
                                                                            // 184 -> 164 us, bench step 119.1 -> 117.5 ms
    }
    switch (best) {
        case 256: return pair ? dispatch_epi<256, A_MN, B_MN, 4, 2>(g, epi, stream)
                              : dispatch_epi<256, A_MN, B_MN, 4, 1>(g, epi, stream);
        case 176: if constexpr (!(A_MN && B_MN)) return pair ? dispatch_epi<176, A_MN, B_MN, B_MN ? 4 : 5, 2>(g, epi, stream)
                                                             : dispatch_epi<176, A_MN, B_MN, B_MN ? 4 : 5, 1>(g, epi, stream);
        case 128: return pair ? dispatch_epi<128, A_MN, B_MN, 6, 2>(g, epi, stream)
                              : dispatch_epi<128, A_MN, B_MN, 6, 1>(g, epi, stream);
        default:  return launch_gemm<64, A_MN, B_MN, 8, 1, EPI_GENERIC>(g, epi, stream);
    }
}

}  // namespace
}  // namespace mico

extern "C" int mico_gemm_plan(int num_tiles, int num_n, int bn, int n_last, int num_kb, int slots, int allow_split, int* ks,
                              int* rounds, double* makespan, int* table, int table_cap) {
    using namespace mico;
    MICO_CHECK_ARG(num_tiles > 0 && num_n > 0 && num_tiles % num_n == 0 && bn > 0 && n_last > 0 && n_last <= bn);
    MICO_CHECK_ARG(num_kb > 0 && slots > 0);
    GemmPlan& p = get_plan(num_tiles, num_n, bn, n_last, num_kb, slots, allow_split != 0);
    if (ks) *ks = p.ks;
    if (rounds) *rounds = p.rounds;
    if (makespan) *makespan = p.makespan;
    if (table) {
        MICO_CHECK_ARG(table_cap >= (int)p.table.size());
        for (size_t i = 0; i < p.table.size(); ++i) table[i] = p.table[i];
    }
    return MICO_OK;
}

extern "C" int mico_gemm_bf16(const MicoGemmArgs* args, void* stream_) {
    using namespace mico;
    static const bool single = [] { const char* e = getenv("MICO_GEMM_SINGLE_CTA"); return e && e[0] == '1'; }();
    static const bool generic = [] { const char* e = getenv("MICO_GEMM_GENERIC_EPI"); return e && e[0] == '1'; }();
    g_force_single_cta = single;
    g_force_generic = generic;
    static const bool pair_mma = [] { const char* e = getenv("MICO_GEMM_PAIR_MMA"); return !(e && e[0] == '0'); }();
    static const bool pair_mma_wgrad = [] { const char* e = getenv("MICO_GEMM_PAIR_MMA_WGRAD"); return !(e && e[0] == '0'); }();
    g_pair_mma = pair_mma;
    g_pair_mma_wgrad = pair_mma_wgrad;
    static const bool pair_mma_128 = [] { const char* e = getenv("MICO_GEMM_PAIR_MMA_128"); return !(e && e[0] == '0'); }();
    g_pair_mma_128 = pair_mma_128;
    MICO_CHECK_ARG(args != nullptr);
    const MicoGemmArgs& g = *args;
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    MICO_CHECK_ARG(g.a && g.b && g.out);
    MICO_CHECK_ARG(g.M > 0 && g.N > 0 && g.K > 0);
    MICO_CHECK_ARG(g.lda % 8 == 0 && g.ldb % 8 == 0);
    MICO_CHECK_ARG((reinterpret_cast<uintptr_t>(g.a) & 15) == 0 && (reinterpret_cast<uintptr_t>(g.b) & 15) == 0);
    MICO_CHECK_ARG(!(g.accumulate && !g.out_fp32));
    MICO_CHECK_ARG(g.act >= MICO_ACT_NONE && g.act <= MICO_ACT_MUL_AUX);
    MICO_CHECK_ARG(!((g.act == MICO_ACT_GELU_BWD || g.act == MICO_ACT_QUICK_GELU_BWD || g.act == MICO_ACT_MUL_AUX) && !g.aux_in));
    MICO_CHECK_ARG(!(g.row_scale && g.rows_per_group <= 0));
    MICO_CHECK_ARG(!(g.remap_gin > 0 && g.remap_gout < g.remap_gin));
    if (g.a_mn_major) MICO_CHECK_ARG(g.lda >= g.M); else MICO_CHECK_ARG(g.lda >= g.K);
    if (g.b_mn_major) MICO_CHECK_ARG(g.ldb >= g.N); else MICO_CHECK_ARG(g.ldb >= g.K);

    GemmEpi e;
    e.out = g.out; e.ldo = g.ldo; e.out_fp32 = g.out_fp32;
    e.bias = g.bias; e.residual = g.residual; e.ldr = g.ldr;
    e.row_scale = g.row_scale; e.rows_per_group = g.rows_per_group;
    e.act = g.act;
    e.aux_out = reinterpret_cast<__nv_bfloat16*>(g.aux_out); e.ld_aux_out = g.ld_aux_out;
    e.aux_in = reinterpret_cast<const __nv_bfloat16*>(g.aux_in); e.ld_aux_in = g.ld_aux_in;
    e.accumulate = g.accumulate; e.alpha = g.alpha;
    e.remap_gin = g.remap_gin; e.remap_gout = g.remap_gout; e.remap_off = g.remap_off;
    e.residual_bcast = g.residual_bcast;
    e.ksplit = 1;
    e.asum_out = g.asum_out;
    e.sched = nullptr;
    e.sched_rounds = 0;
    auto al16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
    const int oel = g.out_fp32 ? 4 : 8;   // elements per 16 bytes
    bool vec = al16(g.out) && (g.ldo % oel == 0);
    if (g.bias) vec = vec && al16(g.bias);
    if (g.residual) vec = vec && al16(g.residual) && (g.ldr % 4 == 0);
    if (g.aux_out) vec = vec && al16(g.aux_out) && (g.ld_aux_out % 8 == 0);
    if (g.aux_in) vec = vec && al16(g.aux_in) && (g.ld_aux_in % 8 == 0);
    e.vec_ok = vec ? 1 : 0;
    ProfScope prof(kProfGemm, 2.0 * g.M * (double)g.N * g.K, stream);

    if (g.asum_out) {
        // bias gradient from the weight-gradient pass: only the cta_group::2 256-wide wgrad kernel carries the ones columns
        MICO_CHECK_ARG(g.ones != nullptr && (reinterpret_cast<uintptr_t>(g.ones) & 127) == 0);
        MICO_CHECK_ARG(g.a_mn_major && g.b_mn_major && g.out_fp32 && !g.bias && !g.residual && !g.row_scale && !g.accumulate);
        MICO_CHECK_ARG(g.act == MICO_ACT_NONE && g.alpha == 1.0f && g.remap_gin == 0 && !g.aux_out);
        const int n_last = g.N % 256;
        const bool ok = e.vec_ok && g.M >= 256 && n_last > 96 && n_last <= 160 && n_last % 32 == 0 && g.N % 128 == 0 &&
                        g_pair_mma && g_pair_mma_wgrad && !g_force_single_cta && !g_force_generic;
        if (!ok) {
            set_last_error(__FILE__, __LINE__, "asum_out: shape / configuration not supported by the fused bias-gradient columns");
            return MICO_ERR_UNSUPPORTED;
        }
        return launch_gemm<256, true, true, 6, 3, EPI_F32>(g, e, stream);
    }

    if (!g.a_mn_major && !g.b_mn_major) return dispatch_bn<false, false>(g, e, stream);
    if (!g.a_mn_major && g.b_mn_major) return dispatch_bn<false, true>(g, e, stream);
    if (g.a_mn_major && !g.b_mn_major) return dispatch_bn<true, false>(g, e, stream);
    return dispatch_bn<true, true>(g, e, stream);
}
