// mico_b200 -- K4 backward: fused attention gradients (sm_100a, tcgen05), no score matrix in HBM.
//
// Autograd of eva_vit_model.py:340-361 / bert.py:233-277.  With P = softmax(scale*QK^T + mask) recomputed
// from Q, K and the saved log-sum-exp, and delta_i = sum_d dO_id O_id:
//     dV = P^T dO          dP = dO V^T          dS = scale * P o (dP - delta)
//     dQ = dS K            dK = dS^T Q
// Launches:
//   attn_delta_kernel   delta[b,h,i]                                     (HBM-bound row reduction)
//   attn_dq_kernel      CTA per (b,h,q-tile), loops over KV tiles:  S, dP -> dS (TMEM) -> dQ += dS K
//   attn_dkv_kernel     CTA per (b,h,kv-tile), loops over Q tiles:  S^T, dP^T -> P^T, dS^T (TMEM)
//                                                                  -> dV += P^T dO, dK += dS^T Q
//   attention_tail_bwd  (attention_tail.cu) the <= 8 remainder rows of the owned side, SIMT
// Both MMA kernels recompute S (7 GEMMs instead of 5) so that no atomics are needed and the result is
// deterministic.  The dS / P^T / dS^T operands never leave tensor memory: the softmax warps write them back as
// packed bf16 (tcgen05.st) and the gradient MMAs read them as the TMEM A operand; K, Q and dO tiles are read
// K-major for the score GEMMs and MN-major (same bytes) for the gradient GEMMs.  Streamed tiles hold 128 rows,
// the last one up to 144 (S = 257 -> 128 + 129); 8 softmax warps (two per TMEM lane quarter) split the columns.
#include "attn_common.cuh"

namespace mico {
namespace {

struct AttnBwdParams {
    int B, H, Sq, Sk, D;
    float scale;
    const float* mask;
    int64_t mask_bs, mask_qs, mask_hs;
    int mask_bmod;
    DropCfg drop;
    const float* lse;     // [B,H,Sq]
    const float* delta;   // [B,H,Sq]
    __nv_bfloat16* dq; int64_t dq_bs, dq_rs, dq_hs;
    __nv_bfloat16* dk; int64_t dk_bs, dk_rs, dk_hs;
    __nv_bfloat16* dv; int64_t dv_bs, dv_rs, dv_hs;
    // K/V shared by several query batch entries (MicoAttnArgs::kv_index): n_kv entries; entry e is read by the query
    // entries grp_list[grp_ptr[e] .. grp_ptr[e+1]).  kv_index == null: identity.
    const int* kv_index; int n_kv; const int* grp_ptr; const int* grp_list;
};

// ------------------------------------------------------------------------------------------------ delta
// A quad of lanes per (b, i, h) row: lane `sub` takes the 16-byte chunks sub, sub+4, ... of O and dO, so a warp
// instruction reads 8 rows x 64 contiguous bytes (the rows of consecutive heads are adjacent in memory).
__global__ void attn_delta_kernel(const __nv_bfloat16* __restrict__ o, int64_t o_bs, int64_t o_rs, int64_t o_hs,
                                  const __nv_bfloat16* __restrict__ d_o, int64_t do_bs, int64_t do_rs, int64_t do_hs,
                                  float* __restrict__ delta, int B, int H, int S, int D) {
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t idx = tid >> 2;                                         // over (b, i, h), h fastest
    const int sub = (int)(tid & 3);
    const bool live = idx < (int64_t)B * S * H;
    float acc = 0.f;
    int h = 0, i = 0, b = 0;
    if (live) {
        h = (int)(idx % H);
        i = (int)((idx / H) % S);
        b = (int)(idx / ((int64_t)H * S));
        const uint4* po = reinterpret_cast<const uint4*>(o + b * o_bs + i * o_rs + h * o_hs);
        const uint4* pd = reinterpret_cast<const uint4*>(d_o + b * do_bs + i * do_rs + h * do_hs);
        uint4 a[4], g[4];
#pragma unroll
        for (int t = 0; t < 4; ++t) {
            const int c = sub + 4 * t;
            const bool ok = c < D / 8;
            a[t] = ok ? po[c] : make_uint4(0, 0, 0, 0);
            g[t] = ok ? pd[c] : make_uint4(0, 0, 0, 0);
        }
#pragma unroll
        for (int t = 0; t < 4; ++t)
            acc += bf16_lo(a[t].x) * bf16_lo(g[t].x) + bf16_hi(a[t].x) * bf16_hi(g[t].x) + bf16_lo(a[t].y) * bf16_lo(g[t].y) +
                   bf16_hi(a[t].y) * bf16_hi(g[t].y) + bf16_lo(a[t].z) * bf16_lo(g[t].z) + bf16_hi(a[t].z) * bf16_hi(g[t].z) +
                   bf16_lo(a[t].w) * bf16_lo(g[t].w) + bf16_hi(a[t].w) * bf16_hi(g[t].w);
    }
    acc += __shfl_xor_sync(0xffffffffu, acc, 1);      // grid is a multiple of the warp size: all lanes are here
    acc += __shfl_xor_sync(0xffffffffu, acc, 2);
    if (live && sub == 0) delta[((int64_t)b * H + h) * S + i] = acc;
}

// ------------------------------------------------------------------------------------------------ dQ
// TMEM columns: S [0,160)  dP [160,320)  dS as bf16 A operand [320,392)  dQ accumulator [400,400+HD_PAD)
struct DqSmem {
    static constexpr int Q = 0;                                   // 2 atoms x 128 rows
    static constexpr int DO = 2 * kAtomBytes;
    static constexpr int K0 = 4 * kAtomBytes;                     // 2 stages x 2 atoms x 144 rows
    static constexpr int V0 = K0 + 4 * kAtomBytesN;
    static constexpr int BARS = V0 + 4 * kAtomBytesN;
    static constexpr int TOTAL = BARS + 256 + 1024;
};
constexpr uint32_t kDqS = 0, kDqDP = 160, kDqDS = 320, kDqAcc = 400;

// kMode 1 (plain): no additive mask and no dropout (every ViT tower): the per-element mask / dropout code is compiled out.
// kMode 2: dropout without a mask (fusion-encoder cross-attention): no mask code, full chunks take a branch-free path.
// kMode 0: mask and dropout decided at run time.
template <int HD_PAD, int kMode>
__global__ void __launch_bounds__(kBwdThreads, 1)
attn_dq_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
               const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmDO,
               const __grid_constant__ CUtensorMap tmKx, const __grid_constant__ CUtensorMap tmVx, AttnBwdParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + DqSmem::BARS);
    uint64_t* q_full = bars + 0;     // Q + dO tiles landed
    uint64_t* q_empty = bars + 1;    // all MMAs of the work item done
    uint64_t* kv_full = bars + 2;    // [2]
    uint64_t* kv_empty = bars + 4;   // [2]
    uint64_t* sdp_full = bars + 6;   // S and dP in TMEM
    uint64_t* sdp_empty = bars + 7;  // softmax finished reading them
    uint64_t* ds_full = bars + 8;    // dS operand in TMEM
    uint64_t* ds_empty = bars + 9;   // dQ MMA finished reading it
    uint64_t* dq_full = bars + 10;   // dQ accumulator final
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 12);

    const int warp = threadIdx.x >> 5;
    const int nqt = m_tiles(p.Sq);
    const NTiling kt = n_tiling(p.Sk);
    const int nkv = kt.n;
    const int num_work = p.B * p.H * nqt;
    constexpr int kAtoms = (HD_PAD + 63) / 64;
    constexpr uint32_t kTileBytes = kAtoms * kAtomBytes;

    if (warp == 8) {
        if (elect_one()) {
            tma_prefetch_desc(&tmQ); tma_prefetch_desc(&tmK); tma_prefetch_desc(&tmV); tma_prefetch_desc(&tmDO);
            tma_prefetch_desc(&tmKx); tma_prefetch_desc(&tmVx);
        }
    } else if (warp == 9) {
        if (elect_one()) {
            mbar_init(q_full, 1); mbar_init(q_empty, 1);
            for (int i = 0; i < 2; ++i) { mbar_init(&kv_full[i], 1); mbar_init(&kv_empty[i], 1); }
            mbar_init(sdp_full, 1); mbar_init(sdp_empty, 8);
            mbar_init(ds_full, 8); mbar_init(ds_empty, 1);
            mbar_init(dq_full, 1);
            fence_mbar_init();
        }
        __syncwarp();
        tmem_alloc<512>(tmem_slot);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t tmem_S = tmem_base + kDqS, tmem_dP = tmem_base + kDqDP, tmem_dS = tmem_base + kDqDS,
                   tmem_dQ = tmem_base + kDqAcc;

    if (warp == 8) {
        if (elect_one()) {
            uint32_t wcount = 0, kvcount = 0;
            for (int w = blockIdx.x; w < num_work; w += gridDim.x, ++wcount) {
                const int qt = w % nqt, bh = w / nqt;
                const int h = bh % p.H, b = bh / p.H;
                mbar_wait(q_empty, (wcount & 1) ^ 1);
                mbar_arrive_expect_tx(q_full, 2 * kTileBytes);
#pragma unroll
                for (int a = 0; a < kAtoms; ++a) {
                    tma_load_4d(smem + DqSmem::Q + a * kAtomBytes, &tmQ, q_full, a * 64, h, qt * kTile, b);
                    tma_load_4d(smem + DqSmem::DO + a * kAtomBytes, &tmDO, q_full, a * 64, h, qt * kTile, b);
                }
                const int kvb = p.kv_index ? __ldg(p.kv_index + b) : b;
                for (int j = 0; j < nkv; ++j, ++kvcount) {
                    const int s = kvcount & 1;
                    mbar_wait(&kv_empty[s], ((kvcount >> 1) & 1) ^ 1);
                    const bool ext = n_valid(kt, j) > kTile;
                    mbar_arrive_expect_tx(&kv_full[s], 2 * n_tile_bytes(kAtoms, ext));
                    load_n_tile<kAtoms>(smem + DqSmem::K0 + s * 2 * kAtomBytesN, &tmK, &tmKx, &kv_full[s], h, j * kTile, kvb, ext);
                    load_n_tile<kAtoms>(smem + DqSmem::V0 + s * 2 * kAtomBytesN, &tmV, &tmVx, &kv_full[s], h, j * kTile, kvb, ext);
                }
            }
        }
    } else if (warp == 9) {
        if (elect_one()) {
            uint32_t wcount = 0, kvcount = 0, tcount = 0;   // tcount: kv tiles processed (sdp / ds barrier phases)
            const uint32_t sQ = smem_u32(smem + DqSmem::Q), sDO = smem_u32(smem + DqSmem::DO);
            constexpr uint32_t idesc_dq = umma_idesc_bf16(HD_PAD, false, true);
            auto issue_scores = [&](int j, uint32_t kvc, uint32_t tc) {
                const int s = kvc & 1;
                mbar_wait(&kv_full[s], (kvc >> 1) & 1);
                mbar_wait(sdp_empty, (tc & 1) ^ 1);
                tc_fence_after();
                const int valid = n_valid(kt, j);
                const uint32_t idesc = umma_idesc_bf16(max(16, (valid + 15) & ~15), false, false);
                const uint32_t sK = smem_u32(smem + DqSmem::K0 + s * 2 * kAtomBytesN);
                const uint32_t sV = smem_u32(smem + DqSmem::V0 + s * 2 * kAtomBytesN);
#pragma unroll
                for (int k = 0; k < HD_PAD / 16; ++k)
                    umma_bf16_ss(tmem_S, umma_smem_desc_sw128(sQ + (k >> 2) * kAtomBytes + (k & 3) * 32, 16, 1024),
                                 umma_smem_desc_sw128(sK + (k >> 2) * kAtomBytesN + (k & 3) * 32, 16, 1024), idesc, k != 0);
#pragma unroll
                for (int k = 0; k < HD_PAD / 16; ++k)
                    umma_bf16_ss(tmem_dP, umma_smem_desc_sw128(sDO + (k >> 2) * kAtomBytes + (k & 3) * 32, 16, 1024),
                                 umma_smem_desc_sw128(sV + (k >> 2) * kAtomBytesN + (k & 3) * 32, 16, 1024), idesc, k != 0);
                umma_commit(sdp_full);
                // Q and dO are read by the score MMAs only: once the last pair of the work item is issued, the producer
                // may refill them for the next item while this item's softmax, dQ MMA and epilogue are still running
                if (j == nkv - 1) umma_commit(q_empty);
            };
            // The (work item, kv tile) steps form one flat stream: the scores of step t+1 -- the next kv tile, or the
            // first tile of the NEXT work item -- are issued before the dQ MMA of step t is waited for, so the TMA
            // latency and the score MMAs of a new work item hide behind the tail of the previous one.
            if ((int)blockIdx.x < num_work) {
                mbar_wait(q_full, 0);
                tc_fence_after();
                issue_scores(0, kvcount, tcount);
            }
            for (int w = blockIdx.x; w < num_work; w += gridDim.x, ++wcount) {
                for (int j = 0; j < nkv; ++j) {
                    if (j + 1 < nkv) {
                        issue_scores(j + 1, kvcount + 1, tcount + 1);
                    } else if (w + (int)gridDim.x < num_work) {
                        mbar_wait(q_full, (wcount + 1) & 1);
                        tc_fence_after();
                        issue_scores(0, kvcount + 1, tcount + 1);
                    }
                    const int s = kvcount & 1;
                    mbar_wait(ds_full, tcount & 1);
                    tc_fence_after();
                    const int valid = n_valid(kt, j);
                    const int ksteps = (valid + 15) >> 4;
                    const uint32_t sK = smem_u32(smem + DqSmem::K0 + s * 2 * kAtomBytesN);
                    for (int k = 0; k < ksteps; ++k)    // dQ += dS(TMEM, 16 keys = 8 columns per step) . K(MN-major)
                        umma_bf16_ts(tmem_dQ, tmem_dS + k * 8, umma_smem_desc_sw128(sK + k * 2048, kAtomBytesN, 1024),
                                     idesc_dq, (j | k) != 0);
                    umma_commit(&kv_empty[s]);
                    umma_commit(ds_empty);
                    ++kvcount; ++tcount;
                }
                umma_commit(dq_full);
            }
        }
    } else {
        // softmax warps 0..7: TMEM lane quarter = warp & 3; the two warps of a quarter alternate 32-key chunks
        const int r = threadIdx.x & 127;
        const int half = warp >> 2;
        const uint32_t lane_off = (uint32_t)((warp & 3) * 32) << 16;
        const float sc2 = p.scale * kLog2e;
        uint32_t tcount = 0, wcount = 0;
        // row statistics of a work item (rows past Sq get lse = +huge -> p = 0 -> dS = 0 with no per-element predicate);
        // the next item's are requested during the last kv tile of the current one
        auto load_stats = [&](int w, float& l, float& d) {
            const int qt = w % nqt, bh = w / nqt;
            const int qi = qt * kTile + r;
            if (qi < p.Sq) {
                const int64_t stat = (int64_t)bh * p.Sq + qi;
                l = ldg_f32_pinned(p.lse + stat);
                d = ldg_f32_pinned(p.delta + stat);
            } else {
                l = 1e30f;
                d = 0.f;
            }
        };
        float lse_n = 0.f, dlt_n = 0.f;
        if ((int)blockIdx.x < num_work) load_stats(blockIdx.x, lse_n, dlt_n);
        for (int w = blockIdx.x; w < num_work; w += gridDim.x, ++wcount) {
            const int qt = w % nqt, bh = w / nqt;
            const int h = bh % p.H, b = bh / p.H;
            const int qi = qt * kTile + r;
            const bool row_ok = qi < p.Sq;
            const float nlse2 = -lse_n * kLog2e;
            const float ndlt = -dlt_n * p.scale;     // dS = p * (dP*scale - delta*scale)
            const float* mrow = (kMode == 0 && p.mask) ? (p.mask + (int64_t)(p.mask_bmod ? b % p.mask_bmod : b) * p.mask_bs + (int64_t)h * p.mask_hs) + (int64_t)(row_ok ? qi : 0) * p.mask_qs : nullptr;
            const bool dropping = kMode == 2 || (kMode == 0 && p.drop.p > 0.f);
            const uint32_t drop_key = dropping ? drop_row_key(p.drop, (uint64_t)(b * p.H + h) * p.Sq + (row_ok ? qi : 0)) : 0u;
            const uint32_t drop_thr = drop_thresh16(p.drop);
            const float inv_keep = p.drop.inv_keep, scale = p.scale;
            for (int j = 0; j < nkv; ++j, ++tcount) {
                const int valid = n_valid(kt, j);
                const int nch = (valid + 31) >> 5;
                if (j == nkv - 1 && w + (int)gridDim.x < num_work) load_stats(w + gridDim.x, lse_n, dlt_n);
                mbar_wait(sdp_full, tcount & 1);
                tc_fence_after();
                mbar_wait(ds_empty, (tcount & 1) ^ 1);   // previous dQ MMA no longer reads the dS operand
                tc_fence_after();
                for (int c = half; c < nch; c += 2) {
                    uint32_t sv[32], dv[32];
                    tmem_ld_x32(tmem_S + lane_off + c * 32, sv);
                    tmem_ld_x32(tmem_dP + lane_off + c * 32, dv);
                    tmem_ld_wait();
                    float ds[32];
                    const int lim = valid - c * 32;
                    if (!mrow && lim >= 32 && !dropping) {
#pragma unroll
                        for (int i = 0; i < 32; ++i)
                            ds[i] = ex2_fast(fmaf(__uint_as_float(sv[i]), sc2, nlse2)) * fmaf(__uint_as_float(dv[i]), p.scale, ndlt);
                    } else if (!mrow && lim >= 32) {       // dropout only, full chunk: no per-element predicates
                        const uint32_t pair0 = (uint32_t)(j * kTile + c * 32) >> 1;
#pragma unroll
                        for (int i = 0; i < 32; i += 2) {
                            const uint32_t bits = drop_pair_bits(drop_key, pair0 + (i >> 1));
                            // dP flows only through the kept probabilities
                            const float m0 = (bits & 0xFFFFu) >= drop_thr ? inv_keep : 0.0f;
                            const float m1 = (bits >> 16) >= drop_thr ? inv_keep : 0.0f;
                            ds[i] = ex2_fast(fmaf(__uint_as_float(sv[i]), sc2, nlse2)) * fmaf(__uint_as_float(dv[i]) * m0, scale, ndlt);
                            ds[i + 1] = ex2_fast(fmaf(__uint_as_float(sv[i + 1]), sc2, nlse2)) * fmaf(__uint_as_float(dv[i + 1]) * m1, scale, ndlt);
                        }
                    } else {
#pragma unroll
                        for (int i = 0; i < 32; i += 2) {
                            const uint32_t bits = dropping ? drop_pair_bits(drop_key, (uint32_t)(j * kTile + c * 32 + i) >> 1) : 0u;
#pragma unroll
                            for (int e = 0; e < 2; ++e) {
                                float s = fmaf(__uint_as_float(sv[i + e]), sc2, nlse2);
                                if (mrow && i + e < lim) s = fmaf(mrow[j * kTile + c * 32 + i + e], kLog2e, s);
                                float dpv = __uint_as_float(dv[i + e]);
                                if (dropping)      // dP flows only through the kept probabilities
                                    dpv *= ((e ? bits >> 16 : bits & 0xFFFFu) >= drop_thr) ? p.drop.inv_keep : 0.0f;
                                ds[i + e] = i + e < lim ? ex2_fast(s) * fmaf(dpv, p.scale, ndlt) : 0.f;
                            }
                        }
                    }
                    tmem_store_bf16x32(tmem_dS + lane_off + c * 16, ds);
                }
                tmem_st_wait();
                tc_fence_before();
                __syncwarp();
                if (lane_id() == 0) {
                    mbar_arrive(sdp_empty);
                    mbar_arrive(ds_full);
                }
            }
            if (half == 0) {
                mbar_wait(dq_full, wcount & 1);
                tc_fence_after();
                float acc[HD_PAD];
                tmem_load_row<HD_PAD, false>(tmem_dQ + lane_off, acc);
                tc_fence_before();
                if (row_ok)
                    store_row_bf16<HD_PAD>(p.dq + (int64_t)b * p.dq_bs + (int64_t)qi * p.dq_rs + (int64_t)h * p.dq_hs, acc, p.D, 1.0f);
                // the next work item's first dQ MMA is ordered after these loads: it waits for ds_full, on which
                // this warp arrives only later in program order
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 9) {
        tc_fence_after();
        tmem_dealloc<512>(tmem_base);
    }
}

// ------------------------------------------------------------------------------------------------ dK, dV
// TMEM columns: S^T [0,144)  dP^T [144,288)  dV [288,288+HD_PAD)  dK [384,384+HD_PAD).  P^T and dS^T (bf16 A operands)
// overwrite the S^T / dP^T columns in place: the two warps of a lane quarter own disjoint column ranges
// [0,hA) and [hA,n16) of the tile and write their packed output at the start of their own range, behind their reads.
struct DkvSmem {
    static constexpr int K = 0;                                   // 2 atoms x 128 rows
    static constexpr int V = 2 * kAtomBytes;
    static constexpr int Q0 = 4 * kAtomBytes;                     // 2 stages x 2 atoms x 144 rows
    static constexpr int DO0 = Q0 + 4 * kAtomBytesN;
    static constexpr int STATS = DO0 + 4 * kAtomBytesN;           // 2 stages x (-lse*log2e [256], -delta*scale [256], dropout row key [256])
    static constexpr int BARS = STATS + 2 * 3072;
    static constexpr int TOTAL = BARS + 256 + 1024;
};
constexpr uint32_t kKvST = 0, kKvDPT = 144, kKvDV = 288, kKvDK = 384;

__device__ __forceinline__ int split_a(int n16) { return ((n16 / 16 + 1) / 2) * 16; }   // columns owned by half 0

template <int HD_PAD, int kMode>
__global__ void __launch_bounds__(kBwdThreads, 1)
attn_dkv_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmDO,
                const __grid_constant__ CUtensorMap tmQx, const __grid_constant__ CUtensorMap tmDOx, AttnBwdParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + DkvSmem::BARS);
    uint64_t* kv_full = bars + 0;
    uint64_t* kv_empty = bars + 1;
    uint64_t* qdo_full = bars + 2;    // [2]
    uint64_t* qdo_empty = bars + 4;   // [2]
    uint64_t* st_full = bars + 6;     // S^T, dP^T in TMEM
    uint64_t* pds_full = bars + 7;    // P^T, dS^T operands in TMEM
    uint64_t* acc_full = bars + 9;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 12);

    const int warp = threadIdx.x >> 5;
    const bool shared_kv = p.kv_index != nullptr;
    const int nkt = m_tiles_nt(p.Sk, shared_kv);
    const NTiling qtl = n_tiling(p.Sq);
    const int nqt = qtl.n;
    const int num_work = (shared_kv ? p.n_kv : p.B) * p.H * nkt;
    constexpr int kAtoms = (HD_PAD + 63) / 64;
    constexpr uint32_t kTileBytes = kAtoms * kAtomBytes;
    // readers of K/V entry e: the query batch entries whose Q / dO tiles are streamed against this entry's key tile
    auto n_members = [&](int e) { return shared_kv ? __ldg(p.grp_ptr + e + 1) - __ldg(p.grp_ptr + e) : 1; };
    auto member = [&](int e, int mi) { return shared_kv ? __ldg(p.grp_list + __ldg(p.grp_ptr + e) + mi) : e; };

    if (warp == 8) {
        if (elect_one()) {
            tma_prefetch_desc(&tmQ); tma_prefetch_desc(&tmK); tma_prefetch_desc(&tmV); tma_prefetch_desc(&tmDO);
            tma_prefetch_desc(&tmQx); tma_prefetch_desc(&tmDOx);
        }
    } else if (warp == 9) {
        if (elect_one()) {
            mbar_init(kv_full, 1); mbar_init(kv_empty, 1);
            for (int i = 0; i < 2; ++i) { mbar_init(&qdo_full[i], 1); mbar_init(&qdo_empty[i], 1); }
            mbar_init(st_full, 1);
            mbar_init(pds_full, 8);
            mbar_init(acc_full, 1);
            fence_mbar_init();
        }
        __syncwarp();
        tmem_alloc<512>(tmem_slot);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t tmem_ST = tmem_base + kKvST, tmem_dPT = tmem_base + kKvDPT, tmem_dV = tmem_base + kKvDV,
                   tmem_dK = tmem_base + kKvDK;

    if (warp == 8) {
        if (elect_one()) {
            uint32_t wcount = 0, qcount = 0;
            for (int w = blockIdx.x; w < num_work; w += gridDim.x, ++wcount) {
                const int jt = w % nkt, bh = w / nkt;
                const int h = bh % p.H, e = bh / p.H;
                mbar_wait(kv_empty, (wcount & 1) ^ 1);
                mbar_arrive_expect_tx(kv_full, 2 * kTileBytes);
#pragma unroll
                for (int a = 0; a < kAtoms; ++a) {
                    tma_load_4d(smem + DkvSmem::K + a * kAtomBytes, &tmK, kv_full, a * 64, h, jt * kTile, e);
                    tma_load_4d(smem + DkvSmem::V + a * kAtomBytes, &tmV, kv_full, a * 64, h, jt * kTile, e);
                }
                const int nmem = n_members(e);
                for (int mi = 0; mi < nmem; ++mi) {
                    const int qb = member(e, mi);
                    for (int i = 0; i < nqt; ++i, ++qcount) {
                        const int s = qcount & 1;
                        mbar_wait(&qdo_empty[s], ((qcount >> 1) & 1) ^ 1);
                        const bool ext = n_valid(qtl, i) > kTile;
                        mbar_arrive_expect_tx(&qdo_full[s], 2 * n_tile_bytes(kAtoms, ext));
                        load_n_tile<kAtoms>(smem + DkvSmem::Q0 + s * 2 * kAtomBytesN, &tmQ, &tmQx, &qdo_full[s], h, i * kTile, qb, ext);
                        load_n_tile<kAtoms>(smem + DkvSmem::DO0 + s * 2 * kAtomBytesN, &tmDO, &tmDOx, &qdo_full[s], h, i * kTile, qb, ext);
                    }
                }
            }
        }
    } else if (warp == 9) {
        if (elect_one()) {
            uint32_t wcount = 0, qcount = 0;
            const uint32_t sK = smem_u32(smem + DkvSmem::K), sV = smem_u32(smem + DkvSmem::V);
            constexpr uint32_t idesc_g = umma_idesc_bf16(HD_PAD, false, true);
            for (int w = blockIdx.x; w < num_work; w += gridDim.x, ++wcount) {
                mbar_wait(kv_full, wcount & 1);
                tc_fence_after();
                const int nsteps = n_members((w / nkt) / p.H) * nqt;      // (reader, q tile) steps of this work item
                if (nsteps == 0) umma_commit(kv_empty);                   // an entry nobody reads: dK = dV = 0
                for (int st = 0; st < nsteps; ++st, ++qcount) {
                    const int i = st % nqt;
                    const int s = qcount & 1;
                    const uint32_t sQ = smem_u32(smem + DkvSmem::Q0 + s * 2 * kAtomBytesN);
                    const uint32_t sDO = smem_u32(smem + DkvSmem::DO0 + s * 2 * kAtomBytesN);
                    mbar_wait(&qdo_full[s], (qcount >> 1) & 1);
                    tc_fence_after();
                    // S^T / dP^T columns are free: the previous tile's gradient MMAs (which read P^T / dS^T from
                    // them) were issued by this thread earlier and tcgen05.mma executes in issue order.
                    const int validq = n_valid(qtl, i);
                    const int n16 = max(16, (validq + 15) & ~15);
                    const uint32_t idesc = umma_idesc_bf16(n16, false, false);
#pragma unroll
                    for (int k = 0; k < HD_PAD / 16; ++k)
                        umma_bf16_ss(tmem_ST, umma_smem_desc_sw128(sK + (k >> 2) * kAtomBytes + (k & 3) * 32, 16, 1024),
                                     umma_smem_desc_sw128(sQ + (k >> 2) * kAtomBytesN + (k & 3) * 32, 16, 1024), idesc, k != 0);
#pragma unroll
                    for (int k = 0; k < HD_PAD / 16; ++k)
                        umma_bf16_ss(tmem_dPT, umma_smem_desc_sw128(sV + (k >> 2) * kAtomBytes + (k & 3) * 32, 16, 1024),
                                     umma_smem_desc_sw128(sDO + (k >> 2) * kAtomBytesN + (k & 3) * 32, 16, 1024), idesc, k != 0);
                    umma_commit(st_full);
                    // K and V are read by the score MMAs only: after the last pair the producer may already fetch the
                    // next work item's K / V while this item's softmax, gradient MMAs and epilogue run
                    if (st == nsteps - 1) umma_commit(kv_empty);
                    mbar_wait(pds_full, qcount & 1);
                    tc_fence_after();
                    const int ksteps = n16 >> 4;
                    const int hA = split_a(n16);
                    for (int k = 0; k < ksteps; ++k) {
                        const int q0 = k * 16;
                        const uint32_t acol = q0 < hA ? q0 / 2 : hA + (q0 - hA) / 2;
                        umma_bf16_ts(tmem_dV, tmem_ST + acol, umma_smem_desc_sw128(sDO + k * 2048, kAtomBytesN, 1024),
                                     idesc_g, (st | k) != 0);
                    }
                    for (int k = 0; k < ksteps; ++k) {
                        const int q0 = k * 16;
                        const uint32_t acol = q0 < hA ? q0 / 2 : hA + (q0 - hA) / 2;
                        umma_bf16_ts(tmem_dK, tmem_dPT + acol, umma_smem_desc_sw128(sQ + k * 2048, kAtomBytesN, 1024),
                                     idesc_g, (st | k) != 0);
                    }
                    umma_commit(&qdo_empty[s]);
                }
                umma_commit(acc_full);
            }
        }
    } else {
        const int r = threadIdx.x & 127;              // key row within the tile
        const int half = warp >> 2;
        const uint32_t lane_off = (uint32_t)((warp & 3) * 32) << 16;
        const float sc2 = p.scale * kLog2e;
        uint32_t qcount = 0, wcount = 0;
        bool first_tile = true;
        // -lse*log2(e) and -delta*scale of query t (= threadIdx.x < 160) of q tile i of work item w; queries past the
        // tile's valid count get lse = +huge -> p = 0, dS = 0 with no per-element predicate
        const bool dropping = kMode == 2 || (kMode == 0 && p.drop.p > 0.f);
        const uint32_t drop_thr = drop_thresh16(p.drop);
        const float inv_keep = p.drop.inv_keep, scale = p.scale;
        auto load_stats = [&](int qb_, int h_, int i_, float& l, float& d, uint32_t& key) {
            l = 1e30f;       // raw values: the scaling is applied where they are stored, long after the loads were issued
            d = 0.f;
            key = 0u;
            const int t = threadIdx.x;
            if (t < 160 && t < n_valid(qtl, i_)) {
                const int64_t at = ((int64_t)qb_ * p.H + h_) * p.Sq + i_ * kTile + t;
                l = ldg_f32_pinned(p.lse + at);
                d = ldg_f32_pinned(p.delta + at);
                if (dropping) key = drop_row_key(p.drop, (uint64_t)at);       // row = (b*H + h)*Sq + query
            }
        };
        for (int w = blockIdx.x; w < num_work; w += gridDim.x, ++wcount) {
            const int jt = w % nkt, bh = w / nkt;
            const int h = bh % p.H, e = bh / p.H;
            const int kj = jt * kTile + r;
            const bool key_ok = kj < p.Sk;
            const int nmem = n_members(e);
            for (int mi = 0; mi < nmem; ++mi) {
            const int b = member(e, mi);           // the query batch entry of this step
            for (int i = 0; i < nqt; ++i, ++qcount) {
                const int validq = n_valid(qtl, i);
                const int n16 = max(16, (validq + 15) & ~15);
                const int hA = split_a(n16);
                // The lse / delta of a q tile are staged in the stats buffer of parity qcount & 1 ONE TILE AHEAD: the
                // global loads of tile qcount + 1 (possibly the first tile of the next work item) are issued here and
                // stored after this tile's math.  The previous readers of that buffer (tile qcount - 1) all passed this
                // tile's 256-thread barrier before the store.
                float* s_nlse2 = reinterpret_cast<float*>(smem + DkvSmem::STATS + (qcount & 1) * 3072);
                float* s_ndlt = s_nlse2 + 256;
                const uint32_t* s_key = reinterpret_cast<const uint32_t*>(s_nlse2 + 512);
                const uint32_t s_stats = smem_u32(s_nlse2);
                if (first_tile) {      // very first tile of this CTA: nothing was prefetched
                    float l, d;
                    uint32_t key;
                    load_stats(b, h, i, l, d, key);
                    if (threadIdx.x < 160) {
                        s_nlse2[threadIdx.x] = -l * kLog2e; s_ndlt[threadIdx.x] = -d * p.scale;
                        reinterpret_cast<uint32_t*>(s_nlse2 + 512)[threadIdx.x] = key;
                    }
                    first_tile = false;
                }
                float l_n = 0.f, d_n = 0.f;
                uint32_t key_n = 0u;
                const bool more = (i + 1 < nqt) || (mi + 1 < nmem) || (w + (int)gridDim.x < num_work);
                if (more) {      // the step that follows: next q tile of this reader, next reader, or the next work item's first
                    if (i + 1 < nqt) load_stats(b, h, i + 1, l_n, d_n, key_n);
                    else if (mi + 1 < nmem) load_stats(member(e, mi + 1), h, 0, l_n, d_n, key_n);
                    else {
                        const int bh2 = (w + (int)gridDim.x) / nkt;
                        load_stats(member(bh2 / p.H, 0), bh2 % p.H, 0, l_n, d_n, key_n);
                    }
                }
                softmax_group_sync256();
                mbar_wait(st_full, qcount & 1);
                tc_fence_after();
                const int c_begin = half == 0 ? 0 : hA, c_end = half == 0 ? hA : n16;
                const bool masked = kMode == 0 && p.mask != nullptr;
                // dropout without a mask: this thread's key column enters the pair hash as a constant
                const uint32_t kj_d = (uint32_t)min(kj, p.Sk - 1);
                const uint32_t kjc = (kj_d >> 1) * 0x9E3779B1u;
                const bool kodd = kj_d & 1u;
                for (int c0 = c_begin; c0 < c_end; c0 += 32) {
                    uint32_t sv[32], dv[32];
                    float pt[32], dst[32];
                    const bool full = c_end - c0 >= 32;
                    if (full) {
                        tmem_ld_x32(tmem_ST + lane_off + c0, sv);
                        tmem_ld_x32(tmem_dPT + lane_off + c0, dv);
                    } else {   // 16-column remainder of this half
                        uint32_t a16[16], b16[16];
                        tmem_ld_x16(tmem_ST + lane_off + c0, a16);
                        tmem_ld_x16(tmem_dPT + lane_off + c0, b16);
#pragma unroll
                        for (int q = 0; q < 16; ++q) { sv[q] = a16[q]; dv[q] = b16[q]; sv[16 + q] = 0; dv[16 + q] = 0; }
                    }
                    tmem_ld_wait();
                    if (!masked && dropping) {       // no per-element predicates; the row keys of four queries per shared load
#pragma unroll
                        for (int q4 = 0; q4 < 8; ++q4) {
                            const uint4 lu = lds128(s_stats + (c0 + q4 * 4) * 4);
                            const uint4 du = lds128(s_stats + (256 + c0 + q4 * 4) * 4);
                            const uint4 ku = lds128(s_stats + (512 + c0 + q4 * 4) * 4);
                            const uint32_t ls[4] = {lu.x, lu.y, lu.z, lu.w}, dl[4] = {du.x, du.y, du.z, du.w}, ks[4] = {ku.x, ku.y, ku.z, ku.w};
#pragma unroll
                            for (int e = 0; e < 4; ++e) {
                                const int q = q4 * 4 + e;
                                uint32_t x = (ks[e] ^ kjc) * 0x85EBCA6Bu;
                                x ^= x >> 15;
                                const float m = (kodd ? x >> 16 : x & 0xFFFFu) >= drop_thr ? inv_keep : 0.0f;
                                const float pu = ex2_fast(fmaf(__uint_as_float(sv[q]), sc2, __uint_as_float(ls[e])));
                                pt[q] = pu * m;                                          // dV uses the dropped probabilities
                                dst[q] = pu * fmaf(__uint_as_float(dv[q]) * m, scale, __uint_as_float(dl[e]));
                            }
                        }
                    } else
#pragma unroll
                    for (int q4 = 0; q4 < 8; ++q4) {
                        const uint4 lu = lds128(s_stats + (c0 + q4 * 4) * 4);                          // c0 + 32 <= 160
                        const uint4 du = lds128(s_stats + (256 + c0 + q4 * 4) * 4);
                        const float4 l4 = make_float4(__uint_as_float(lu.x), __uint_as_float(lu.y), __uint_as_float(lu.z), __uint_as_float(lu.w));
                        const float4 d4 = make_float4(__uint_as_float(du.x), __uint_as_float(du.y), __uint_as_float(du.z), __uint_as_float(du.w));
                        const float ls[4] = {l4.x, l4.y, l4.z, l4.w}, dl[4] = {d4.x, d4.y, d4.z, d4.w};
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            const int q = q4 * 4 + e;
                            float s = fmaf(__uint_as_float(sv[q]), sc2, ls[e]);
                            if (masked && key_ok && c0 + q < validq)
                                s = fmaf((p.mask + (int64_t)(p.mask_bmod ? b % p.mask_bmod : b) * p.mask_bs + (int64_t)h * p.mask_hs)[(int64_t)(i * kTile + c0 + q) * p.mask_qs + kj], kLog2e, s);
                            const float pu = ex2_fast(s);
                            float m = 1.0f;
                            if (dropping) m = drop_mult_rc(p.drop, s_key[c0 + q], drop_thr, min(kj, p.Sk - 1));
                            pt[q] = pu * m;                                          // dV uses the dropped probabilities
                            dst[q] = pu * fmaf(__uint_as_float(dv[q]) * m, p.scale, dl[e]);
                        }
                    }
                    // packed bf16 output: 16 (or 8) columns at the start of this half's own range, behind its reads
                    const uint32_t ocol = c_begin + (c0 - c_begin) / 2;
                    if (full) {
                        tmem_store_bf16x32(tmem_ST + lane_off + ocol, pt);
                        tmem_store_bf16x32(tmem_dPT + lane_off + ocol, dst);
                    } else {
                        uint32_t w0[8], w1[8];
#pragma unroll
                        for (int q = 0; q < 8; ++q) {
                            w0[q] = pack_bf16x2(pt[2 * q], pt[2 * q + 1]);
                            w1[q] = pack_bf16x2(dst[2 * q], dst[2 * q + 1]);
                        }
                        tmem_st_x8(tmem_ST + lane_off + ocol, w0);
                        tmem_st_x8(tmem_dPT + lane_off + ocol, w1);
                    }
                }
                tmem_st_wait();
                tc_fence_before();
                __syncwarp();
                if (lane_id() == 0) mbar_arrive(pds_full);
                if (more && threadIdx.x < 160) {
                    float* n_nlse2 = reinterpret_cast<float*>(smem + DkvSmem::STATS + ((qcount + 1) & 1) * 3072);
                    n_nlse2[threadIdx.x] = -l_n * kLog2e;
                    n_nlse2[256 + threadIdx.x] = -d_n * p.scale;
                    reinterpret_cast<uint32_t*>(n_nlse2 + 512)[threadIdx.x] = key_n;
                }
            }
            }
            mbar_wait(acc_full, wcount & 1);
            tc_fence_after();
            {
                float acc[HD_PAD];
                if (nmem == 0) {
#pragma unroll
                    for (int q = 0; q < HD_PAD; ++q) acc[q] = 0.f;
                    if (key_ok) {
                        __nv_bfloat16* dst = half == 0 ? p.dv + (int64_t)e * p.dv_bs + (int64_t)kj * p.dv_rs + (int64_t)h * p.dv_hs
                                                       : p.dk + (int64_t)e * p.dk_bs + (int64_t)kj * p.dk_rs + (int64_t)h * p.dk_hs;
                        store_row_bf16<HD_PAD>(dst, acc, p.D, 1.0f);
                    }
                } else if (half == 0) {
                    tmem_load_row<HD_PAD, false>(tmem_dV + lane_off, acc);
                    if (key_ok)
                        store_row_bf16<HD_PAD>(p.dv + (int64_t)e * p.dv_bs + (int64_t)kj * p.dv_rs + (int64_t)h * p.dv_hs, acc, p.D, 1.0f);
                } else {
                    tmem_load_row<HD_PAD, false>(tmem_dK + lane_off, acc);
                    if (key_ok)
                        store_row_bf16<HD_PAD>(p.dk + (int64_t)e * p.dk_bs + (int64_t)kj * p.dk_rs + (int64_t)h * p.dk_hs, acc, p.D, 1.0f);
                }
            }
            tc_fence_before();
            // the next work item's first gradient MMAs overwrite dV/dK: they wait for pds_full, on which every
            // softmax warp arrives only after these loads in program order.
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 9) {
        tc_fence_after();
        tmem_dealloc<512>(tmem_base);
    }
}

}  // namespace
}  // namespace mico

namespace mico {
// attention_bwd_dq.cu: the dQ kernel with double-buffered score tiles (head_dim 81..96)
int attention_dq_pipelined(const MicoAttnArgs* a, cudaStream_t stream);
static bool dq_pipelined_enabled() {      // MICO_ATTN_DQ_PIPELINED=0: A/B switch for measurements
    static const bool on = [] { const char* e = getenv("MICO_ATTN_DQ_PIPELINED"); return !(e && e[0] == '0'); }();
    return on;
}
}  // namespace mico

extern "C" int mico_attention_bwd(const MicoAttnArgs* a, void* stream_) {
    using namespace mico;
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    MICO_CHECK_ARG(a && a->q && a->k && a->v && a->o && a->dout && a->lse && a->delta && a->dq && a->dk && a->dv);
    MICO_CHECK_ARG(a->B > 0 && a->H > 0 && a->Sq > 0 && a->Sk > 0);
    MICO_CHECK_ARG(a->D % 8 == 0 && a->D >= 16 && a->D <= 128);
    for (const int64_t s : {a->q_bs, a->q_rs, a->q_hs, a->k_bs, a->k_rs, a->k_hs, a->v_bs, a->v_rs, a->v_hs, a->o_bs,
                            a->o_rs, a->o_hs, a->do_bs, a->do_rs, a->do_hs, a->dq_bs, a->dq_rs, a->dq_hs, a->dk_bs,
                            a->dk_rs, a->dk_hs, a->dv_bs, a->dv_rs, a->dv_hs})
        MICO_CHECK_ARG(s % 8 == 0);
    for (const void* ptr : {a->q, a->k, a->v, (const void*)a->o, a->dout, (const void*)a->dq, (const void*)a->dk,
                            (const void*)a->dv})
        MICO_CHECK_ARG((reinterpret_cast<uintptr_t>(ptr) & 15) == 0);

    ProfScope prof(kProfAttnBwd, 10.0 * a->B * a->H * (double)a->Sq * a->Sk * a->D, stream);
    {
        const int64_t n = (int64_t)a->B * a->Sq * a->H * 4;      // a quad of lanes per row
        attn_delta_kernel<<<(int)((n + 255) / 256), 256, 0, stream>>>(
            reinterpret_cast<const __nv_bfloat16*>(a->o), a->o_bs, a->o_rs, a->o_hs,
            reinterpret_cast<const __nv_bfloat16*>(a->dout), a->do_bs, a->do_rs, a->do_hs, a->delta, a->B, a->H, a->Sq, a->D);
        MICO_CHECK_CUDA(cudaGetLastError());
    }
    CUtensorMap tq, tk, tv, tdo, tqx, tkx, tvx, tdox;
    int rc;
    if ((rc = make_attn_tmap(&tq, a->q, a->D, a->H, a->Sq, a->B, a->q_bs, a->q_rs, a->q_hs))) return rc;
    MICO_CHECK_ARG(a->kv_index == nullptr || (a->n_kv > 0 && a->grp_ptr && a->grp_list));
    const int nkvb = a->kv_index ? a->n_kv : a->B;
    if ((rc = make_attn_tmap(&tk, a->k, a->D, a->H, a->Sk, nkvb, a->k_bs, a->k_rs, a->k_hs))) return rc;
    if ((rc = make_attn_tmap(&tv, a->v, a->D, a->H, a->Sk, nkvb, a->v_bs, a->v_rs, a->v_hs))) return rc;
    if ((rc = make_attn_tmap(&tdo, a->dout, a->D, a->H, a->Sq, a->B, a->do_bs, a->do_rs, a->do_hs))) return rc;
    if ((rc = make_attn_tmap(&tqx, a->q, a->D, a->H, a->Sq, a->B, a->q_bs, a->q_rs, a->q_hs, kExtRows))) return rc;
    if ((rc = make_attn_tmap(&tkx, a->k, a->D, a->H, a->Sk, nkvb, a->k_bs, a->k_rs, a->k_hs, kExtRows))) return rc;
    if ((rc = make_attn_tmap(&tvx, a->v, a->D, a->H, a->Sk, nkvb, a->v_bs, a->v_rs, a->v_hs, kExtRows))) return rc;
    if ((rc = make_attn_tmap(&tdox, a->dout, a->D, a->H, a->Sq, a->B, a->do_bs, a->do_rs, a->do_hs, kExtRows))) return rc;
    AttnBwdParams p;
    p.B = a->B; p.H = a->H; p.Sq = a->Sq; p.Sk = a->Sk; p.D = a->D; p.scale = a->scale;
    p.mask = a->mask; p.mask_bs = a->mask_bs; p.mask_qs = a->mask_qs; p.mask_hs = a->mask_hs; p.mask_bmod = a->mask_bmod;
    p.drop.p = a->dropout_p; p.drop.inv_keep = a->dropout_p < 1.f ? 1.f / (1.f - a->dropout_p) : 0.f; p.drop.seed = a->dropout_seed;
    p.lse = a->lse; p.delta = a->delta;
    p.dq = reinterpret_cast<__nv_bfloat16*>(a->dq); p.dq_bs = a->dq_bs; p.dq_rs = a->dq_rs; p.dq_hs = a->dq_hs;
    p.dk = reinterpret_cast<__nv_bfloat16*>(a->dk); p.dk_bs = a->dk_bs; p.dk_rs = a->dk_rs; p.dk_hs = a->dk_hs;
    p.dv = reinterpret_cast<__nv_bfloat16*>(a->dv); p.dv_bs = a->dv_bs; p.dv_rs = a->dv_rs; p.dv_hs = a->dv_hs;
    const int hd_pad = (a->D + 15) & ~15;
    const int work_q = a->B * a->H * m_tiles(a->Sq);
    const int work_k = nkvb * a->H * m_tiles_nt(a->Sk, a->kv_index != nullptr);
    p.kv_index = a->kv_index; p.n_kv = a->n_kv; p.grp_ptr = a->grp_ptr; p.grp_list = a->grp_list;
    auto launch = [&](auto kq, auto kkv) -> int {
        MICO_CHECK_CUDA(cudaFuncSetAttribute(kq, cudaFuncAttributeMaxDynamicSharedMemorySize, DqSmem::TOTAL));
        MICO_CHECK_CUDA(cudaFuncSetAttribute(kkv, cudaFuncAttributeMaxDynamicSharedMemorySize, DkvSmem::TOTAL));
        // remainder rows (SIMT) run on the side stream next to the tile kernels; they need delta, which is already queued
        const bool tail = m_tail_rows(a->Sq) || (m_tail_rows(a->Sk) && a->kv_index == nullptr);
        cudaStream_t side = tail ? side_fork(stream) : nullptr;
        int rc_tail = MICO_OK;
        if (hd_pad == 96 && dq_pipelined_enabled()) {
            // ViT-g (d = 88): double-buffered score tiles + deferred read-out, 105 -> 85 us per layer at bs 64 (ncu, one box)
            const int rc_dq = attention_dq_pipelined(a, stream);
            if (rc_dq) return rc_dq;
        } else {
            kq<<<work_q < num_sms() ? work_q : num_sms(), kBwdThreads, DqSmem::TOTAL, stream>>>(tq, tk, tv, tdo, tkx, tvx, p);
            MICO_CHECK_CUDA(cudaGetLastError());
        }
        // queued after the first tile kernel so that the persistent CTAs get their SMs first; the small tail blocks then fill
        // the shared memory / thread slots the tile kernels leave free
        if (tail && side) rc_tail = attention_tail_bwd(a, side);
        kkv<<<work_k < num_sms() ? work_k : num_sms(), kBwdThreads, DkvSmem::TOTAL, stream>>>(tq, tk, tv, tdo, tqx, tdox, p);
        MICO_CHECK_CUDA(cudaGetLastError());
        count_launch(3);
        if (tail && !side) return attention_tail_bwd(a, stream);
        if (side) side_join(stream);
        return rc_tail;
    };
    const bool plain = a->mask == nullptr && a->dropout_p == 0.0f;
    const bool drop_only = a->mask == nullptr && a->dropout_p > 0.0f && drop_only_enabled();
    switch (hd_pad) {
        case 32: return plain ? launch(attn_dq_kernel<32, true>, attn_dkv_kernel<32, true>)
                              : launch(attn_dq_kernel<32, false>, attn_dkv_kernel<32, false>);
        case 64: return plain ? launch(attn_dq_kernel<64, 1>, attn_dkv_kernel<64, 1>)
                     : drop_only ? launch(attn_dq_kernel<64, 2>, attn_dkv_kernel<64, 2>)
                                 : launch(attn_dq_kernel<64, 0>, attn_dkv_kernel<64, 0>);
        case 96: return plain ? launch(attn_dq_kernel<96, true>, attn_dkv_kernel<96, true>)
                              : launch(attn_dq_kernel<96, false>, attn_dkv_kernel<96, false>);
        default:
            // head_dim 128 would need 2 x 144 + 2 x 128 TMEM columns in attn_dkv_kernel; no tower on the MiCo path
            // has it (ViT-g 88, BERT / CLIP 64, Swin 32)
            set_last_error(__FILE__, __LINE__, "attention backward: head_dim must pad to 32, 64 or 96");
            return MICO_ERR_UNSUPPORTED;
    }
}
