// mico_b200 -- K4 backward: fused attention gradients (sm_100a, tcgen05), no score matrix in HBM.
//
// Autograd of eva_vit_model.py:340-361 / bert.py:233-277.  With P = softmax(scale*QK^T + mask) recomputed
// from Q, K and the saved log-sum-exp, and delta_i = sum_d dO_id O_id:
//     dV = P^T dO          dP = dO V^T          dS = scale * P o (dP - delta)
//     dQ = dS K            dK = dS^T Q
// Three launches:
//   attn_delta_kernel   delta[b,h,i]                                     (HBM-bound row reduction)
//   attn_dq_kernel      CTA per (b,h,q-tile), loops over KV tiles:  S, dP -> dS (smem) -> dQ += dS K
//   attn_dkv_kernel     CTA per (b,h,kv-tile), loops over Q tiles:  S^T, dP^T -> P^T, dS^T (smem)
//                                                                  -> dV += P^T dO, dK += dS^T Q
// Both MMA kernels recompute S (7 GEMMs instead of 5) so that no atomics are needed and the result is
// deterministic.  The dS / P^T tiles written to shared memory serve as K-major A operands; K, Q and dO tiles
// are read K-major for the score GEMMs and MN-major (same bytes) for the gradient GEMMs.
#include "attn_common.cuh"

namespace mico {
namespace {

struct AttnBwdParams {
    int B, H, Sq, Sk, D;
    float scale;
    const float* mask;
    int64_t mask_bs, mask_qs;
    const float* lse;     // [B,H,Sq]
    const float* delta;   // [B,H,Sq]
    __nv_bfloat16* dq; int64_t dq_bs, dq_rs, dq_hs;
    __nv_bfloat16* dk; int64_t dk_bs, dk_rs, dk_hs;
    __nv_bfloat16* dv; int64_t dv_bs, dv_rs, dv_hs;
};

// ------------------------------------------------------------------------------------------------ delta
__global__ void attn_delta_kernel(const __nv_bfloat16* __restrict__ o, int64_t o_bs, int64_t o_rs, int64_t o_hs,
                                  const __nv_bfloat16* __restrict__ d_o, int64_t do_bs, int64_t do_rs, int64_t do_hs,
                                  float* __restrict__ delta, int B, int H, int S, int D) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;   // over (b, i, h), h fastest
    if (idx >= (int64_t)B * S * H) return;
    const int h = (int)(idx % H);
    const int i = (int)((idx / H) % S);
    const int b = (int)(idx / ((int64_t)H * S));
    const uint4* po = reinterpret_cast<const uint4*>(o + b * o_bs + i * o_rs + h * o_hs);
    const uint4* pd = reinterpret_cast<const uint4*>(d_o + b * do_bs + i * do_rs + h * do_hs);
    float acc = 0.f;
    for (int c = 0; c < D / 8; ++c) {
        const uint4 a = po[c], g = pd[c];
        acc += bf16_lo(a.x) * bf16_lo(g.x) + bf16_hi(a.x) * bf16_hi(g.x) + bf16_lo(a.y) * bf16_lo(g.y) +
               bf16_hi(a.y) * bf16_hi(g.y) + bf16_lo(a.z) * bf16_lo(g.z) + bf16_hi(a.z) * bf16_hi(g.z) +
               bf16_lo(a.w) * bf16_lo(g.w) + bf16_hi(a.w) * bf16_hi(g.w);
    }
    delta[((int64_t)b * H + h) * S + i] = acc;
}

// ------------------------------------------------------------------------------------------------ dQ
struct DqSmem {
    static constexpr int Q = 0;
    static constexpr int DO = 2 * kAtomBytes;
    static constexpr int K0 = 4 * kAtomBytes;               // 2 stages x 2 atoms
    static constexpr int V0 = 8 * kAtomBytes;
    static constexpr int DS = 12 * kAtomBytes;
    static constexpr int BARS = 14 * kAtomBytes;
    static constexpr int TOTAL = BARS + 256 + 1024;
};

template <int HD_PAD>
__global__ void __launch_bounds__(kAttThreads, 1)
attn_dq_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
               const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmDO, AttnBwdParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + DqSmem::BARS);
    uint64_t* q_full = bars + 0;     // Q + dO tiles landed
    uint64_t* q_empty = bars + 1;    // all MMAs of the work item done
    uint64_t* kv_full = bars + 2;    // [2]
    uint64_t* kv_empty = bars + 4;   // [2]
    uint64_t* sdp_full = bars + 6;   // S and dP in TMEM
    uint64_t* sdp_empty = bars + 7;  // softmax finished reading them
    uint64_t* ds_full = bars + 8;    // dS tile in smem
    uint64_t* ds_empty = bars + 9;   // dQ MMA finished reading it
    uint64_t* dq_full = bars + 10;   // dQ accumulator final
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 12);

    const int warp = threadIdx.x >> 5;
    const int nqt = (p.Sq + kTile - 1) / kTile;
    const int nkv = (p.Sk + kTile - 1) / kTile;
    const int num_work = p.B * p.H * nqt;
    constexpr int kAtoms = (HD_PAD + 63) / 64;
    constexpr uint32_t kTileBytes = kAtoms * kAtomBytes;

    if (warp == 4) {
        if (elect_one()) {
            tma_prefetch_desc(&tmQ); tma_prefetch_desc(&tmK); tma_prefetch_desc(&tmV); tma_prefetch_desc(&tmDO);
        }
    } else if (warp == 5) {
        if (elect_one()) {
            mbar_init(q_full, 1); mbar_init(q_empty, 1);
            for (int i = 0; i < 2; ++i) { mbar_init(&kv_full[i], 1); mbar_init(&kv_empty[i], 1); }
            mbar_init(sdp_full, 1); mbar_init(sdp_empty, 4);
            mbar_init(ds_full, 4); mbar_init(ds_empty, 1);
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
    const uint32_t tmem_S = tmem_base, tmem_dP = tmem_base + 128, tmem_dQ = tmem_base + 256;

    if (warp == 4) {
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
                for (int j = 0; j < nkv; ++j, ++kvcount) {
                    const int s = kvcount & 1;
                    mbar_wait(&kv_empty[s], ((kvcount >> 1) & 1) ^ 1);
                    mbar_arrive_expect_tx(&kv_full[s], 2 * kTileBytes);
#pragma unroll
                    for (int a = 0; a < kAtoms; ++a) {
                        tma_load_4d(smem + DqSmem::K0 + (s * 2 + a) * kAtomBytes, &tmK, &kv_full[s], a * 64, h, j * kTile, b);
                        tma_load_4d(smem + DqSmem::V0 + (s * 2 + a) * kAtomBytes, &tmV, &kv_full[s], a * 64, h, j * kTile, b);
                    }
                }
            }
        }
    } else if (warp == 5) {
        if (elect_one()) {
            uint32_t wcount = 0, kvcount = 0, tcount = 0;   // tcount: kv tiles processed (sdp / ds barrier phases)
            const uint32_t sQ = smem_u32(smem + DqSmem::Q), sDO = smem_u32(smem + DqSmem::DO);
            const uint32_t sDS = smem_u32(smem + DqSmem::DS);
            constexpr uint32_t idesc_dq = umma_idesc_bf16(HD_PAD, false, true);
            for (int w = blockIdx.x; w < num_work; w += gridDim.x, ++wcount) {
                mbar_wait(q_full, wcount & 1);
                tc_fence_after();
                auto issue_scores = [&](int j, uint32_t kvc, uint32_t tc) {
                    const int s = kvc & 1;
                    mbar_wait(&kv_full[s], (kvc >> 1) & 1);
                    mbar_wait(sdp_empty, (tc & 1) ^ 1);
                    tc_fence_after();
                    const int valid = min(kTile, p.Sk - j * kTile);
                    const uint32_t idesc = umma_idesc_bf16(max(16, (valid + 15) & ~15), false, false);
                    const uint32_t sK = smem_u32(smem + DqSmem::K0 + s * 2 * kAtomBytes);
                    const uint32_t sV = smem_u32(smem + DqSmem::V0 + s * 2 * kAtomBytes);
#pragma unroll
                    for (int k = 0; k < HD_PAD / 16; ++k) {
                        const uint32_t off = (k >> 2) * kAtomBytes + (k & 3) * 32;
                        umma_bf16_ss(tmem_S, umma_smem_desc_sw128(sQ + off, 16, 1024),
                                     umma_smem_desc_sw128(sK + off, 16, 1024), idesc, k != 0);
                    }
#pragma unroll
                    for (int k = 0; k < HD_PAD / 16; ++k) {
                        const uint32_t off = (k >> 2) * kAtomBytes + (k & 3) * 32;
                        umma_bf16_ss(tmem_dP, umma_smem_desc_sw128(sDO + off, 16, 1024),
                                     umma_smem_desc_sw128(sV + off, 16, 1024), idesc, k != 0);
                    }
                    umma_commit(sdp_full);
                };
                issue_scores(0, kvcount, tcount);
                for (int j = 0; j < nkv; ++j) {
                    if (j + 1 < nkv) issue_scores(j + 1, kvcount + 1, tcount + 1);
                    const int s = kvcount & 1;
                    mbar_wait(ds_full, tcount & 1);
                    tc_fence_after();
                    const int valid = min(kTile, p.Sk - j * kTile);
                    const int ksteps = (valid + 15) >> 4;
                    const uint32_t sK = smem_u32(smem + DqSmem::K0 + s * 2 * kAtomBytes);
                    for (int k = 0; k < ksteps; ++k) {
                        const uint32_t aoff = (k >> 2) * kAtomBytes + (k & 3) * 32;
                        umma_bf16_ss(tmem_dQ, umma_smem_desc_sw128(sDS + aoff, 16, 1024),
                                     umma_smem_desc_sw128(sK + k * 2048, kAtomBytes, 1024), idesc_dq, (j | k) != 0);
                    }
                    umma_commit(&kv_empty[s]);
                    umma_commit(ds_empty);
                    ++kvcount; ++tcount;
                }
                umma_commit(dq_full);
                umma_commit(q_empty);
            }
        }
    } else {
        const int r = threadIdx.x;
        const uint32_t lane_off = (uint32_t)(warp * 32) << 16;
        const float sc2 = p.scale * kLog2e;
        uint32_t tcount = 0, wcount = 0;
        uint8_t* sDS = smem + DqSmem::DS;
        for (int w = blockIdx.x; w < num_work; w += gridDim.x, ++wcount) {
            const int qt = w % nqt, bh = w / nqt;
            const int h = bh % p.H, b = bh / p.H;
            const int qi = qt * kTile + r;
            const bool row_ok = qi < p.Sq;
            const int64_t stat = ((int64_t)b * p.H + h) * p.Sq + (row_ok ? qi : 0);
            const float lse2 = row_ok ? p.lse[stat] * kLog2e : 0.f;
            const float dlt = row_ok ? p.delta[stat] : 0.f;
            const float* mrow = p.mask ? p.mask + (int64_t)b * p.mask_bs + (int64_t)(row_ok ? qi : 0) * p.mask_qs : nullptr;
            for (int j = 0; j < nkv; ++j, ++tcount) {
                const int valid = min(kTile, p.Sk - j * kTile);
                const int nch = (valid + 31) >> 5;
                mbar_wait(sdp_full, tcount & 1);
                tc_fence_after();
                mbar_wait(ds_empty, (tcount & 1) ^ 1);   // previous dQ MMA no longer reads the dS tile
                for (int c = 0; c < nch; ++c) {
                    uint32_t sv[32], dv[32];
                    tmem_ld_x32(tmem_S + lane_off + c * 32, sv);
                    tmem_ld_x32(tmem_dP + lane_off + c * 32, dv);
                    tmem_ld_wait();
                    float ds[32];
#pragma unroll
                    for (int i = 0; i < 32; ++i) {
                        const int col = c * 32 + i;
                        float s = __uint_as_float(sv[i]) * sc2 - lse2;
                        if (mrow) s += (col < valid ? mrow[j * kTile + col] : 0.f) * kLog2e;
                        const float pr = exp2f(s);
                        ds[i] = (col < valid && row_ok) ? pr * (__uint_as_float(dv[i]) - dlt) * p.scale : 0.f;
                    }
                    store_tile_chunk32(sDS, r, c * 32, ds);
                }
                tc_fence_before();
                fence_proxy_async_smem();
                __syncwarp();
                if (lane_id() == 0) {
                    mbar_arrive(sdp_empty);
                    mbar_arrive(ds_full);
                }
            }
            mbar_wait(dq_full, wcount & 1);
            tc_fence_after();
            float acc[HD_PAD];
            tmem_load_row<HD_PAD, false>(tmem_dQ + lane_off, acc);
            tc_fence_before();
            if (row_ok)
                store_row_bf16<HD_PAD>(p.dq + (int64_t)b * p.dq_bs + (int64_t)qi * p.dq_rs + (int64_t)h * p.dq_hs, acc, p.D, 1.0f);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 5) {
        tc_fence_after();
        tmem_dealloc<512>(tmem_base);
    }
}

// ------------------------------------------------------------------------------------------------ dK, dV
struct DkvSmem {
    static constexpr int K = 0;
    static constexpr int V = 2 * kAtomBytes;
    static constexpr int Q = 4 * kAtomBytes;
    static constexpr int DO = 6 * kAtomBytes;
    static constexpr int PT = 8 * kAtomBytes;
    static constexpr int DST = 10 * kAtomBytes;
    static constexpr int STATS = 12 * kAtomBytes;             // lse2[128], delta[128]
    static constexpr int BARS = STATS + 1024;
    static constexpr int TOTAL = BARS + 256 + 1024;
};

template <int HD_PAD>
__global__ void __launch_bounds__(kAttThreads, 1)
attn_dkv_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmDO, AttnBwdParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + DkvSmem::BARS);
    uint64_t* kv_full = bars + 0;
    uint64_t* kv_empty = bars + 1;
    uint64_t* qdo_full = bars + 2;
    uint64_t* qdo_empty = bars + 3;
    uint64_t* st_full = bars + 4;
    uint64_t* st_empty = bars + 5;
    uint64_t* pds_full = bars + 6;
    uint64_t* pds_empty = bars + 7;
    uint64_t* acc_full = bars + 8;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 10);
    float* s_lse2 = reinterpret_cast<float*>(smem + DkvSmem::STATS);
    float* s_delta = s_lse2 + kTile;

    const int warp = threadIdx.x >> 5;
    const int nqt = (p.Sq + kTile - 1) / kTile;
    const int nkv = (p.Sk + kTile - 1) / kTile;
    const int num_work = p.B * p.H * nkv;
    constexpr int kAtoms = (HD_PAD + 63) / 64;
    constexpr uint32_t kTileBytes = kAtoms * kAtomBytes;

    if (warp == 4) {
        if (elect_one()) {
            tma_prefetch_desc(&tmQ); tma_prefetch_desc(&tmK); tma_prefetch_desc(&tmV); tma_prefetch_desc(&tmDO);
        }
    } else if (warp == 5) {
        if (elect_one()) {
            mbar_init(kv_full, 1); mbar_init(kv_empty, 1);
            mbar_init(qdo_full, 1); mbar_init(qdo_empty, 1);
            mbar_init(st_full, 1); mbar_init(st_empty, 4);
            mbar_init(pds_full, 4); mbar_init(pds_empty, 1);
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
    const uint32_t tmem_ST = tmem_base, tmem_dPT = tmem_base + 128, tmem_dV = tmem_base + 256, tmem_dK = tmem_base + 384;

    if (warp == 4) {
        if (elect_one()) {
            uint32_t wcount = 0, qcount = 0;
            for (int w = blockIdx.x; w < num_work; w += gridDim.x, ++wcount) {
                const int jt = w % nkv, bh = w / nkv;
                const int h = bh % p.H, b = bh / p.H;
                mbar_wait(kv_empty, (wcount & 1) ^ 1);
                mbar_arrive_expect_tx(kv_full, 2 * kTileBytes);
#pragma unroll
                for (int a = 0; a < kAtoms; ++a) {
                    tma_load_4d(smem + DkvSmem::K + a * kAtomBytes, &tmK, kv_full, a * 64, h, jt * kTile, b);
                    tma_load_4d(smem + DkvSmem::V + a * kAtomBytes, &tmV, kv_full, a * 64, h, jt * kTile, b);
                }
                for (int i = 0; i < nqt; ++i, ++qcount) {
                    mbar_wait(qdo_empty, (qcount & 1) ^ 1);
                    mbar_arrive_expect_tx(qdo_full, 2 * kTileBytes);
#pragma unroll
                    for (int a = 0; a < kAtoms; ++a) {
                        tma_load_4d(smem + DkvSmem::Q + a * kAtomBytes, &tmQ, qdo_full, a * 64, h, i * kTile, b);
                        tma_load_4d(smem + DkvSmem::DO + a * kAtomBytes, &tmDO, qdo_full, a * 64, h, i * kTile, b);
                    }
                }
            }
        }
    } else if (warp == 5) {
        if (elect_one()) {
            uint32_t wcount = 0, qcount = 0;
            const uint32_t sK = smem_u32(smem + DkvSmem::K), sV = smem_u32(smem + DkvSmem::V);
            const uint32_t sQ = smem_u32(smem + DkvSmem::Q), sDO = smem_u32(smem + DkvSmem::DO);
            const uint32_t sPT = smem_u32(smem + DkvSmem::PT), sDST = smem_u32(smem + DkvSmem::DST);
            constexpr uint32_t idesc_g = umma_idesc_bf16(HD_PAD, false, true);
            for (int w = blockIdx.x; w < num_work; w += gridDim.x, ++wcount) {
                mbar_wait(kv_full, wcount & 1);
                tc_fence_after();
                for (int i = 0; i < nqt; ++i, ++qcount) {
                    mbar_wait(qdo_full, qcount & 1);
                    mbar_wait(st_empty, (qcount & 1) ^ 1);
                    tc_fence_after();
                    const int validq = min(kTile, p.Sq - i * kTile);
                    const uint32_t idesc = umma_idesc_bf16(max(16, (validq + 15) & ~15), false, false);
#pragma unroll
                    for (int k = 0; k < HD_PAD / 16; ++k) {
                        const uint32_t off = (k >> 2) * kAtomBytes + (k & 3) * 32;
                        umma_bf16_ss(tmem_ST, umma_smem_desc_sw128(sK + off, 16, 1024),
                                     umma_smem_desc_sw128(sQ + off, 16, 1024), idesc, k != 0);
                    }
#pragma unroll
                    for (int k = 0; k < HD_PAD / 16; ++k) {
                        const uint32_t off = (k >> 2) * kAtomBytes + (k & 3) * 32;
                        umma_bf16_ss(tmem_dPT, umma_smem_desc_sw128(sV + off, 16, 1024),
                                     umma_smem_desc_sw128(sDO + off, 16, 1024), idesc, k != 0);
                    }
                    umma_commit(st_full);
                    mbar_wait(pds_full, qcount & 1);
                    tc_fence_after();
                    const int ksteps = (validq + 15) >> 4;
                    for (int k = 0; k < ksteps; ++k) {
                        const uint32_t aoff = (k >> 2) * kAtomBytes + (k & 3) * 32;
                        umma_bf16_ss(tmem_dV, umma_smem_desc_sw128(sPT + aoff, 16, 1024),
                                     umma_smem_desc_sw128(sDO + k * 2048, kAtomBytes, 1024), idesc_g, (i | k) != 0);
                    }
                    for (int k = 0; k < ksteps; ++k) {
                        const uint32_t aoff = (k >> 2) * kAtomBytes + (k & 3) * 32;
                        umma_bf16_ss(tmem_dK, umma_smem_desc_sw128(sDST + aoff, 16, 1024),
                                     umma_smem_desc_sw128(sQ + k * 2048, kAtomBytes, 1024), idesc_g, (i | k) != 0);
                    }
                    umma_commit(qdo_empty);
                    umma_commit(pds_empty);
                }
                umma_commit(acc_full);
                umma_commit(kv_empty);
            }
        }
    } else {
        const int r = threadIdx.x;                    // key row within the tile
        const uint32_t lane_off = (uint32_t)(warp * 32) << 16;
        const float sc2 = p.scale * kLog2e;
        uint32_t qcount = 0, wcount = 0;
        for (int w = blockIdx.x; w < num_work; w += gridDim.x, ++wcount) {
            const int jt = w % nkv, bh = w / nkv;
            const int h = bh % p.H, b = bh / p.H;
            const int kj = jt * kTile + r;
            const bool key_ok = kj < p.Sk;
            const int64_t stat0 = ((int64_t)b * p.H + h) * p.Sq;
            for (int i = 0; i < nqt; ++i, ++qcount) {
                const int validq = min(kTile, p.Sq - i * kTile);
                const int nch = (validq + 31) >> 5;
                // stage this q tile's lse / delta (previous tile's readers are past their last use: they have
                // all arrived on pds_full, and we only get here after waiting st_full of this tile)
                softmax_group_sync();
                s_lse2[r] = (r < validq) ? p.lse[stat0 + i * kTile + r] * kLog2e : 0.f;
                s_delta[r] = (r < validq) ? p.delta[stat0 + i * kTile + r] : 0.f;
                softmax_group_sync();
                mbar_wait(st_full, qcount & 1);
                tc_fence_after();
                mbar_wait(pds_empty, (qcount & 1) ^ 1);
                for (int c = 0; c < nch; ++c) {
                    uint32_t sv[32], dv[32];
                    tmem_ld_x32(tmem_ST + lane_off + c * 32, sv);
                    tmem_ld_x32(tmem_dPT + lane_off + c * 32, dv);
                    tmem_ld_wait();
                    float pt[32], dst[32];
#pragma unroll
                    for (int q = 0; q < 32; ++q) {
                        const int col = c * 32 + q;                 // query within the tile
                        float s = __uint_as_float(sv[q]) * sc2 - s_lse2[col];
                        if (p.mask && key_ok && col < validq)
                            s += p.mask[(int64_t)b * p.mask_bs + (int64_t)(i * kTile + col) * p.mask_qs + kj] * kLog2e;
                        const bool ok = key_ok && col < validq;
                        const float pr = ok ? exp2f(s) : 0.f;
                        pt[q] = pr;
                        dst[q] = ok ? pr * (__uint_as_float(dv[q]) - s_delta[col]) * p.scale : 0.f;
                    }
                    store_tile_chunk32(smem + DkvSmem::PT, r, c * 32, pt);
                    store_tile_chunk32(smem + DkvSmem::DST, r, c * 32, dst);
                }
                tc_fence_before();
                fence_proxy_async_smem();
                __syncwarp();
                if (lane_id() == 0) {
                    mbar_arrive(st_empty);
                    mbar_arrive(pds_full);
                }
            }
            mbar_wait(acc_full, wcount & 1);
            tc_fence_after();
            {
                float acc[HD_PAD];
                tmem_load_row<HD_PAD, false>(tmem_dV + lane_off, acc);
                if (key_ok)
                    store_row_bf16<HD_PAD>(p.dv + (int64_t)b * p.dv_bs + (int64_t)kj * p.dv_rs + (int64_t)h * p.dv_hs, acc, p.D, 1.0f);
                tmem_load_row<HD_PAD, false>(tmem_dK + lane_off, acc);
                if (key_ok)
                    store_row_bf16<HD_PAD>(p.dk + (int64_t)b * p.dk_bs + (int64_t)kj * p.dk_rs + (int64_t)h * p.dk_hs, acc, p.D, 1.0f);
            }
            tc_fence_before();
            // the next work item's first MMAs overwrite dV/dK: they are ordered after our reads through
            // st_empty (arrived only after these loads in program order on the next tile).
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 5) {
        tc_fence_after();
        tmem_dealloc<512>(tmem_base);
    }
}

}  // namespace
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
        const int64_t n = (int64_t)a->B * a->Sq * a->H;
        attn_delta_kernel<<<(int)((n + 255) / 256), 256, 0, stream>>>(
            reinterpret_cast<const __nv_bfloat16*>(a->o), a->o_bs, a->o_rs, a->o_hs,
            reinterpret_cast<const __nv_bfloat16*>(a->dout), a->do_bs, a->do_rs, a->do_hs, a->delta, a->B, a->H, a->Sq, a->D);
        MICO_CHECK_CUDA(cudaGetLastError());
    }
    CUtensorMap tq, tk, tv, tdo;
    int rc;
    if ((rc = make_attn_tmap(&tq, a->q, a->D, a->H, a->Sq, a->B, a->q_bs, a->q_rs, a->q_hs))) return rc;
    if ((rc = make_attn_tmap(&tk, a->k, a->D, a->H, a->Sk, a->B, a->k_bs, a->k_rs, a->k_hs))) return rc;
    if ((rc = make_attn_tmap(&tv, a->v, a->D, a->H, a->Sk, a->B, a->v_bs, a->v_rs, a->v_hs))) return rc;
    if ((rc = make_attn_tmap(&tdo, a->dout, a->D, a->H, a->Sq, a->B, a->do_bs, a->do_rs, a->do_hs))) return rc;
    AttnBwdParams p;
    p.B = a->B; p.H = a->H; p.Sq = a->Sq; p.Sk = a->Sk; p.D = a->D; p.scale = a->scale;
    p.mask = a->mask; p.mask_bs = a->mask_bs; p.mask_qs = a->mask_qs;
    p.lse = a->lse; p.delta = a->delta;
    p.dq = reinterpret_cast<__nv_bfloat16*>(a->dq); p.dq_bs = a->dq_bs; p.dq_rs = a->dq_rs; p.dq_hs = a->dq_hs;
    p.dk = reinterpret_cast<__nv_bfloat16*>(a->dk); p.dk_bs = a->dk_bs; p.dk_rs = a->dk_rs; p.dk_hs = a->dk_hs;
    p.dv = reinterpret_cast<__nv_bfloat16*>(a->dv); p.dv_bs = a->dv_bs; p.dv_rs = a->dv_rs; p.dv_hs = a->dv_hs;
    const int hd_pad = (a->D + 15) & ~15;
    const int work_q = a->B * a->H * ceil_div(a->Sq, kTile);
    const int work_k = a->B * a->H * ceil_div(a->Sk, kTile);
    auto launch = [&](auto kq, auto kkv) -> int {
        MICO_CHECK_CUDA(cudaFuncSetAttribute(kq, cudaFuncAttributeMaxDynamicSharedMemorySize, DqSmem::TOTAL));
        MICO_CHECK_CUDA(cudaFuncSetAttribute(kkv, cudaFuncAttributeMaxDynamicSharedMemorySize, DkvSmem::TOTAL));
        kq<<<work_q < num_sms() ? work_q : num_sms(), kAttThreads, DqSmem::TOTAL, stream>>>(tq, tk, tv, tdo, p);
        MICO_CHECK_CUDA(cudaGetLastError());
        kkv<<<work_k < num_sms() ? work_k : num_sms(), kAttThreads, DkvSmem::TOTAL, stream>>>(tq, tk, tv, tdo, p);
        MICO_CHECK_CUDA(cudaGetLastError());
        count_launch(3);
        return MICO_OK;
    };
    switch (hd_pad) {
        case 32: return launch(attn_dq_kernel<32>, attn_dkv_kernel<32>);
        case 64: return launch(attn_dq_kernel<64>, attn_dkv_kernel<64>);
        case 96: return launch(attn_dq_kernel<96>, attn_dkv_kernel<96>);
        case 128: return launch(attn_dq_kernel<128>, attn_dkv_kernel<128>);
        default:
            set_last_error(__FILE__, __LINE__, "head_dim must pad to 32, 64, 96 or 128");
            return MICO_ERR_UNSUPPORTED;
    }
}
