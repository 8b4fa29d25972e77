// mico_b200 -- K4 backward, dQ kernel with double-buffered score tiles (head_dim 81..96: the ViT-g tower).
//
// Same math as attn_dq_kernel of attention_bwd.cu (S, dP -> dS in tensor memory -> dQ += dS K); what differs is the
// schedule, see "pipelining" below.  The matching restructure of the dK/dV kernel (80-query streamed tiles, deferred
// read-out, per-warp statistics) was written and measured too: 126 -> 136 us per layer at bs 64, so it was dropped and
// attention_bwd.cu keeps the single-buffer dK/dV kernel ("measured and rejected").
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

// ------------------------------------------------------------------------------------------------ pipelining (round 2)
// The kernels of attention_bwd.cu hold ONE score buffer in tensor memory, so every (work item, streamed tile) step runs
// score MMA -> softmax -> gradient MMA strictly in series: ncu's source view put 33 % of the dQ kernel's stall samples
// on the softmax warps' wait for the score MMAs and another 10 % on the wait for the last dQ MMA (tensor pipe 20 %
// active).  Here the kernel keeps TWO score buffers: keys stream in tiles of 96 so that 2 x (S + dP) and the
// accumulator fit the 512 columns, the issuer runs the score MMAs TWO steps ahead of the dQ MMAs -- while the softmax
// warps work on step t the tensor pipe computes the scores of step t+1 and the gradient product of step t-1 -- and a work
// item's accumulator is read out only after the first softmax step of the next item (deferred read-out).  Streamed
// tiles are plain TMA boxes (rows past the end of the sequence are zero-filled by the TMA unit: their scores are 0 and
// their K rows contribute nothing).  ncu, ViT-g layer at bs 64: 105 -> 85 us, tensor pipe 20 -> 25 %; what is left is the
// Q / dO tile of the next work item (one shared-memory stage: a second one does not fit next to three K / V stages).
template <int W>
__host__ __device__ inline int s_tiles(int S) { return (S + W - 1) / W; }
template <int W>
__host__ __device__ inline int s_valid(int S, int j, int n) { return j < n - 1 ? W : S - W * (n - 1); }
__device__ __forceinline__ int split_a(int n16) { return ((n16 / 16 + 1) / 2) * 16; }   // columns owned by half 0

// this warp's 32 lanes x NC (16 / 32 / 48) consecutive fp32 columns -> registers, one wait
template <int NC>
__device__ __forceinline__ void tmem_load_cols(uint32_t taddr, float (&acc)[NC]) {
    static_assert(NC == 16 || NC == 32 || NC == 48, "column count");
    uint32_t v[NC];
    if constexpr (NC >= 32) {
        uint32_t (&c0)[32] = *reinterpret_cast<uint32_t (*)[32]>(&v[0]);
        tmem_ld_x32(taddr, c0);
    }
    if constexpr (NC % 32 != 0) {
        uint32_t (&c1)[16] = *reinterpret_cast<uint32_t (*)[16]>(&v[(NC / 32) * 32]);
        tmem_ld_x16(taddr + (NC / 32) * 32, c1);
    }
    tmem_ld_wait();
#pragma unroll
    for (int i = 0; i < NC; ++i) acc[i] = __uint_as_float(v[i]);
}
// NC columns [col0, col0 + NC) of a bf16 row (16-byte chunks; only columns < D exist)
template <int NC>
__device__ __forceinline__ void store_cols_bf16(__nv_bfloat16* row, int col0, const float (&acc)[NC], int D) {
#pragma unroll
    for (int g = 0; g < NC / 8; ++g) {
        if (col0 + g * 8 < D)
            *reinterpret_cast<uint4*>(row + col0 + g * 8) =
                make_uint4(pack_bf16x2(acc[g * 8 + 0], acc[g * 8 + 1]), pack_bf16x2(acc[g * 8 + 2], acc[g * 8 + 3]),
                           pack_bf16x2(acc[g * 8 + 4], acc[g * 8 + 5]), pack_bf16x2(acc[g * 8 + 6], acc[g * 8 + 7]));
    }
}

// ------------------------------------------------------------------------------------------------ dQ
// Work item = (b, h, 128-row query tile); keys stream in tiles of WK = 96.
// TMEM columns: buffer u in {0, 1}: S [192u, 192u+96)  dP [192u+96, 192u+192);  dQ accumulator [384, 384+HD_PAD).
// dS (packed bf16, the A operand of dQ += dS K) overwrites dP in place: the two warps of a lane quarter own the key
// ranges [0,hA) and [hA,n16) and write their packed output at the start of their own range, behind their reads.
template <int HD_PAD>
struct DqCfg {
    static constexpr int kAtoms = (HD_PAD + 63) / 64;
    static constexpr int WK = 96;
    static constexpr int QST = HD_PAD <= 64 ? 2 : 1;          // Q / dO stages
    static constexpr int KVST = HD_PAD <= 64 ? 4 : 3;         // K / V stages (a stage is held until its dQ MMA is done)
    static constexpr int kTileBytes = kAtoms * kAtomBytes;    // a Q or dO tile
    static constexpr int kAtomN = WK * 128;
    static constexpr int kStageN = kAtoms * kAtomN;           // a K or V tile
    static constexpr int Q0 = 0;
    static constexpr int DO0 = QST * kTileBytes;
    static constexpr int K0 = 2 * QST * kTileBytes;
    static constexpr int V0 = K0 + KVST * kStageN;
    static constexpr int BARS = V0 + KVST * kStageN;
    static constexpr int TOTAL = BARS + 512 + 1024;
    static constexpr uint32_t kBuf = 2 * WK, kAcc = 2 * kBuf;
    static_assert(kAcc + HD_PAD <= 512, "TMEM budget");
    static_assert(TOTAL <= 232448, "shared memory budget");
};

// kPlain: no additive mask and no dropout (every ViT tower): the per-element mask / dropout code is compiled out
template <int HD_PAD, bool kPlain>
__global__ void __launch_bounds__(kBwdThreads, 1)
attn_dq_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
               const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmDO, AttnBwdParams p) {
    using C = DqCfg<HD_PAD>;
    constexpr int kAtoms = C::kAtoms, WK = C::WK, QST = C::QST, KVST = C::KVST;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::BARS);
    uint64_t* q_full = bars + 0;      // [2] Q + dO tiles landed
    uint64_t* q_empty = bars + 2;     // [2] the item's last score MMAs are done
    uint64_t* kv_full = bars + 4;     // [4]
    uint64_t* kv_empty = bars + 8;    // [4] the step's dQ MMA is done
    uint64_t* sdp_full = bars + 12;   // [2] S and dP of the buffer are in TMEM
    uint64_t* ds_full = bars + 14;    // [2] dS operand written (and S / dP of the buffer no longer read)
    uint64_t* dq_full = bars + 16;    // dQ accumulator final
    uint64_t* dq_empty = bars + 17;   // ... and read out by the softmax warps
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 20);

    const int warp = threadIdx.x >> 5;
    const int nqt = m_tiles(p.Sq);
    const int nkv = s_tiles<WK>(p.Sk);
    const int num_work = p.B * p.H * nqt;

    if (warp == 8) {
        if (elect_one()) {
            tma_prefetch_desc(&tmQ); tma_prefetch_desc(&tmK); tma_prefetch_desc(&tmV); tma_prefetch_desc(&tmDO);
        }
    } else if (warp == 9) {
        if (elect_one()) {
            for (int i = 0; i < 2; ++i) {
                mbar_init(&q_full[i], 1); mbar_init(&q_empty[i], 1);
                mbar_init(&sdp_full[i], 1); mbar_init(&ds_full[i], 8);
            }
            for (int i = 0; i < 4; ++i) { mbar_init(&kv_full[i], 1); mbar_init(&kv_empty[i], 1); }
            mbar_init(dq_full, 1);
            mbar_init(dq_empty, 8);
            fence_mbar_init();
        }
        __syncwarp();
        tmem_alloc<512>(tmem_slot);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t tmem_dQ = tmem_base + C::kAcc;

    if (warp == 8) {
        // ------------------------------------------------------------------ TMA producer
        if (elect_one()) {
            uint32_t wc = 0, t = 0;
            for (int w = blockIdx.x; w < num_work; w += gridDim.x, ++wc) {
                const int qt = w % nqt, bh = w / nqt;
                const int h = bh % p.H, b = bh / p.H;
                const int qs = wc % QST;
                mbar_wait_relaxed(&q_empty[qs], ((wc / QST) & 1) ^ 1);
                mbar_arrive_expect_tx(&q_full[qs], 2 * C::kTileBytes);
#pragma unroll
                for (int a = 0; a < kAtoms; ++a) {
                    tma_load_4d(smem + C::Q0 + qs * C::kTileBytes + a * kAtomBytes, &tmQ, &q_full[qs], a * 64, h, qt * kTile, b);
                    tma_load_4d(smem + C::DO0 + qs * C::kTileBytes + a * kAtomBytes, &tmDO, &q_full[qs], a * 64, h, qt * kTile, b);
                }
                if (QST == 1 && w + (int)gridDim.x < num_work) {
                    // single Q / dO stage (head_dim > 64): the next item's tiles can only be loaded once this item's last score
                    // MMAs are done, one streamed tile before they are needed -- pull them into L2 now so that load is short
                    const int w2 = w + gridDim.x, qt2 = w2 % nqt, bh2 = w2 / nqt;
#pragma unroll
                    for (int a = 0; a < kAtoms; ++a) {
                        tma_prefetch_4d(&tmQ, a * 64, bh2 % p.H, qt2 * kTile, bh2 / p.H);
                        tma_prefetch_4d(&tmDO, a * 64, bh2 % p.H, qt2 * kTile, bh2 / p.H);
                    }
                }
                const int kvb = p.kv_index ? __ldg(p.kv_index + b) : b;
                for (int j = 0; j < nkv; ++j, ++t) {
                    const int s = t % KVST;
                    mbar_wait_relaxed(&kv_empty[s], ((t / KVST) & 1) ^ 1);
                    mbar_arrive_expect_tx(&kv_full[s], 2 * C::kStageN);
#pragma unroll
                    for (int a = 0; a < kAtoms; ++a) {
                        tma_load_4d(smem + C::K0 + s * C::kStageN + a * C::kAtomN, &tmK, &kv_full[s], a * 64, h, j * WK, kvb);
                        tma_load_4d(smem + C::V0 + s * C::kStageN + a * C::kAtomN, &tmV, &kv_full[s], a * 64, h, j * WK, kvb);
                    }
                }
            }
        }
    } else if (warp == 9) {
        // ------------------------------------------------------------------ MMA issuer: score MMAs run two steps ahead
        if (elect_one()) {
            struct Cur { int w, j; uint32_t wc, t; };
            auto adv = [&](Cur& c) {
                ++c.t;
                if (++c.j == nkv) { c.j = 0; c.w += gridDim.x; ++c.wc; }
            };
            const uint32_t sbase = smem_u32(smem);
            constexpr uint32_t idesc_dq = umma_idesc_bf16(HD_PAD, false, true);
            auto scores = [&](const Cur& c) {
                const int qs = c.wc % QST, s = c.t % KVST;
                const uint32_t buf = tmem_base + (c.t & 1) * C::kBuf;
                if (c.j == 0) mbar_wait(&q_full[qs], (c.wc / QST) & 1);
                mbar_wait(&kv_full[s], (c.t / KVST) & 1);
                tc_fence_after();
                // the buffer is free: its previous user (step t-2) had its dQ MMA issued by this thread earlier, after the
                // softmax warps reported ds_full, and tcgen05.mma executes in issue order
                const int valid = s_valid<WK>(p.Sk, c.j, nkv);
                const uint32_t idesc = umma_idesc_bf16(max(16, (valid + 15) & ~15), false, false);
                const uint32_t sQ = sbase + C::Q0 + qs * C::kTileBytes, sDO = sbase + C::DO0 + qs * C::kTileBytes;
                const uint32_t sK = sbase + C::K0 + s * C::kStageN, sV = sbase + C::V0 + s * C::kStageN;
#pragma unroll
                for (int k = 0; k < HD_PAD / 16; ++k)
                    umma_bf16_ss(buf, umma_smem_desc_sw128(sQ + (k >> 2) * kAtomBytes + (k & 3) * 32, 16, 1024),
                                 umma_smem_desc_sw128(sK + (k >> 2) * C::kAtomN + (k & 3) * 32, 16, 1024), idesc, k != 0);
#pragma unroll
                for (int k = 0; k < HD_PAD / 16; ++k)
                    umma_bf16_ss(buf + WK, umma_smem_desc_sw128(sDO + (k >> 2) * kAtomBytes + (k & 3) * 32, 16, 1024),
                                 umma_smem_desc_sw128(sV + (k >> 2) * C::kAtomN + (k & 3) * 32, 16, 1024), idesc, k != 0);
                umma_commit(&sdp_full[c.t & 1]);
                // Q and dO are read by the score MMAs only
                if (c.j == nkv - 1) umma_commit(&q_empty[qs]);
            };
            Cur cs{(int)blockIdx.x, 0, 0u, 0u}, cg = cs;
            for (int i = 0; i < 2 && cs.w < num_work; ++i) { scores(cs); adv(cs); }
            while (cg.w < num_work) {
                const int s = cg.t % KVST;
                mbar_wait(&ds_full[cg.t & 1], (cg.t >> 1) & 1);
                // use_acc = 0 overwrites the accumulator: the previous item's rows must have been read out (the softmax
                // warps do that right after this step's softmax, see below)
                if (cg.j == 0 && cg.wc > 0) mbar_wait(dq_empty, (cg.wc - 1) & 1);
                tc_fence_after();
                const int valid = s_valid<WK>(p.Sk, cg.j, nkv);
                const int n16 = max(16, (valid + 15) & ~15), hA = split_a(n16);
                const uint32_t tdS = tmem_base + (cg.t & 1) * C::kBuf + WK;
                const uint32_t sK = sbase + C::K0 + s * C::kStageN;
                for (int k = 0; k < (n16 >> 4); ++k) {     // dQ += dS (TMEM, 16 keys = 8 columns per step) . K (MN-major)
                    const int q0 = k * 16;
                    const uint32_t acol = q0 < hA ? q0 / 2 : hA + (q0 - hA) / 2;
                    umma_bf16_ts(tmem_dQ, tdS + acol, umma_smem_desc_sw128(sK + k * 2048, C::kAtomN, 1024), idesc_dq,
                                 (cg.j | k) != 0);
                }
                umma_commit(&kv_empty[s]);
                if (cg.j == nkv - 1) umma_commit(dq_full);
                adv(cg);
                if (cs.w < num_work) { scores(cs); adv(cs); }
            }
        }
    } else {
        // ------------------------------------------------------------------ softmax warps 0..7
        // TMEM lane quarter = warp & 3 (row r of the tile); half = warp >> 2 picks the key range of the row
        const int r = threadIdx.x & 127;
        const int half = warp >> 2;
        const uint32_t lane_off = (uint32_t)((warp & 3) * 32) << 16;
        const float sc2 = p.scale * kLog2e;
        uint32_t t = 0, wc = 0;
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
        // The accumulator read-out of a work item is DEFERRED until this warp has finished the first softmax step of the
        // NEXT item (ncu: with the read-out right after the last step, 26 % of the stall samples sat on the wait for the
        // last dQ MMA): by then the MMA has long completed, and the tensor pipe already holds the next scores.
        __nv_bfloat16* pend = nullptr;     // this thread's dQ row of the previous item (null: none / row past Sq)
        bool have_pend = false;
        uint32_t pend_wc = 0;
        auto read_out = [&]() {
            constexpr int HC = HD_PAD / 2;
            mbar_wait(dq_full, pend_wc & 1);
            tc_fence_after();
            float acc[HC];
            tmem_load_cols<HC>(tmem_dQ + lane_off + half * HC, acc);
            tc_fence_before();
            __syncwarp();
            if (lane_id() == 0) mbar_arrive(dq_empty);
            if (pend) store_cols_bf16<HC>(pend, half * HC, acc, p.D);
            have_pend = false;
        };
        for (int w = blockIdx.x; w < num_work; w += gridDim.x, ++wc) {
            const int qt = w % nqt, bh = w / nqt;
            const int h = bh % p.H, b = bh / p.H;
            const int qi = qt * kTile + r;
            const bool row_ok = qi < p.Sq;
            const float nlse2 = -lse_n * kLog2e;
            const float ndlt = -dlt_n * p.scale;     // dS = p * (dP*scale - delta*scale)
            // the next item's statistics are requested now: a whole item of lead over their first use
            if (w + (int)gridDim.x < num_work) load_stats(w + gridDim.x, lse_n, dlt_n);
            const float* mrow = (!kPlain && p.mask) ? (p.mask + (int64_t)(p.mask_bmod ? b % p.mask_bmod : b) * p.mask_bs + (int64_t)h * p.mask_hs) + (int64_t)(row_ok ? qi : 0) * p.mask_qs : nullptr;
            const bool dropping = !kPlain && p.drop.p > 0.f;
            const uint32_t drop_key = dropping ? drop_row_key(p.drop, (uint64_t)(b * p.H + h) * p.Sq + (row_ok ? qi : 0)) : 0u;
            const uint32_t drop_thr = drop_thresh16(p.drop);
            for (int j = 0; j < nkv; ++j, ++t) {
                const int valid = s_valid<WK>(p.Sk, j, nkv);
                const int n16 = max(16, (valid + 15) & ~15), hA = split_a(n16);
                const int c_begin = half == 0 ? 0 : hA, c_end = half == 0 ? hA : n16;
                const uint32_t tS = tmem_base + (t & 1) * C::kBuf + lane_off, tdP = tS + WK;
                mbar_wait(&sdp_full[t & 1], (t >> 1) & 1);
                tc_fence_after();
                for (int c0 = c_begin; c0 < c_end; c0 += 32) {
                    const bool full = c_end - c0 >= 32;
                    uint32_t sv[32], dv[32];
                    if (full) {
                        tmem_ld_x32(tS + c0, sv);
                        tmem_ld_x32(tdP + c0, dv);
                    } else {      // 16-column remainder of this half
                        uint32_t a16[16], b16[16];
                        tmem_ld_x16(tS + c0, a16);
                        tmem_ld_x16(tdP + c0, b16);
#pragma unroll
                        for (int q = 0; q < 16; ++q) { sv[q] = a16[q]; dv[q] = b16[q]; sv[16 + q] = 0; dv[16 + q] = 0; }
                    }
                    tmem_ld_wait();
                    float ds[32];
                    const int lim = valid - c0;       // keys of this chunk that exist
                    if ((kPlain || (!mrow && !dropping)) && lim >= 32) {
#pragma unroll
                        for (int i = 0; i < 32; ++i)
                            ds[i] = ex2_fast(fmaf(__uint_as_float(sv[i]), sc2, nlse2)) * fmaf(__uint_as_float(dv[i]), p.scale, ndlt);
                    } else {
#pragma unroll
                        for (int i = 0; i < 32; i += 2) {
                            const uint32_t bits = dropping ? drop_pair_bits(drop_key, (uint32_t)(j * WK + c0 + i) >> 1) : 0u;
#pragma unroll
                            for (int e = 0; e < 2; ++e) {
                                float s = fmaf(__uint_as_float(sv[i + e]), sc2, nlse2);
                                if (mrow && i + e < lim) s = fmaf(mrow[j * WK + c0 + i + e], kLog2e, s);
                                float dpv = __uint_as_float(dv[i + e]);
                                if (dropping)      // dP flows only through the kept probabilities
                                    dpv *= ((e ? bits >> 16 : bits & 0xFFFFu) >= drop_thr) ? p.drop.inv_keep : 0.0f;
                                ds[i + e] = i + e < lim ? ex2_fast(s) * fmaf(dpv, p.scale, ndlt) : 0.f;
                            }
                        }
                    }
                    // packed bf16 output: 16 (or 8) columns at the start of this half's own range, behind its reads
                    const uint32_t ocol = c_begin + (c0 - c_begin) / 2;
                    if (full) {
                        tmem_store_bf16x32(tdP + ocol, ds);
                    } else {
                        uint32_t w0[8];
#pragma unroll
                        for (int q = 0; q < 8; ++q) w0[q] = pack_bf16x2(ds[2 * q], ds[2 * q + 1]);
                        tmem_st_x8(tdP + ocol, w0);
                    }
                }
                tmem_st_wait();
                tc_fence_before();
                __syncwarp();
                if (lane_id() == 0) mbar_arrive(&ds_full[t & 1]);
                if (j == 0 && have_pend) read_out();      // the previous item's dQ rows (the issuer waits for dq_empty)
            }
            pend = row_ok ? p.dq + (int64_t)b * p.dq_bs + (int64_t)qi * p.dq_rs + (int64_t)h * p.dq_hs : nullptr;
            pend_wc = wc;
            have_pend = true;
        }
        if (have_pend) read_out();
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
// dQ of mico_attention_bwd for head_dim 81..96 (arguments validated and delta already queued by the caller)
int attention_dq_pipelined(const MicoAttnArgs* a, cudaStream_t stream) {
    const int nkvb = a->kv_index ? a->n_kv : a->B;
    CUtensorMap tq, tks, tvs, tdo;
    int rc;
    if ((rc = make_attn_tmap(&tq, a->q, a->D, a->H, a->Sq, a->B, a->q_bs, a->q_rs, a->q_hs))) return rc;
    if ((rc = make_attn_tmap(&tdo, a->dout, a->D, a->H, a->Sq, a->B, a->do_bs, a->do_rs, a->do_hs))) return rc;
    if ((rc = make_attn_tmap(&tks, a->k, a->D, a->H, a->Sk, nkvb, a->k_bs, a->k_rs, a->k_hs, DqCfg<96>::WK))) return rc;
    if ((rc = make_attn_tmap(&tvs, a->v, a->D, a->H, a->Sk, nkvb, a->v_bs, a->v_rs, a->v_hs, DqCfg<96>::WK))) return rc;
    AttnBwdParams p;
    p.B = a->B; p.H = a->H; p.Sq = a->Sq; p.Sk = a->Sk; p.D = a->D; p.scale = a->scale;
    p.mask = a->mask; p.mask_bs = a->mask_bs; p.mask_qs = a->mask_qs; p.mask_hs = a->mask_hs; p.mask_bmod = a->mask_bmod;
    p.drop.p = a->dropout_p; p.drop.inv_keep = a->dropout_p < 1.f ? 1.f / (1.f - a->dropout_p) : 0.f; p.drop.seed = a->dropout_seed;
    p.lse = a->lse; p.delta = a->delta;
    p.dq = reinterpret_cast<__nv_bfloat16*>(a->dq); p.dq_bs = a->dq_bs; p.dq_rs = a->dq_rs; p.dq_hs = a->dq_hs;
    p.dk = nullptr; p.dv = nullptr;
    p.kv_index = a->kv_index; p.n_kv = a->n_kv; p.grp_ptr = a->grp_ptr; p.grp_list = a->grp_list;
    const int work_q = a->B * a->H * m_tiles(a->Sq);
    const int grid = work_q < num_sms() ? work_q : num_sms();
    auto launch = [&](auto kq) -> int {
        MICO_CHECK_CUDA(cudaFuncSetAttribute(kq, cudaFuncAttributeMaxDynamicSharedMemorySize, DqCfg<96>::TOTAL));
        kq<<<grid, kBwdThreads, DqCfg<96>::TOTAL, stream>>>(tq, tks, tvs, tdo, p);
        MICO_CHECK_CUDA(cudaGetLastError());
        return MICO_OK;
    };
    const bool plain = a->mask == nullptr && a->dropout_p == 0.0f;
    return plain ? launch(attn_dq_kernel<96, true>) : launch(attn_dq_kernel<96, false>);
}
}  // namespace mico
