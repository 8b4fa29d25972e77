// mico_b200 -- shared pieces of the attention forward/backward kernels.
#pragma once
#include "common.cuh"
#include <stdlib.h>
#include "host_utils.h"

namespace mico {

constexpr int kAttThreads = 192;      // forward: 4 softmax warps + TMA + MMA
constexpr int kBwdThreads = 320;      // backward: 8 softmax warps + TMA + MMA
constexpr int kTile = 128;            // rows per M tile (TMEM lanes) and the base size of a streamed N tile
constexpr int kTileN = 144;           // a streamed tile may absorb a remainder of up to 16 rows (257 = 128 + 129)
constexpr int kTailMax = 8;           // M-side remainders of at most this many rows go to the SIMT tail kernels
constexpr int kAtomBytes = 16384;     // 128 rows x 128 B (64 bf16): one 128B-swizzle atom column
constexpr int kAtomBytesN = kTileN * 128;   // same for a 144-row streamed tile (18432 = 18 x 1024)
constexpr int kExtRows = kTileN - kTile;    // rows of the remainder extension box
constexpr float kLog2e = 1.4426950408889634f;
constexpr float kLn2 = 0.6931471805599453f;

// 4-D tensor map over a [B,S,H,D] strided bf16 tensor: dims (d, head, row, batch), box (64,1,rows,1).
int make_attn_tmap(CUtensorMap* tm, const void* base, int D, int H, int S, int B, int64_t bs, int64_t rs, int64_t hs,
                   int box_rows = kTile);

// SIMT kernels for the M-side remainder rows (attention_tail.cu); called by the entry points when m_tail_rows() > 0.
int attention_tail_fwd(const MicoAttnArgs* a, cudaStream_t stream);
int attention_tail_bwd(const MicoAttnArgs* a, cudaStream_t stream);

// Tiling of a sequence of S rows.
//   streamed (N) side: tiles start at 128*j; all hold 128 rows except the last, which holds `last` <= 144:
//                      a remainder of <= 16 rows is absorbed into the previous tile (S = 257 -> 128 + 129).
//   owned (M) side:    128-row tiles (one TMEM lane per row); a remainder of <= kTailMax rows is left to the
//                      SIMT tail kernels instead of paying a whole tile pipeline for it (S = 257 -> 2 tiles + 1 row).
struct NTiling { int n, last; };
__host__ __device__ inline NTiling n_tiling(int S) {
    NTiling t;
    if (S <= kTileN) { t.n = 1; t.last = S; return t; }
    const int full = S / kTile, r = S % kTile;
    if (r == 0) { t.n = full; t.last = kTile; }
    else if (r <= kExtRows) { t.n = full; t.last = kTile + r; }
    else { t.n = full + 1; t.last = r; }
    return t;
}
__host__ __device__ inline int n_valid(const NTiling& t, int j) { return j < t.n - 1 ? kTile : t.last; }
__host__ __device__ inline int m_tail_rows(int S) {
    const int r = S % kTile;
    return (S > kTile && r > 0 && r <= kTailMax) ? r : 0;
}
__host__ __device__ inline int m_tiles(int S) { return m_tail_rows(S) ? S / kTile : (S + kTile - 1) / kTile; }
// same with the SIMT tail path switched off (shared K/V entries: the remainder rows take one more, mostly empty, tile)
__host__ __device__ inline int m_tiles_nt(int S, bool no_tail) { return no_tail ? (S + kTile - 1) / kTile : m_tiles(S); }

// TMA-load one streamed tile (rows 128*j ..) of a [B,S,H,D] tensor into `dst` (kAtoms atoms of kAtomBytesN):
// the 128-row box, plus the 16-row extension box when the tile holds more than 128 rows.
template <int kAtoms>
__device__ __forceinline__ void load_n_tile(uint8_t* dst, const CUtensorMap* tm, const CUtensorMap* tm_ext, uint64_t* bar,
                                            int h, int row0, int b, bool ext) {
#pragma unroll
    for (int a = 0; a < kAtoms; ++a) {
        tma_load_4d(dst + a * kAtomBytesN, tm, bar, a * 64, h, row0, b);
        if (ext) tma_load_4d(dst + a * kAtomBytesN + kAtomBytes, tm_ext, bar, a * 64, h, row0 + kTile, b);
    }
}
__host__ __device__ constexpr uint32_t n_tile_bytes(int atoms, bool ext) {
    return (uint32_t)atoms * (kAtomBytes + (ext ? kExtRows * 128 : 0));
}

// Store 32 consecutive columns [col0, col0+32) of row r of a 128 x 128 bf16 tile into shared memory in the
// canonical K-major 128B-swizzle layout (two 64-column atoms), as expected by a UMMA smem descriptor.
__device__ __forceinline__ void store_tile_chunk32(uint8_t* tile, int r, int col0, const float (&v)[32]) {
#pragma unroll
    for (int g = 0; g < 4; ++g) {
        const int col = col0 + g * 8;
        const int atom = col >> 6, chunk = (col & 63) >> 3;
        uint8_t* dst = tile + atom * kAtomBytes + r * 128 + ((chunk ^ (r & 7)) << 4);
        *reinterpret_cast<uint4*>(dst) =
            make_uint4(pack_bf16x2(v[g * 8 + 0], v[g * 8 + 1]), pack_bf16x2(v[g * 8 + 2], v[g * 8 + 3]),
                       pack_bf16x2(v[g * 8 + 4], v[g * 8 + 5]), pack_bf16x2(v[g * 8 + 6], v[g * 8 + 7]));
    }
}

// Write one row of HD_PAD fp32 accumulators (held by this thread) as bf16; D % 8 == 0.
template <int HD_PAD>
__device__ __forceinline__ void store_row_bf16(__nv_bfloat16* dst, const float (&acc)[HD_PAD], int D, float mul) {
#pragma unroll
    for (int g = 0; g < HD_PAD / 8; ++g) {
        if (g * 8 < D)
            *reinterpret_cast<uint4*>(dst + g * 8) =
                make_uint4(pack_bf16x2(acc[g * 8 + 0] * mul, acc[g * 8 + 1] * mul),
                           pack_bf16x2(acc[g * 8 + 2] * mul, acc[g * 8 + 3] * mul),
                           pack_bf16x2(acc[g * 8 + 4] * mul, acc[g * 8 + 5] * mul),
                           pack_bf16x2(acc[g * 8 + 6] * mul, acc[g * 8 + 7] * mul));
    }
}

// TMEM [this warp's 32 lanes] x HD_PAD columns -> registers.  All loads are issued before the single wait: one TMEM
// round trip per row instead of one per 32 columns (ncu round 1: the accumulator read-out of the backward kernels held
// ~20 % of their stall samples with a wait after every chunk).
template <int HD_PAD, bool kAccumulate>
__device__ __forceinline__ void tmem_load_row(uint32_t taddr, float (&acc)[HD_PAD]) {
    uint32_t v[HD_PAD];
#pragma unroll
    for (int c = 0; c < HD_PAD / 32; ++c) {
        uint32_t (&chunk)[32] = *reinterpret_cast<uint32_t (*)[32]>(&v[c * 32]);
        tmem_ld_x32(taddr + c * 32, chunk);
    }
    if constexpr (HD_PAD % 32 != 0) {
        uint32_t (&chunk)[16] = *reinterpret_cast<uint32_t (*)[16]>(&v[(HD_PAD / 32) * 32]);
        tmem_ld_x16(taddr + (HD_PAD / 32) * 32, chunk);
    }
    tmem_ld_wait();
#pragma unroll
    for (int i = 0; i < HD_PAD; ++i) {
        if (kAccumulate) acc[i] += __uint_as_float(v[i]);
        else acc[i] = __uint_as_float(v[i]);
    }
}

// a global fp32 load that stays where it is written (the compiler otherwise sinks prefetches down to their first use)
__device__ __forceinline__ float ldg_f32_pinned(const float* p) {
    float v;
    asm volatile("ld.global.nc.f32 %0, [%1];" : "=f"(v) : "l"(p) : "memory");
    return v;
}

// named barrier among the softmax warps (id 1; id 0 is __syncthreads)
__device__ __forceinline__ void softmax_group_sync() { asm volatile("bar.sync 1, 128;" ::: "memory"); }
__device__ __forceinline__ void softmax_group_sync256() { asm volatile("bar.sync 1, 256;" ::: "memory"); }

// 32 fp32 values -> 16 packed bf16x2 words -> this warp's lanes x 16 TMEM columns (A operand of a TS-mode MMA:
// column c of the operand region holds elements 2c (low half) and 2c+1 (high half) of the row).
__device__ __forceinline__ void tmem_store_bf16x32(uint32_t taddr, const float (&v)[32]) {
    uint32_t w[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) w[i] = pack_bf16x2(v[2 * i], v[2 * i + 1]);
    tmem_st_x16(taddr, w);
}

// MICO_ATTN_DROP_ONLY=0: calls with dropout and no mask go through the general (kMode 0) kernels (A/B switch for measurements)
inline bool drop_only_enabled() {
    static const bool on = [] { const char* e = getenv("MICO_ATTN_DROP_ONLY"); return !(e && e[0] == '0'); }();
    return on;
}

}  // namespace mico
