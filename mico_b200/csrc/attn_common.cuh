// mico_b200 -- shared pieces of the attention forward/backward kernels.
#pragma once
#include "common.cuh"
#include "host_utils.h"

namespace mico {

constexpr int kAttThreads = 192;
constexpr int kTile = 128;            // rows per query tile == keys per KV tile
constexpr int kAtomBytes = 16384;     // 128 rows x 128 B (64 bf16): one 128B-swizzle atom column
constexpr float kLog2e = 1.4426950408889634f;
constexpr float kLn2 = 0.6931471805599453f;

// 4-D tensor map over a [B,S,H,D] strided bf16 tensor: dims (d, head, row, batch), box (64,1,128,1).
int make_attn_tmap(CUtensorMap* tm, const void* base, int D, int H, int S, int B, int64_t bs, int64_t rs, int64_t hs);

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

// TMEM [this warp's 32 lanes] x HD_PAD columns -> registers
template <int HD_PAD, bool kAccumulate>
__device__ __forceinline__ void tmem_load_row(uint32_t taddr, float (&acc)[HD_PAD]) {
#pragma unroll
    for (int c = 0; c < HD_PAD / 32; ++c) {
        uint32_t v[32];
        tmem_ld_x32(taddr + c * 32, v);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) {
            if (kAccumulate) acc[c * 32 + i] += __uint_as_float(v[i]);
            else acc[c * 32 + i] = __uint_as_float(v[i]);
        }
    }
    if constexpr (HD_PAD % 32 != 0) {
        uint32_t v[16];
        tmem_ld_x16(taddr + (HD_PAD / 32) * 32, v);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            if (kAccumulate) acc[(HD_PAD / 32) * 32 + i] += __uint_as_float(v[i]);
            else acc[(HD_PAD / 32) * 32 + i] = __uint_as_float(v[i]);
        }
    }
}

// 128-thread named barrier among the softmax warps (id 1; id 0 is __syncthreads)
__device__ __forceinline__ void softmax_group_sync() { asm volatile("bar.sync 1, 128;" ::: "memory"); }

}  // namespace mico
