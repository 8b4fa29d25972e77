// mico_b200 -- host-side helpers shared by the C-ABI translation units.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/mico_b200.h"

namespace mico {

#define MICO_CHECK_ARG(cond)                                                             \
    do {                                                                                 \
        if (!(cond)) {                                                                   \
            mico::set_last_error(__FILE__, __LINE__, "invalid argument: " #cond);        \
            return MICO_ERR_INVALID_ARG;                                                 \
        }                                                                                \
    } while (0)

#define MICO_CHECK_CUDA(expr)                                                            \
    do {                                                                                 \
        cudaError_t _e = (expr);                                                         \
        if (_e != cudaSuccess) {                                                         \
            mico::set_last_error(__FILE__, __LINE__, cudaGetErrorString(_e));            \
            return MICO_ERR_CUDA;                                                        \
        }                                                                                \
    } while (0)

void set_last_error(const char* file, int line, const char* msg);

// Build a tiled bf16 tensor map with 128-byte swizzle.  dims/strides are innermost-first;
// strides[0] is implied (2 bytes), strides[i>0] in BYTES.  Returns 0 or a MICO_ERR_* code.
int make_tmap_bf16(CUtensorMap* out, const void* base, int rank, const uint64_t* dims,
                   const uint64_t* strides_bytes, const uint32_t* box);

// [rows x cols] bf16 / fp32 matrix seen through 64-byte x 32-row boxes with 64-byte swizzle: the per-warp epilogue
// tiles of the GEMM (TMA stores).  pitch in bytes (multiple of 16).
int make_tmap_tile64(CUtensorMap* out, const void* base, bool fp32, uint64_t cols, uint64_t rows, uint64_t pitch_bytes);

inline int ceil_div(int a, int b) { return (a + b - 1) / b; }

constexpr int kDefaultSMs = 148;
int num_sms();

void count_launch(int n = 1);

// Fork / join onto a library-owned side stream (one per device) so that a small independent kernel -- the SIMT tail rows of
// attention -- runs CONCURRENTLY with the tile kernel it complements instead of after it.  side_fork(main): the side stream
// waits for everything queued on `main` so far and is returned; side_join(main): `main` waits for everything queued on the
// side stream so far.  Returns nullptr / does nothing when overlap is switched off (MICO_ATTN_TAIL_OVERLAP=0) or on error:
// callers then launch on `main`.
cudaStream_t side_fork(cudaStream_t main);
void side_join(cudaStream_t main);

// Per-kernel-family device timing (mico_profile_* in the C-ABI): when enabled, every entry point brackets its
// launches with a cudaEvent pair on the launching stream; mico_profile_collect() reads them back.
enum ProfKind { kProfGemm = 0, kProfAttnFwd, kProfAttnBwd, kProfLnFwd, kProfLnBwd, kProfOther, kProfKinds };
struct ProfScope {
    int slot;
    cudaStream_t stream;
    ProfScope(int kind, double work, cudaStream_t s);
    ~ProfScope();
};

}  // namespace mico
