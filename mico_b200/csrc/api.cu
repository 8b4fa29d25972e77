// mico_b200 -- library-wide host utilities behind the C-ABI (error slot, TMA descriptor encode, counters).
#include <atomic>
#include <mutex>
#include <string.h>

#include "host_utils.h"

namespace mico {

static thread_local char g_err[512] = "";
static std::atomic<int64_t> g_launches{0};

void set_last_error(const char* file, int line, const char* msg) {
    snprintf(g_err, sizeof(g_err), "%s:%d: %s", file, line, msg);
}

void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

int num_sms() {
    static int n = [] {
        int dev = 0, v = 0;
        if (cudaGetDevice(&dev) != cudaSuccess) return kDefaultSMs;
        if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v <= 0)
            return kDefaultSMs;
        return v;
    }();
    return n;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
            q != cudaDriverEntryPointSuccess)
            p = nullptr;
        return reinterpret_cast<EncodeTiledFn>(p);
    }();
    return fn;
}

int make_tmap_bf16(CUtensorMap* out, const void* base, int rank, const uint64_t* dims,
                   const uint64_t* strides_bytes, const uint32_t* box) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) {
        set_last_error(__FILE__, __LINE__, "cuTensorMapEncodeTiled not available (no CUDA driver?)");
        return MICO_ERR_DRIVER;
    }
    cuuint64_t gdim[5];
    cuuint64_t gstride[4];
    cuuint32_t bx[5], es[5];
    for (int i = 0; i < rank; ++i) {
        gdim[i] = dims[i];
        bx[i] = box[i];
        es[i] = 1;
        if (i > 0) gstride[i - 1] = strides_bytes[i];
    }
    CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(base), gdim, gstride,
                    bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                    CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        char msg[200];
        snprintf(msg, sizeof(msg), "cuTensorMapEncodeTiled failed (%d): rank %d dims %llu,%llu box %u,%u stride %llu",
                 (int)r, rank, (unsigned long long)dims[0], (unsigned long long)(rank > 1 ? dims[1] : 0), box[0],
                 rank > 1 ? box[1] : 0, (unsigned long long)(rank > 1 ? strides_bytes[1] : 0));
        set_last_error(__FILE__, __LINE__, msg);
        return MICO_ERR_DRIVER;
    }
    return MICO_OK;
}

}  // namespace mico

extern "C" int mico_version(void) { return 100; }
extern "C" const char* mico_last_error(void) { return mico::g_err; }
extern "C" int64_t mico_launch_count(void) { return mico::g_launches.load(); }
extern "C" void mico_reset_launch_count(void) { mico::g_launches.store(0); }
