// mico_b200 -- library-wide host utilities behind the C-ABI (error slot, TMA descriptor encode, counters).
#include <atomic>
#include <mutex>
#include <vector>
#include <string.h>

#include "host_utils.h"

namespace mico {

static thread_local char g_err[512] = "";
static std::atomic<int64_t> g_launches{0};

void set_last_error(const char* file, int line, const char* msg) {
    snprintf(g_err, sizeof(g_err), "%s:%d: %s", file, line, msg);
}

void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

// ------------------------------------------------------------------ per-family event timing
struct ProfRec { cudaEvent_t e0, e1; int kind; double work; };
static std::mutex g_prof_mu;
static std::vector<ProfRec> g_prof;        // records of the current session
static std::vector<cudaEvent_t> g_ev_pool; // recycled events
static std::atomic<bool> g_prof_on{false};

static cudaEvent_t prof_event() {
    if (!g_ev_pool.empty()) { cudaEvent_t e = g_ev_pool.back(); g_ev_pool.pop_back(); return e; }
    cudaEvent_t e = nullptr;
    cudaEventCreate(&e);
    return e;
}

ProfScope::ProfScope(int kind, double work, cudaStream_t s) : slot(-1), stream(s) {
    if (!g_prof_on.load(std::memory_order_relaxed)) return;
    std::lock_guard<std::mutex> lk(g_prof_mu);
    ProfRec r;
    r.e0 = prof_event(); r.e1 = prof_event(); r.kind = kind; r.work = work;
    if (!r.e0 || !r.e1) return;
    cudaEventRecord(r.e0, s);
    g_prof.push_back(r);
    slot = (int)g_prof.size() - 1;
}
ProfScope::~ProfScope() {
    if (slot < 0) return;
    std::lock_guard<std::mutex> lk(g_prof_mu);
    if (slot < (int)g_prof.size()) cudaEventRecord(g_prof[slot].e1, stream);
}

// ---- side stream (see host_utils.h)
namespace {
struct SideCtx {
    cudaStream_t stream = nullptr;
    cudaEvent_t ev[64] = {};
    unsigned next = 0;
};
std::mutex g_side_mu;
SideCtx g_side[16];
cudaEvent_t side_event(SideCtx& c) {
    cudaEvent_t& e = c.ev[c.next++ & 63];
    if (!e && cudaEventCreateWithFlags(&e, cudaEventDisableTiming) != cudaSuccess) return nullptr;
    return e;
}
SideCtx* side_ctx() {
    static const bool on = [] { const char* e = getenv("MICO_ATTN_TAIL_OVERLAP"); return !(e && e[0] == '0'); }();
    int dev = 0;
    if (!on || cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 16) return nullptr;
    SideCtx& c = g_side[dev];
    if (!c.stream && cudaStreamCreateWithFlags(&c.stream, cudaStreamNonBlocking) != cudaSuccess) return nullptr;
    return &c;
}
}  // namespace

cudaStream_t side_fork(cudaStream_t main) {
    std::lock_guard<std::mutex> lk(g_side_mu);
    SideCtx* c = side_ctx();
    if (!c) return nullptr;
    cudaEvent_t e = side_event(*c);
    if (!e || cudaEventRecord(e, main) != cudaSuccess || cudaStreamWaitEvent(c->stream, e, 0) != cudaSuccess) return nullptr;
    return c->stream;
}

void side_join(cudaStream_t main) {
    std::lock_guard<std::mutex> lk(g_side_mu);
    SideCtx* c = side_ctx();
    if (!c) return;
    cudaEvent_t e = side_event(*c);
    if (e && cudaEventRecord(e, c->stream) == cudaSuccess) cudaStreamWaitEvent(main, e, 0);
}

static std::atomic<int> g_reserved_sms{0};

int num_sms() {
    static int n = [] {
        int dev = 0, v = 0;
        if (cudaGetDevice(&dev) != cudaSuccess) return kDefaultSMs;
        if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v <= 0)
            return kDefaultSMs;
        return v;
    }();
    const int r = g_reserved_sms.load(std::memory_order_relaxed);
    return n - r > 2 ? n - r : 2;
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


int make_tmap_tile64(CUtensorMap* out, const void* base, bool fp32, uint64_t cols, uint64_t rows, uint64_t pitch_bytes) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) {
        set_last_error(__FILE__, __LINE__, "cuTensorMapEncodeTiled not available (no CUDA driver?)");
        return MICO_ERR_DRIVER;
    }
    const cuuint64_t gdim[2] = {cols, rows};
    const cuuint64_t gstride[1] = {pitch_bytes};
    const cuuint32_t bx[2] = {fp32 ? 16u : 32u, 32u};      // 64 bytes x 32 rows
    const cuuint32_t es[2] = {1, 1};
    CUresult r = fn(out, fp32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2,
                    const_cast<void*>(base), gdim, gstride, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        char msg[200];
        snprintf(msg, sizeof(msg), "cuTensorMapEncodeTiled (epilogue tile) failed (%d): %llu x %llu pitch %llu", (int)r,
                 (unsigned long long)cols, (unsigned long long)rows, (unsigned long long)pitch_bytes);
        set_last_error(__FILE__, __LINE__, msg);
        return MICO_ERR_DRIVER;
    }
    return MICO_OK;
}

}  // namespace mico

extern "C" int mico_version(void) { return 100; }
extern "C" int mico_set_reserved_sms(int n) {
    if (n < 0 || n > 64 || (n & 1)) {
        mico::set_last_error(__FILE__, __LINE__, "reserved SMs must be an even number in [0, 64]");
        return MICO_ERR_INVALID_ARG;
    }
    mico::g_reserved_sms.store(n);
    return MICO_OK;
}
extern "C" const char* mico_last_error(void) { return mico::g_err; }
extern "C" int64_t mico_launch_count(void) { return mico::g_launches.load(); }
extern "C" void mico_reset_launch_count(void) { mico::g_launches.store(0); }

extern "C" int mico_profile_enable(int on) {
    using namespace mico;
    std::lock_guard<std::mutex> lk(g_prof_mu);
    for (auto& r : g_prof) { g_ev_pool.push_back(r.e0); g_ev_pool.push_back(r.e1); }
    g_prof.clear();
    g_prof_on.store(on != 0);
    return MICO_OK;
}

extern "C" int mico_profile_collect(double* ms, double* work, int64_t* count, int nkinds) {
    using namespace mico;
    MICO_CHECK_ARG(ms && work && count && nkinds >= kProfKinds);
    MICO_CHECK_CUDA(cudaDeviceSynchronize());
    std::lock_guard<std::mutex> lk(g_prof_mu);
    for (int i = 0; i < nkinds; ++i) { ms[i] = 0.0; work[i] = 0.0; count[i] = 0; }
    for (auto& r : g_prof) {
        float t = 0.f;
        if (cudaEventElapsedTime(&t, r.e0, r.e1) != cudaSuccess) continue;
        ms[r.kind] += t; work[r.kind] += r.work; count[r.kind] += 1;
    }
    return MICO_OK;
}
