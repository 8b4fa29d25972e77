// mico_b200 -- K4 remainder rows: the <= 8 rows of the "owned" side of attention that do not fill a 128-row
// tensor-core tile (the ViT's 257th token, eva_vit_model.py:613-616: 257 = 2 x 128 + 1) are computed here with
// plain fp32 SIMT math, one 128-thread block per (batch, head, row), instead of paying a whole TMA/tcgen05 tile pipeline
// for a tile that is 99 % padding.  Same arithmetic as the tile kernels:
//   forward   o_i = softmax(scale q_i K^T + mask_i) V,  lse_i                       (eva_vit_model.py:340-361)
//   dQ row    ds_ij = scale p_ij (dO_i.v_j - delta_i);  dq_i = sum_j ds_ij k_j
//   dK/dV row (a remainder KEY j): dv_j = sum_i p_ij dO_i;  dk_j = sum_i ds_ij q_i  over all queries i
// HBM-light: each block streams the other side's rows once (257 x 176 B for the ViT).
#include "attn_common.cuh"

namespace mico {
namespace {

constexpr int kTailMaxLen = 8192;   // streamed rows a tail warp can hold scores for (2 x 32 KB of smem)

struct TailParams {
    const __nv_bfloat16 *q, *k, *v, *o, *d_o;
    int64_t q_bs, q_rs, q_hs, k_bs, k_rs, k_hs, v_bs, v_rs, v_hs, o_bs, o_rs, o_hs, do_bs, do_rs, do_hs;
    __nv_bfloat16 *out0, *out1;     // fwd: o ; dq: dq ; dkv: dk, dv
    int64_t o0_bs, o0_rs, o0_hs, o1_bs, o1_rs, o1_hs;
    const float* mask; int64_t mask_bs, mask_qs, mask_hs; int mask_bmod;
    DropCfg drop;
    float* lse; const float* delta;
    int B, H, Sq, Sk, D, row0, rows;
    float scale;
    const int* kv_index;      // q-side kernels: K/V batch entry of query entry b (null = identity)
};

constexpr int kTailThreads = 128;

// dot product of a bf16 row (D % 8 == 0, 16-byte aligned) with an fp32 vector in shared memory
__device__ __forceinline__ float dot8(const uint4& u, const float* vec) {
    return bf16_lo(u.x) * vec[0] + bf16_hi(u.x) * vec[1] + bf16_lo(u.y) * vec[2] + bf16_hi(u.y) * vec[3] +
           bf16_lo(u.z) * vec[4] + bf16_hi(u.z) * vec[5] + bf16_lo(u.w) * vec[6] + bf16_hi(u.w) * vec[7];
}
// Dot product of a bf16 row with an fp32 vector, computed by a QUAD of lanes: lane `sub` of the quad takes the 16-byte
// chunks sub, sub+4, sub+8, ... so that the four lanes read 64 contiguous bytes per instruction (8 rows x 64 B per
// warp instruction instead of 32 rows x 16 B: the one-row-per-lane version was bound by L1 tag lookups, one per row
// and chunk).  All four lanes return the full sum.
__device__ __forceinline__ float row_dot_quad(const __nv_bfloat16* row, const float* vec, int D) {
    const int nch = D >> 3, sub = (int)lane_id() & 3;
    uint4 u[4];
#pragma unroll
    for (int t = 0; t < 4; ++t) {
        const int c = sub + 4 * t;
        u[t] = c < nch ? *reinterpret_cast<const uint4*>(row + 8 * c) : make_uint4(0, 0, 0, 0);
    }
    float acc = 0.f;
#pragma unroll
    for (int t = 0; t < 4; ++t) {
        const int c = sub + 4 * t;
        if (c < nch) acc += dot8(u[t], vec + 8 * c);
    }
    const unsigned qm = 0xFu << ((int)lane_id() & ~3);
    acc += __shfl_xor_sync(qm, acc, 1);
    acc += __shfl_xor_sync(qm, acc, 2);
    return acc;
}

__device__ __forceinline__ void load_vec(const __nv_bfloat16* row, float* vec, int D) {
    for (int c = threadIdx.x; c < D; c += kTailThreads) vec[c] = __bfloat162float(row[c]);
}

__device__ __forceinline__ float block_reduce(float v, float* red, bool is_max) {
    v = is_max ? warp_max(v) : warp_sum(v);
    __syncthreads();
    if (lane_id() == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    float r = red[0];
    for (int w = 1; w < kTailThreads / 32; ++w) r = is_max ? fmaxf(r, red[w]) : r + red[w];
    return r;
}

// out[d] = mul * sum_j w[j] * M[j][d].  Thread = (row group g of 8, 8-column chunk c of 16); 4 loads in flight.
__device__ __forceinline__ void weighted_rows(const __nv_bfloat16* M, int64_t rs, int n, const float* w, int D,
                                              float mul, __nv_bfloat16* out, float* red /* [8][128] */) {
    const int c = (threadIdx.x & 15) * 8, g = threadIdx.x >> 4;
    float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    if (c < D) {
        int j = g;
        for (; j + 56 < n; j += 64) {
            uint4 u[8];
#pragma unroll
            for (int t = 0; t < 8; ++t) u[t] = *reinterpret_cast<const uint4*>(M + (int64_t)(j + 8 * t) * rs + c);
#pragma unroll
            for (int t = 0; t < 8; ++t) {
                const float x = w[j + 8 * t];
                acc[0] += x * bf16_lo(u[t].x); acc[1] += x * bf16_hi(u[t].x); acc[2] += x * bf16_lo(u[t].y); acc[3] += x * bf16_hi(u[t].y);
                acc[4] += x * bf16_lo(u[t].z); acc[5] += x * bf16_hi(u[t].z); acc[6] += x * bf16_lo(u[t].w); acc[7] += x * bf16_hi(u[t].w);
            }
        }
        for (; j < n; j += 8) {
            const uint4 u = *reinterpret_cast<const uint4*>(M + (int64_t)j * rs + c);
            const float x = w[j];
            acc[0] += x * bf16_lo(u.x); acc[1] += x * bf16_hi(u.x); acc[2] += x * bf16_lo(u.y); acc[3] += x * bf16_hi(u.y);
            acc[4] += x * bf16_lo(u.z); acc[5] += x * bf16_hi(u.z); acc[6] += x * bf16_lo(u.w); acc[7] += x * bf16_hi(u.w);
        }
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 8; ++i) red[g * 128 + c + i] = acc[i];
    __syncthreads();
    if ((int)threadIdx.x < D) {
        float t = 0.f;
#pragma unroll
        for (int gg = 0; gg < 8; ++gg) t += red[gg * 128 + threadIdx.x];
        out[threadIdx.x] = __float2bfloat16(t * mul);
    }
}

// shared memory: vec0[128] vec1[128] red[8*128] small[8] then the score arrays
constexpr int kTailFixedFloats = 256 + 1024 + 8;

// mode 0: forward row; mode 1: dQ row.  One 128-thread block per (b, h, tail query row).
template <int MODE>
__global__ void __launch_bounds__(kTailThreads) attn_tail_q_kernel(TailParams p) {
    extern __shared__ float sm[];
    float* qv = sm;               // [128] q_i
    float* dov = sm + 128;        // [128] dO_i (mode 1)
    float* red = sm + 256;
    float* small = red + 1024;
    float* sc = sm + kTailFixedFloats;   // [Sk]
    const int t = blockIdx.x % p.rows, bh = blockIdx.x / p.rows;
    const int h = bh % p.H, b = bh / p.H;
    const int i = p.row0 + t;
    load_vec(p.q + b * p.q_bs + (int64_t)i * p.q_rs + h * p.q_hs, qv, p.D);
    if (MODE == 1) load_vec(p.d_o + b * p.do_bs + (int64_t)i * p.do_rs + h * p.do_hs, dov, p.D);
    __syncthreads();
    const int kvb = p.kv_index ? __ldg(p.kv_index + b) : b;
    const __nv_bfloat16* K = p.k + kvb * p.k_bs + h * p.k_hs;
    const __nv_bfloat16* V = p.v + kvb * p.v_bs + h * p.v_hs;
    const float* mrow = p.mask ? (p.mask + (int64_t)(p.mask_bmod ? b % p.mask_bmod : b) * p.mask_bs + (int64_t)h * p.mask_hs) + (int64_t)i * p.mask_qs : nullptr;
    const int64_t stat = ((int64_t)b * p.H + h) * p.Sq + i;
    if (MODE == 0) {
        float mx = -INFINITY;
        for (int j = threadIdx.x >> 2; j < p.Sk; j += kTailThreads / 4) {
            float s = p.scale * row_dot_quad(K + (int64_t)j * p.k_rs, qv, p.D);
            if (mrow) s += mrow[j];
            if ((threadIdx.x & 3) == 0) sc[j] = s;
            mx = fmaxf(mx, s);
        }
        mx = block_reduce(mx, small, true);
        float sum = 0.f;
        for (int j = threadIdx.x; j < p.Sk; j += kTailThreads) {
            const float e = __expf(sc[j] - mx);
            sum += e;
            sc[j] = p.drop.p > 0.f ? e * drop_mult_rc(p.drop, drop_row_key(p.drop, (uint64_t)stat), drop_thresh16(p.drop), j) : e;
        }
        sum = block_reduce(sum, small, false);
        weighted_rows(V, p.v_rs, p.Sk, sc, p.D, 1.0f / sum, p.out0 + b * p.o0_bs + (int64_t)i * p.o0_rs + h * p.o0_hs, red);
        if (p.lse && threadIdx.x == 0) p.lse[stat] = mx + __logf(sum);
    } else {
        const float lse = p.lse[stat], dlt = p.delta[stat];
        for (int j = threadIdx.x >> 2; j < p.Sk; j += kTailThreads / 4) {
            float s = p.scale * row_dot_quad(K + (int64_t)j * p.k_rs, qv, p.D);
            if (mrow) s += mrow[j];
            const float pr = __expf(s - lse);
            const float dp = row_dot_quad(V + (int64_t)j * p.v_rs, dov, p.D);
            const float m = p.drop.p > 0.f ? drop_mult_rc(p.drop, drop_row_key(p.drop, (uint64_t)stat), drop_thresh16(p.drop), j) : 1.0f;
            if ((threadIdx.x & 3) == 0) sc[j] = pr * (dp * m - dlt) * p.scale;
        }
        __syncthreads();
        weighted_rows(K, p.k_rs, p.Sk, sc, p.D, 1.0f, p.out0 + b * p.o0_bs + (int64_t)i * p.o0_rs + h * p.o0_hs, red);
    }
}

// One 128-thread block per (b, h, tail key row j): dk_j (out0), dv_j (out1).
__global__ void __launch_bounds__(kTailThreads) attn_tail_kv_kernel(TailParams p) {
    extern __shared__ float sm[];
    float* kv = sm;               // [128] k_j
    float* vv = sm + 128;         // [128] v_j
    float* red = sm + 256;
    float* pa = sm + kTailFixedFloats;   // [Sq] p_ij
    float* dsa = pa + p.Sq;              // [Sq] ds_ij
    const int t = blockIdx.x % p.rows, bh = blockIdx.x / p.rows;
    const int h = bh % p.H, b = bh / p.H;
    const int j = p.row0 + t;
    load_vec(p.k + b * p.k_bs + (int64_t)j * p.k_rs + h * p.k_hs, kv, p.D);
    load_vec(p.v + b * p.v_bs + (int64_t)j * p.v_rs + h * p.v_hs, vv, p.D);
    __syncthreads();
    const __nv_bfloat16* Q = p.q + b * p.q_bs + h * p.q_hs;
    const __nv_bfloat16* DO = p.d_o + b * p.do_bs + h * p.do_hs;
    const int64_t stat0 = ((int64_t)b * p.H + h) * p.Sq;
    for (int i = threadIdx.x >> 2; i < p.Sq; i += kTailThreads / 4) {
        float s = p.scale * row_dot_quad(Q + (int64_t)i * p.q_rs, kv, p.D);
        if (p.mask) s += (p.mask + (int64_t)(p.mask_bmod ? b % p.mask_bmod : b) * p.mask_bs + (int64_t)h * p.mask_hs)[(int64_t)i * p.mask_qs + j];
        const float pr = __expf(s - p.lse[stat0 + i]);
        const float dp = row_dot_quad(DO + (int64_t)i * p.do_rs, vv, p.D);
        if ((threadIdx.x & 3) == 0) {
            const float m = p.drop.p > 0.f ? drop_mult_rc(p.drop, drop_row_key(p.drop, (uint64_t)(stat0 + i)), drop_thresh16(p.drop), j) : 1.0f;
            pa[i] = pr * m;
            dsa[i] = pr * (dp * m - p.delta[stat0 + i]) * p.scale;
        }
    }
    __syncthreads();
    weighted_rows(Q, p.q_rs, p.Sq, dsa, p.D, 1.0f, p.out0 + b * p.o0_bs + (int64_t)j * p.o0_rs + h * p.o0_hs, red);
    weighted_rows(DO, p.do_rs, p.Sq, pa, p.D, 1.0f, p.out1 + b * p.o1_bs + (int64_t)j * p.o1_rs + h * p.o1_hs, red);
}

// Gradient of an additive attention bias that is a PARAMETER (Swin's relative position bias, swin.py:135-139):
//   d mask[g,h,i,j] += p_ij * (dO_i . v_j - delta_i)      summed over the batch entries b with b % G == g.
// One block per (b, h); K and V of the (small) window are staged in shared memory as fp32, threads own (i, j) pairs.
__global__ void __launch_bounds__(256) attn_dmask_kernel(TailParams p, float* __restrict__ dmask) {
    extern __shared__ float sm[];
    float* ks = sm;                              // [Sk][D]
    float* vs = ks + (size_t)p.Sk * p.D;         // [Sk][D]
    float* qs = vs + (size_t)p.Sk * p.D;         // [Sq][D]
    float* gs = qs + (size_t)p.Sq * p.D;         // [Sq][D]  dO
    const int h = blockIdx.x % p.H, b = blockIdx.x / p.H;
    const __nv_bfloat16* K = p.k + b * p.k_bs + h * p.k_hs;
    const __nv_bfloat16* V = p.v + b * p.v_bs + h * p.v_hs;
    const __nv_bfloat16* Q = p.q + b * p.q_bs + h * p.q_hs;
    const __nv_bfloat16* DO = p.d_o + b * p.do_bs + h * p.do_hs;
    for (int idx = threadIdx.x; idx < p.Sk * p.D; idx += blockDim.x) {
        const int j = idx / p.D, d = idx - j * p.D;
        ks[idx] = __bfloat162float(K[(int64_t)j * p.k_rs + d]);
        vs[idx] = __bfloat162float(V[(int64_t)j * p.v_rs + d]);
    }
    for (int idx = threadIdx.x; idx < p.Sq * p.D; idx += blockDim.x) {
        const int i = idx / p.D, d = idx - i * p.D;
        qs[idx] = __bfloat162float(Q[(int64_t)i * p.q_rs + d]);
        gs[idx] = __bfloat162float(DO[(int64_t)i * p.do_rs + d]);
    }
    __syncthreads();
    const int g = p.mask_bmod ? b % p.mask_bmod : b;
    const float* mbase = p.mask + (int64_t)g * p.mask_bs + (int64_t)h * p.mask_hs;
    float* dbase = dmask + (int64_t)g * p.mask_bs + (int64_t)h * p.mask_hs;
    const int64_t stat0 = ((int64_t)b * p.H + h) * p.Sq;
    for (int idx = threadIdx.x; idx < p.Sq * p.Sk; idx += blockDim.x) {
        const int i = idx / p.Sk, j = idx - i * p.Sk;
        float s = 0.f, dp = 0.f;
        for (int d = 0; d < p.D; ++d) {
            s += qs[i * p.D + d] * ks[j * p.D + d];
            dp += gs[i * p.D + d] * vs[j * p.D + d];
        }
        s = s * p.scale + mbase[(int64_t)i * p.mask_qs + j];
        const float pr = __expf(s - p.lse[stat0 + i]);
        atomicAdd(dbase + (int64_t)i * p.mask_qs + j, pr * (dp - p.delta[stat0 + i]));
    }
}

TailParams base_params(const MicoAttnArgs* a) {
    TailParams p{};
    p.q = reinterpret_cast<const __nv_bfloat16*>(a->q); p.q_bs = a->q_bs; p.q_rs = a->q_rs; p.q_hs = a->q_hs;
    p.k = reinterpret_cast<const __nv_bfloat16*>(a->k); p.k_bs = a->k_bs; p.k_rs = a->k_rs; p.k_hs = a->k_hs;
    p.v = reinterpret_cast<const __nv_bfloat16*>(a->v); p.v_bs = a->v_bs; p.v_rs = a->v_rs; p.v_hs = a->v_hs;
    p.o = reinterpret_cast<const __nv_bfloat16*>(a->o); p.o_bs = a->o_bs; p.o_rs = a->o_rs; p.o_hs = a->o_hs;
    p.d_o = reinterpret_cast<const __nv_bfloat16*>(a->dout); p.do_bs = a->do_bs; p.do_rs = a->do_rs; p.do_hs = a->do_hs;
    p.mask = a->mask; p.mask_bs = a->mask_bs; p.mask_qs = a->mask_qs; p.mask_hs = a->mask_hs; p.mask_bmod = a->mask_bmod;
    p.drop.p = a->dropout_p; p.drop.inv_keep = a->dropout_p < 1.f ? 1.f / (1.f - a->dropout_p) : 0.f; p.drop.seed = a->dropout_seed;
    p.lse = a->lse; p.delta = a->delta;
    p.B = a->B; p.H = a->H; p.Sq = a->Sq; p.Sk = a->Sk; p.D = a->D; p.scale = a->scale;
    p.kv_index = a->kv_index;
    return p;
}

template <typename K>
int launch_tail(K kern, const TailParams& p, int len, cudaStream_t stream) {
    if (len > kTailMaxLen) {
        set_last_error(__FILE__, __LINE__, "attention tail rows: the other side is longer than 8192 rows");
        return MICO_ERR_UNSUPPORTED;
    }
    const size_t smem = (kTailFixedFloats + 2 * (size_t)len) * sizeof(float);
    if (smem > 48 * 1024) MICO_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<p.B * p.H * p.rows, kTailThreads, smem, stream>>>(p);
    MICO_CHECK_CUDA(cudaGetLastError());
    count_launch();
    return MICO_OK;
}

}  // namespace

int attention_tail_fwd(const MicoAttnArgs* a, cudaStream_t stream) {
    TailParams p = base_params(a);
    p.rows = m_tail_rows(a->Sq);
    p.row0 = a->Sq - p.rows;
    p.out0 = reinterpret_cast<__nv_bfloat16*>(a->o); p.o0_bs = a->o_bs; p.o0_rs = a->o_rs; p.o0_hs = a->o_hs;
    return launch_tail(attn_tail_q_kernel<0>, p, a->Sk, stream);
}

int attention_tail_bwd(const MicoAttnArgs* a, cudaStream_t stream) {
    int rc = MICO_OK;
    if (m_tail_rows(a->Sq)) {
        TailParams p = base_params(a);
        p.rows = m_tail_rows(a->Sq);
        p.row0 = a->Sq - p.rows;
        p.out0 = reinterpret_cast<__nv_bfloat16*>(a->dq); p.o0_bs = a->dq_bs; p.o0_rs = a->dq_rs; p.o0_hs = a->dq_hs;
        if ((rc = launch_tail(attn_tail_q_kernel<1>, p, a->Sk, stream))) return rc;
    }
    if (m_tail_rows(a->Sk) && a->kv_index == nullptr) {      // shared K/V entries: the tile kernels take the remainder keys
        TailParams p = base_params(a);
        p.rows = m_tail_rows(a->Sk);
        p.row0 = a->Sk - p.rows;
        p.out0 = reinterpret_cast<__nv_bfloat16*>(a->dk); p.o0_bs = a->dk_bs; p.o0_rs = a->dk_rs; p.o0_hs = a->dk_hs;
        p.out1 = reinterpret_cast<__nv_bfloat16*>(a->dv); p.o1_bs = a->dv_bs; p.o1_rs = a->dv_rs; p.o1_hs = a->dv_hs;
        if ((rc = launch_tail(attn_tail_kv_kernel, p, a->Sq, stream))) return rc;
    }
    return rc;
}

}  // namespace mico

extern "C" int mico_attention_dmask(const MicoAttnArgs* a, float* dmask, void* stream_) {
    using namespace mico;
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    MICO_CHECK_ARG(a && dmask && a->mask && a->q && a->k && a->v && a->dout && a->lse && a->delta);
    MICO_CHECK_ARG(a->dropout_p == 0.0f);   // the learnable-bias towers (Swin) run without attention dropout
    TailParams p = base_params(a);
    const size_t smem = (size_t)2 * (a->Sk + a->Sq) * a->D * sizeof(float);
    if (smem > 200 * 1024) {
        set_last_error(__FILE__, __LINE__, "attention bias gradient: window too large for the SIMT kernel (Sq+Sk)*D*8 > 200 KB");
        return MICO_ERR_UNSUPPORTED;
    }
    if (smem > 48 * 1024)
        MICO_CHECK_CUDA(cudaFuncSetAttribute(attn_dmask_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    ProfScope prof(kProfAttnBwd, 4.0 * a->B * a->H * (double)a->Sq * a->Sk * a->D, stream);
    attn_dmask_kernel<<<a->B * a->H, 256, smem, stream>>>(p, dmask);
    MICO_CHECK_CUDA(cudaGetLastError());
    count_launch();
    return MICO_OK;
}
