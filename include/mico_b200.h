/* mico_b200 -- C-ABI of the B200 (sm_100a) kernels behind the MiCo hot path.
 *
 * The reference (invictus717/MiCo) has no FFI: its boundary is the Python nn.Module surface
 * (SURVEY.md section 8b).  This header is the boundary *beneath* that surface: one entry point per
 * fused op that the reference dispatches to ATen/cuBLAS/cuDNN.  Every function
 *   - takes raw DEVICE pointers borrowed for the call (no allocation inside, no torch types),
 *   - is stream-ordered on `stream` (a cudaStream_t passed as void*), re-entrant, never throws,
 *   - returns 0 on success or a negative MICO_ERR_* code; mico_last_error() gives the message.
 * Each declaration cites the reference call site (file:line under /root/reference) it replaces.
 */
#ifndef MICO_B200_H_
#define MICO_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MICO_OK 0
#define MICO_ERR_INVALID_ARG (-1)
#define MICO_ERR_CUDA (-2)
#define MICO_ERR_UNSUPPORTED (-3)
#define MICO_ERR_DRIVER (-4)

/* activation selector for mico_gemm_bf16 epilogues */
#define MICO_ACT_NONE 0
#define MICO_ACT_GELU 1          /* exact erf GELU: nn.GELU (eva_vit_model.py:173), ACT2FN["gelu"] (bert.py:354) */
#define MICO_ACT_QUICK_GELU 2    /* x*sigmoid(1.702x): model/clip/clip.py:168-170 */
#define MICO_ACT_GELU_BWD 3      /* out = acc * gelu'(aux_in) */
#define MICO_ACT_QUICK_GELU_BWD 4

int mico_version(void);
const char* mico_last_error(void);
/* number of kernels this library has launched since load / since the last reset (bench "gpu_launches") */
int64_t mico_launch_count(void);
void mico_reset_launch_count(void);

/* ---------------------------------------------------------------------------------------------
 * K3  Linear layers and every other dense contraction on the path (tcgen05 + TMA + TMEM).
 *     out[M,N] = epilogue( alpha * sum_k A(m,k) * B(n,k) )
 * Replaces F.linear / nn.Linear / torch.matmul at: eva_vit_model.py:310 (qkv), :363 (proj),
 * :191 (fc1), :197 (fc2), :446 (patch-embed conv as GEMM); bert.py:196-209, 293, 357, 370, 601, 607;
 * mico.py:41, 51-52, 400-403; and their autograd backward (dgrad / wgrad).
 *
 * Operand layouts (bf16, device):
 *   a_mn_major == 0 : A is [M rows][K] with row pitch lda (elements)       ("K-major", activations x)
 *   a_mn_major == 1 : A is [K rows][M] with row pitch lda                  (wgrad: dY^T without a transpose)
 *   b_mn_major == 0 : B is [N rows][K] with row pitch ldb                  (nn.Linear weight [out,in])
 *   b_mn_major == 1 : B is [K rows][N] with row pitch ldb                  (dgrad against W, wgrad against x)
 * Pitches must be multiples of 8 elements (16 bytes, TMA); base pointers 16-byte aligned.
 * Epilogue, in order:  v = alpha*acc;  v += bias[n];  act (GELU...) with optional aux_out = pre-activation;
 *   v *= row_scale[m / rows_per_group] (DropPath: eva_vit_model.py:121-138);  v += residual[m,n];
 *   if accumulate: v += out[m,n] (fp32 out only);  store as fp32 or bf16.
 * ------------------------------------------------------------------------------------------- */
typedef struct MicoGemmArgs {
    const void* a;        int64_t lda;  int32_t a_mn_major;
    const void* b;        int64_t ldb;  int32_t b_mn_major;
    int32_t M, N, K;
    void* out;            int64_t ldo;  int32_t out_fp32;
    const float* bias;                     /* [N] or NULL */
    const float* residual; int64_t ldr;    /* fp32 [M][ldr] or NULL */
    const float* row_scale; int32_t rows_per_group;  /* fp32 [ceil(M/rows_per_group)] or NULL */
    int32_t act;                           /* MICO_ACT_* */
    void* aux_out;        int64_t ld_aux_out;  /* bf16 [M][N]: value before the activation, or NULL */
    const void* aux_in;   int64_t ld_aux_in;   /* bf16 [M][N]: pre-activation for *_BWD, else NULL */
    int32_t accumulate;                    /* 1: out += v (requires out_fp32) */
    float alpha;
    /* optional output-row remap (patch-embed: 256 patch rows -> tokens 1..256 of a 257-token sample,
     * eva_vit_model.py:613-619): out_row = (m / remap_gin) * remap_gout + m % remap_gin + remap_off.
     * residual_bcast != 0: the residual row is (m % remap_gin) + remap_off (pos_embed broadcast over batch). */
    int32_t remap_gin, remap_gout, remap_off, residual_bcast;
} MicoGemmArgs;

int mico_gemm_bf16(const MicoGemmArgs* args, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MICO_B200_H_ */
