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
#define MICO_ACT_GELU_SAVE_GRAD 5       /* out = gelu(v), aux_out = gelu'(v): backward is one multiply (MUL_AUX) */
#define MICO_ACT_QUICK_GELU_SAVE_GRAD 6
#define MICO_ACT_MUL_AUX 7              /* out = acc * aux_in */

int mico_version(void);
/* Persistent kernels (GEMM, attention, LayerNorm forward) launch one CTA per SM.  When a communication library runs its own
 * kernels concurrently (the overlapped NCCL gradient all-reduce of data-parallel training, pipeline.py:93-99), leave it n SMs
 * (even, a whole TPC each) so that its CTAs never displace one of ours into a second wave.  Default 0. */
int mico_set_reserved_sms(int n);
const char* mico_last_error(void);
/* number of kernels this library has launched since load / since the last reset (bench "gpu_launches") */
int64_t mico_launch_count(void);
void mico_reset_launch_count(void);
/* Device timing per kernel family for bench.py's roofline: while enabled, every entry point brackets its launches
 * with a cudaEvent pair on the launching stream.  mico_profile_collect synchronises the device and returns, per
 * family (MICO_PROF_*), the summed milliseconds, the summed algorithmic work (FLOPs for GEMM/attention, bytes for
 * the HBM-bound families) and the number of calls.  mico_profile_enable(0/1) also clears the records. */
#define MICO_PROF_GEMM 0
#define MICO_PROF_ATTN_FWD 1
#define MICO_PROF_ATTN_BWD 2
#define MICO_PROF_LN_FWD 3
#define MICO_PROF_LN_BWD 4
#define MICO_PROF_OTHER 5
#define MICO_PROF_KINDS 6
int mico_profile_enable(int on);
int mico_profile_collect(double* ms, double* work, int64_t* count, int nkinds);

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
    /* Weight gradients (a_mn_major = b_mn_major = 1, fp32 out, no bias / residual / accumulate): asum_out (fp32 [M]) also
     * receives sum_k A(m,k) -- with A = dY^T the layer's BIAS gradient (the column sum of dY over tokens, what
     * eva_vit_model.py:191,310's nn.Linear bias receives in autograd) -- from the same pass over dY: the last N tile issues 32
     * more MMA columns against a tile of ones.  `ones`: bf16 [>= K rows][64] of 1.0, row pitch 64 elements.  Supported when
     * the last 256-wide N tile has room (N % 256 in (96, 160], M >= 256: the ViT-g tower's N = 1408); otherwise the call
     * returns MICO_ERR_UNSUPPORTED and the caller sums the columns of dY with mico_colsum_bf16. */
    float* asum_out;
    const void* ones;
} MicoGemmArgs;

int mico_gemm_bf16(const MicoGemmArgs* args, void* stream);

/* Host-side view of the work-unit plan mico_gemm_bf16 uses for a launch (no GPU needed; tests and tuning scripts).  A launch
 * covers num_tiles output tiles (num_n of them per M group, bn MMA columns each, the last one of every group n_last), each
 * with num_kb K blocks, on `slots` persistent CTAs / CTA pairs.  Returns the split-K factor (1 unless allow_split), the number
 * of rounds, the planned makespan in (MMA column x K block) units and, if table != NULL, the [slots x rounds] unit table
 * (-1 padded): slot s runs units table[s * rounds + 0..], unit u = (tile u % num_tiles, K part u / num_tiles). */
int mico_gemm_plan(int num_tiles, int num_n, int bn, int n_last, int num_kb, int slots, int allow_split, int* ks, int* rounds,
                   double* makespan, int* table, int table_cap);

/* ---------------------------------------------------------------------------------------------
 * K4  Fused attention (flash-style; tcgen05 QK^T and PV, online softmax, no score matrix in HBM).
 *     O = softmax(scale * Q K^T + mask) V          per (batch b, head h)
 * Replaces eva_vit_model.py:340-361 (q*scale; q@k^T; softmax; @v), bert.py:233-277 (scores/sqrt(d) +
 * additive mask; softmax; @v; self- and cross-attention), transformer.py:121-130, clip.py:185-189.
 * Tensors are bf16 with arbitrary (16-byte aligned) strides so that the fused QKV GEMM output is read in
 * place: element (b, i, h, d) lives at ptr + b*bs + i*rs + h*hs + d (strides in elements, multiples of 8).
 * mask: additive fp32 (the reference's (1-m)*-10000 masks, bert.py:697-781) at mask + b*mask_bs + i*mask_qs + j
 * (mask_qs = 0 for a per-key padding mask), or NULL.   lse: [B,H,Sq] fp32 log-sum-exp saved for backward.
 * head_dim D: multiple of 8, <= 128 forward, <= 96 backward (ViT-g 88, BERT/CLIP 64, Swin 32).
 * Backward (mico_attention_bwd) recomputes P from Q,K and lse; needs delta[b,h,i] = sum_d dO*O
 * (mico_attention_bwd_delta) and writes dQ, dK, dV with the same stride convention.
 * ------------------------------------------------------------------------------------------- */
typedef struct MicoAttnArgs {
    const void* q; int64_t q_bs, q_rs, q_hs;
    const void* k; int64_t k_bs, k_rs, k_hs;
    const void* v; int64_t v_bs, v_rs, v_hs;
    void* o;       int64_t o_bs, o_rs, o_hs;      /* forward: output; backward: forward output (input) */
    float* lse;                                    /* [B,H,Sq] */
    const float* mask; int64_t mask_bs, mask_qs;
    int32_t B, H, Sq, Sk, D;
    float scale;
    /* backward only */
    const void* dout; int64_t do_bs, do_rs, do_hs;
    float* delta;                                  /* [B,H,Sq] workspace: rowsum(dO * O) */
    void* dq; int64_t dq_bs, dq_rs, dq_hs;
    void* dk; int64_t dk_bs, dk_rs, dk_hs;
    void* dv; int64_t dv_bs, dv_rs, dv_hs;
    /* additive bias that also depends on the head, shared by groups of batch entries -- Swin's relative position bias
     * + shifted-window mask (swin.py:135-147): the mask element is at
     *   mask + (mask_bmod ? b % mask_bmod : b) * mask_bs + h * mask_hs + i * mask_qs + j
     * (batch entry = window; mask_bmod = windows per image; 0 / 0 reproduces the per-batch mask above) */
    int64_t mask_hs; int32_t mask_bmod;
    /* attention-probability dropout (bert.py:243-247): P is multiplied by a counter-based Bernoulli(1-p)/(1-p) mask of
     * element ((b*H + h)*Sq + i)*Sk + j under `dropout_seed`; forward and backward regenerate it.  p = 0 disables. */
    float dropout_p; uint64_t dropout_seed;
    /* K/V shared by several query batch entries -- the fusion encoder's ITM (positive, hard-negative-text) and caption
     * sequences of one sample all cross-attend to the same visual tokens (data/model/vast.py:445-451, 504-507), so their K / V
     * projections are computed once: kv_index[b] (device int32, or NULL = identity) is the K/V batch entry that query entry b
     * reads; K, V, dK, dV have n_kv batch entries.  The backward pass also needs the inverse map in CSR form: K/V entry e is
     * read by query entries grp_list[grp_ptr[e] .. grp_ptr[e+1]) (device int32; every entry must have at least one reader). */
    const int32_t* kv_index; int32_t n_kv; const int32_t* grp_ptr; const int32_t* grp_list;
} MicoAttnArgs;

int mico_attention_fwd(const MicoAttnArgs* args, void* stream);
int mico_attention_bwd(const MicoAttnArgs* args, void* stream);
/* gradient of a learnable additive bias (Swin relative position bias, swin.py:135-139): dmask (same layout as mask, zeroed
 * by the caller) += P o (dO V^T - delta), summed over the batch entries that share a mask slice.  Call after
 * mico_attention_bwd with the same arguments (it needs lse and delta).  Small windows only. */
int mico_attention_dmask(const MicoAttnArgs* args, float* dmask, void* stream);

/* ---------------------------------------------------------------------------------------------
 * K2  LayerNorm (eva_vit_model.py:375,382,542 eps 1e-6; bert.py:92,290,368,583 and mico.py:49,400-403
 *     eps 1e-12; swin.py:212,218,329,565 eps 1e-5).  x is fp32 or bf16 [M,D]; y as bf16 and/or fp32.
 *     Backward: dx = [dres +] LN'(dy [+ dy2]); dy2 = optional second (bf16) upstream gradient, for a LayerNorm
 *     output that feeds both the next GEMM and the next residual add (post-LN BERT, bert.py:286-297); optional bf16 copy of dx scaled per row group (DropPath);
 *     dgamma/dbeta reduced deterministically through `workspace` (mico_layernorm_bwd_workspace bytes).
 *     dxb_colsum (optional, fp32 [D], 512 <= D <= 1536): column sums of the scaled output = the bias gradient of the
 *     upstream nn.Linear (eva_vit_model.py:197,363 backward), fused here instead of a separate pass over dx_bf16.
 * ------------------------------------------------------------------------------------------- */
int mico_layernorm_fwd(const void* x, int x_is_bf16, int64_t ldx, const float* gamma, const float* beta,
                       void* y_bf16, float* y_f32, int64_t ldy, float* mean, float* rstd, int M, int D,
                       float eps, void* stream);
size_t mico_layernorm_bwd_workspace(int M, int D);
int mico_layernorm_bwd(const void* dy, int dy_is_bf16, int64_t lddy, const void* dy2_bf16, int64_t lddy2,
                       const float* x, int64_t ldx,
                       const float* mean, const float* rstd, const float* gamma, const float* dres,
                       int64_t lddres, float* dx, int64_t lddx, void* dx_bf16, int64_t lddxb,
                       const float* row_scale, int rows_per_group, float* dgamma, float* dbeta,
                       int accumulate_param_grads, float* dxb_colsum, int M, int D, void* workspace, size_t ws_bytes,
                       void* stream);
/* Same, for a LayerNorm whose input was dropout(dense(x)) + residual (post-LN BERT, bert.py:293-296, 370-373): the bf16 copy
 * of dx (and dxb_colsum) is additionally multiplied by the hidden-dropout mask of the forward pass -- element (row, col) by
 * the multiplier mico_dropout gives flat index drop_site + row * D + col under (drop_p, drop_seed) -- i.e. it is the gradient
 * w.r.t. the dense layer's output and dxb_colsum its bias gradient; the fp32 dx (residual gradient) stays unmasked.  One pass
 * instead of LayerNorm backward + mico_dropout + mico_colsum_bf16.  fp32 dy, 512 <= D <= 1536, aligned rows, dx_bf16 required;
 * otherwise MICO_ERR_UNSUPPORTED. */
int mico_layernorm_bwd_dropout(const void* dy, int dy_is_bf16, int64_t lddy, const void* dy2_bf16, int64_t lddy2,
                               const float* x, int64_t ldx,
                               const float* mean, const float* rstd, const float* gamma, const float* dres,
                               int64_t lddres, float* dx, int64_t lddx, void* dx_bf16, int64_t lddxb,
                               const float* row_scale, int rows_per_group, float* dgamma, float* dbeta,
                               int accumulate_param_grads, float* dxb_colsum, int M, int D, void* workspace, size_t ws_bytes,
                               float drop_p, uint64_t drop_seed, uint64_t drop_site, void* stream);

/* ---------------------------------------------------------------------------------------------
 * HBM-bound helpers on the path (all vectorised 128-bit, grid sized in multiples of the SM count).
 * ------------------------------------------------------------------------------------------- */
/* fp32 master parameters -> bf16 GEMM operands (the reference relies on torch.autocast, pipeline.py:43) */
int mico_cast_f32_to_bf16(const float* src, void* dst, int64_t n, void* stream);
/* bf16 -> fp32: a gradient bucket reduced in bf16 back into the fp32 gradient buffer (data-parallel SUM of
 * data/utils/pipeline.py:93-99 at half the NVLink bytes; mico_b200/dp.py) */
int mico_cast_bf16_to_f32(const void* src, float* dst, int64_t n, void* stream);
/* x *= (*scale_dev if scale_dev else 1) * scale_host, in place (16-byte aligned fp32): gradients of a loss group that was
 * differentiated ahead of the outer backward pass are rescaled by the upstream scalar gradient -- the loss scale of
 * GradScaler, data/utils/pipeline.py:86-88, or 1 (mico_b200/train_step.py) */
int mico_scale_f32(float* x, const float* scale_dev, float scale_host, int64_t n, void* stream);
/* SURVEY 8(f).4 EVA02 towers.  Rotary position embedding (model/evaclip/rope.py:79-136, applied to q and k at
 * eva_vit_model.py:314-322): contiguous [B, T, H, d] tokens, token 0 (cls) passes through, the others are rotated pair-wise
 * with the interleaved tables cos / sin [T-1, d]; fp32 -> bf16 (forward: the attention operand) or bf16 -> fp32 with
 * inverse = 1 (the gradient of the rotation = its transpose). */
int mico_rope(const void* x, int x_is_bf16, void* y, int y_is_bf16, const float* cos_table, const float* sin_table, int B,
              int T, int H, int d, int inverse, void* stream);
/* SwiGLU (eva_vit_model.py:201-224): dg == NULL: out0 = silu(u1) * u2; else out0 = d u1, out1 = d u2. */
int mico_swiglu(const float* u1, const float* u2, const float* dg, float* out0, float* out1, int64_t n, void* stream);
/* same for a [rows, cols] matrix into a wider bf16 pitch ldd, zero-filling columns cols..ldd-1
 * (patch-embed weight (1408, 3*14*14=588) -> K padded to a 16-byte multiple, eva_vit_model.py:440) */
int mico_cast_f32_to_bf16_2d(const float* src, int64_t lds, int rows, int cols, void* dst, int64_t ldd, void* stream);
/* nn.Linear bias gradient: out[n] = sum_m x[m,n] (accumulate 0), out += sum (1), out -= sum (2); x bf16 */
size_t mico_colsum_workspace(int M, int N);
/* column sums of two column ranges of one bf16 matrix in one launch: logical columns [0,n0) -> out0, physical columns
 * [n0+gap, n0+gap+n1) -> out1 (the q and v thirds of the fused qkv gradient; k has no bias, eva_vit_model.py:307).
 * n0, gap, n1 multiples of 8; workspace as for mico_colsum_workspace(M, n0 + n1). */
int mico_colsum2_bf16(const void* x, int64_t ldx, int M, int n0, int gap, int n1, float* out0, float* out1,
                      void* workspace, size_t ws_bytes, void* stream);
int mico_colsum_bf16(const void* x, int64_t ldx, int M, int N, float* out, int accumulate, void* workspace,
                     size_t ws_bytes, void* stream);
/* d pos_embed / d cls_token: out[r] (+)= sum_b x[b*R + r] */
int mico_batch_sum_f32(const float* x, int B, int64_t R, float* out, int accumulate, void* stream);
/* K1 im2col for Conv2d(k=s=P) (eva_vit_model.py:440-447; swin.py:437-475): (B,C,H,W) fp32 -> bf16
 * [B*(H/P)*(W/P), Kpad], columns (c,ky,kx) zero-padded to Kpad; chan_stride=0 replicates one channel
 * (forward_audio_encoder's repeat(1,1,3,1,1), mico.py:139-143).  tokens_per_img > 0 lays image b's patches at
 * rows b*tokens_per_img + token_off ... and zero-fills the other rows of that image (token_off = 1 leaves the
 * cls slot of eva_vit_model.py:615-616 as a zero row so the token matrix is one dense [B*257, Kpad] operand). */
int mico_patchify(const float* img, int64_t img_stride, int64_t chan_stride, int B, int C, int H, int W, int P,
                  int Kpad, int tokens_per_img, int token_off, void* out, void* stream);
/* token 0 of every sample = cls_token + pos_embed[0] (eva_vit_model.py:615-619) */
int mico_cls_pos_row(const float* cls_token, const float* pos0, float* x, int64_t sample_stride, int B, int D,
                     void* stream);
/* K12 DropPath (eva_vit_model.py:121-138, rates linspace(0, 0.4, depth) :533): all per-sample multipliers of a tower
 * in one launch, out[l][j][b] = Bernoulli(1 - drop_prob[l]) / (1 - drop_prob[l]) for branch j in {attn, mlp};
 * counter-based Philox so a (seed, offset) pair reproduces the masks.  drop_prob: device fp32 [L]. */
int mico_drop_path_scales(const float* drop_prob, int L, int B, uint64_t seed, uint64_t offset, float* out,
                          void* stream);
/* y = bf16(x * row_scale[row / rows_per_group]) */
int mico_scale_cast_bf16(const float* x, int64_t ldx, const float* row_scale, int rows_per_group, void* y,
                         int64_t ldy, int M, int D, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Text head and losses (fp32 SIMT kernels; small or HBM-bound).
 * ------------------------------------------------------------------------------------------- */
/* K6 BertEmbeddings (bert.py:139-146): out[m] = word[ids[m]] + type[type_ids[m] or 0] + pos[pos_ids[m] or (m % S) + pos_offset];
 * the LayerNorm that follows is mico_layernorm_fwd.  Backward into the word table: dtable[ids[m]] += dx[m] (atomics). */
int mico_embedding_gather(const int64_t* ids, const int64_t* type_ids, const int64_t* pos_ids, int pos_offset,
                          const float* word, const float* pos, const float* type, float* out, int M, int S, int D, int V,
                          int P, int T, void* stream);
int mico_embedding_scatter_add(const float* dx, const int64_t* ids, float* dtable, int M, int D, int V, void* stream);
/* K7/K8 F.cross_entropy(logits[M,V], labels, ignore_index, label_smoothing), mean over non-ignored rows
 * (bert.py:1088-1090; vast.py:412-415 with label_smoothing 0.1; vast.py:456).  loss_and_count[0] = loss,
 * [1] = number of contributing rows.  Backward writes dlogits = grad[0] * dloss/dlogits as bf16 or fp32. */
int mico_cross_entropy_fwd(const void* logits, int logits_bf16, int64_t ld, const int64_t* labels, int64_t ignore_index,
                           float label_smoothing, float* row_loss, float* lse, float* loss_and_count, int M, int V,
                           void* stream);
int mico_cross_entropy_bwd(const void* logits, int logits_bf16, int64_t ld, const int64_t* labels, int64_t ignore_index,
                           float label_smoothing, const float* lse, const float* grad, const float* loss_and_count,
                           void* dlogits, int dlogits_bf16, int64_t ldd, int M, int V, void* stream);
/* K7, LM head + cross-entropy without the [M, 30522] logits (bert.py:606-608 decoder + bert.py:1084-1090 CrossEntropyLoss,
 * label_smoothing 0): the caller walks the vocabulary in chunks -- logits[:, col0 : col0+Vc] = decoder GEMM into a reusable
 * fp32 buffer -- and these entries keep an online log-sum-exp (run_max, run_sum) and the label's logit per row
 * (mico_ce_chunk_update; first != 0 on the first chunk), turn them into lse / row losses / (loss, n_valid)
 * (mico_ce_chunk_finalize), and in the backward pass turn a RECOMPUTED chunk of logits into its bf16 dlogits
 * = grad[0] / n_valid * (softmax - onehot) (mico_ce_chunk_grad).  mico_b200/bert.py:_LMHeadLossFn is the caller. */
int mico_ce_chunk_update(const float* logits, int64_t ld, int col0, int Vc, const int64_t* labels, float* run_max,
                         float* run_sum, float* label_logit, int M, int first, void* stream);
int mico_ce_chunk_finalize(const float* run_max, const float* run_sum, const float* label_logit, const int64_t* labels,
                           int64_t ignore_index, int V, float* row_loss, float* lse, float* loss_and_count, int M, void* stream);
int mico_ce_chunk_grad(const float* logits, int64_t ld, int col0, int Vc, const int64_t* labels, int64_t ignore_index, int V,
                       const float* lse, const float* grad, const float* loss_and_count, void* dlogits_bf16, int64_t ldd, int M,
                       void* stream);
/* F.normalize(x, dim=-1) (vast.py:225): y = x / max(||x||, eps); norm[m] saved for the backward */
int mico_l2norm_fwd(const float* x, float* y, float* norm, int M, int D, float eps, void* stream);
int mico_l2norm_bwd(const float* y, const float* dy, const float* norm, float* dx, int M, int D, void* stream);
/* Small fp32 GEMM with arbitrary operand strides: C[m,n] (+)= s * sum_k A(m,k) B(n,k) + bias[n],
 * A(m,k) = a[m*a_sm + k*a_sk], B(n,k) = b[n*b_sn + k*b_sk], s = alpha * (alpha_dev ? (alpha_recip ? 1/ *alpha_dev : *alpha_dev) : 1).
 * K8 contrastive logits sim = f . f_all^T / contra_temp (vast.py:405-408, temperature read on the device) and the
 * fp32 heads (Contra_head mico.py:36-41, contra_head_va/id/vs/vas :386-394, Match_head :44-52). */
int mico_sgemm_strided(const float* a, int64_t a_sm, int64_t a_sk, const float* b, int64_t b_sn, int64_t b_sk, float* c,
                       int64_t ldc, const float* bias, int M, int N, int K, float alpha, const float* alpha_dev,
                       int alpha_recip, int accumulate, void* stream);
/* exact-erf GELU on a small fp32 tensor (Match_head, mico.py:44-52): out = gelu(x), or out = dy * gelu'(x) when dy != NULL */
int mico_gelu_f32(const float* x, const float* dy, float* out, int64_t n, void* stream);
/* hidden-state dropout (bert.py:148, 294, 372): out = [res +] x * m, m = Bernoulli(1-p)/(1-p) of element
 * (site_offset + i) under `seed` (same counter-based generator as the attention dropout); the backward pass calls it
 * again on the gradient with the same (seed, site_offset).  x fp32 or bf16; out as fp32 and/or bf16. */
int mico_dropout(const void* x, int x_is_bf16, const float* res, float* out_f32, void* out_bf16, int64_t n, float p,
                 uint64_t seed, uint64_t site_offset, void* stream);
/* out[0] (+)= alpha * <a, b>  (d contra_temp) */
int mico_dot_f32(const float* a, const float* b, int64_t n, float alpha, float* out, int accumulate, void* stream);

/* K11 Kaldi-compatible log-mel filterbank (model/audioprocessor.py:38-46: waveform * 2^15 -> torchaudio
 * compliance.kaldi.fbank(num_mel_bins, 16 kHz, frame_length 25 ms, frame_shift 10 ms; defaults: snip_edges, DC removal,
 * pre-emphasis 0.97, povey window, 512-point power spectrum, mel 20 Hz..Nyquist, log) and the normalisation
 * (x - norm_sub) * norm_mul of :47-48).  wave: fp32 [n_clips][clip_stride]; window: fp32 [frame_len]; mel: fp32
 * [num_mel][257]; out: fp32 [n_clips][n_frames][num_mel], n_frames = 1 + (n_samples - frame_len) / frame_shift. */
int mico_fbank(const float* wave, int64_t clip_stride, int n_clips, int n_samples, int frame_len, int frame_shift,
               const float* window, const float* mel, int num_mel, float in_scale, float preemph, float log_floor,
               float norm_sub, float norm_mul, float* out, int64_t out_clip_stride, void* stream);

/* ---------------------------------------------------------------------------------------------
 * SURVEY 8(f).4  Image / video-frame preprocessing (input side of the path): ToTensor -> Resize((Ho,Wo), bilinear,
 *     align_corners=False) -> Normalize(mean, std) of model/imageprocessor.py:24-29,52-56 and model/videoprocessor.py:36-41
 *     in one kernel.  src: n images, uint8 HWC (src_is_u8_hwc = 1, the decoder's layout; scaled by 1/255) or fp32 CHW;
 *     img_stride in elements.  dst: fp32 [n, C, Ho, Wo].  mean / std: HOST arrays of C floats.  antialias = 1 follows
 *     torchvision >= 0.17 (ATen's separable triangle filter), 0 the reference's pinned torchvision 0.15.2.
 * ------------------------------------------------------------------------------------------- */
int mico_resize_normalize(const void* src, int src_is_u8_hwc, int n, int C, int H, int W, int64_t img_stride, float* dst,
                          int Ho, int Wo, const float* mean_host, const float* std_host, int antialias, void* stream);

/* ---------------------------------------------------------------------------------------------
 * SURVEY 8(f).1  Fused multi-tensor AdamW with the decoupled weight-decay "fix" -- replaces AdamW.step of
 *     data/utils/build_optimizer.py:136-196 (one launch for every parameter of every group instead of ~8 torch
 *     kernels per tensor):  m,v moment updates; p -= step_size * m/(sqrt(v)+eps); then p -= lr*wd*p.
 *     step_size = lr * sqrt(1-beta2^t)/(1-beta1^t) is computed by the host (it knows t); lr is the group's
 *     scheduled rate (pipeline.py:75-78).  `tensors`, `chunk_tensor`, `chunk_index` are DEVICE arrays: work item c
 *     updates elements [chunk_index[c]*chunk_elems, +chunk_elems) of tensors[chunk_tensor[c]].  p_bf16 (optional)
 *     receives the bf16 copy of the updated parameter (the GEMM operand the towers otherwise re-cast every step).
 *     grad_scale multiplies every gradient first (1/world for an averaged DP sum, 1/loss-scale, or 1).
 * ------------------------------------------------------------------------------------------- */
typedef struct MicoAdamTensor {
    float* p; const float* g; float* m; float* v;   /* fp32 [n] */
    void* p_bf16;                                    /* bf16 [n] or NULL */
    int64_t n;
    int32_t group;                                   /* index into the hyper-parameter table */
    int32_t reserved;
} MicoAdamTensor;
typedef struct MicoAdamHyper { float lr, step_size, weight_decay, beta1, beta2, eps; } MicoAdamHyper;
int mico_adamw_multi(const MicoAdamTensor* tensors_dev, const int32_t* chunk_tensor_dev, const int32_t* chunk_index_dev,
                     int n_chunks, int chunk_elems, const MicoAdamHyper* hyper_host, int n_groups, float grad_scale,
                     double total_elems, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MICO_B200_H_ */
