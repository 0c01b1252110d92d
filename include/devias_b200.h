/*
 * devias_b200 -- C-ABI of the B200-native DEVIAS hot path (sm_100a).
 *
 * The reference (KHU-VLL/DEVIAS) is pure PyTorch and has no FFI of its own; its "operator interface"
 * for this path is the set of torch library calls made by model/modeling_slot.py and the agg_block modules
 * (SURVEY.md section 2.2).  Each entry point below replaces one of those call sites and cites it.
 * All pointers are DEVICE pointers owned by the caller (PyTorch); no entry point allocates, synchronises
 * or touches the host copy of the data.  `stream` is a cudaStream_t passed as void*.
 * Every function returns 0 on success, DEVIAS_ERR_* otherwise; devias_last_error() gives the detail.
 * There is no CPU fallback anywhere behind this ABI.
 */
#ifndef DEVIAS_B200_H_
#define DEVIAS_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DEVIAS_OK 0
#define DEVIAS_ERR_ARG 1
#define DEVIAS_ERR_CUDA 2
#define DEVIAS_ERR_UNSUPPORTED 3

/* library / device introspection */
int devias_abi_version(void);
const char* devias_last_error(void);
/* number of kernels this library has launched since load (monotonic; used for bench.py's gpu_launches) */
int64_t devias_launch_count(void);

/* Optional per-kernel timing for bench.py's roofline leg: between begin/end every launch of the instrumented kernel
 * families is bracketed by CUDA events on its own stream; end() synchronises and returns the summed device time (ms),
 * the summed algorithmic work (FLOPs for GEMM/attention kinds, bytes for streaming kinds) and the launch count. */
#define DEVIAS_PROF_GEMM 0
#define DEVIAS_PROF_ATTN 1
#define DEVIAS_PROF_SLOT 2
#define DEVIAS_PROF_NORM 3
int devias_profile_begin(void);
/* same, but only launches made inside a CUDA stream capture are bracketed (eager warm-up launches are ignored) */
int devias_profile_begin_capture(void);
/* stop bracketing new launches; the records made so far stay readable */
int devias_profile_pause(void);
/* When the launches were made inside a CUDA stream capture, the brackets are external event-record nodes of the captured graph:
 * after every replay, end() returns the times of that replay (kernels timed INSIDE the replayed step).  It may be called
 * repeatedly; it never discards the records (begin() does). */
int devias_profile_end(int kind, double* total_ms, double* total_work, int64_t* launches);

/* ---- GEMM on tcgen05/TMEM fed by TMA ------------------------------------------------------------
 * D[m,n] = sum_k A[m,k] * B[n,k]  (bf16 operands, fp32 accumulation in tensor memory) + fused epilogue.
 * Replaces F.linear / nn.Linear / Conv3d-as-GEMM and their autograd dgrad/wgrad:
 *   model/modeling_slot.py:101 (qkv), :113 (proj), :61,:65 (fc1, fc2), :176 (patch embed).
 * Operand layouts: *_mn_major = 0: the operand is stored [rows = m or n][cols = k], k contiguous (ld = row stride);
 *                  *_mn_major = 1: stored [rows = k][cols = m or n] (a transposed view; used by wgrad/dgrad).
 * Epilogues (DEVIAS_EPI_*):
 *   STORE_BF16 : out_bf16[m,n]  = acc + bias[n]
 *   STORE_F32  : out_f32[m,n]   = acc + bias[n]
 *   GELU_BF16  : out_bf16 = acc + bias[n] (pre-activation), out2_bf16 = GELU_erf(acc + bias[n])      (modeling_slot.py:61-62)
 *   DGELU_BF16 : out_bf16 = acc * GELU'(aux_bf16[m,n])                                               (backward of :62)
 *   RESID_F32  : out_f32  = aux_f32[m % aux_row_mod (or m), n] + row_scale[m / rows_per_scale] * (acc + bias[n])
 *                (residual add :150-151 with per-sample drop-path scale; patch-embed + sin-cos table :354-355)
 *   ATOMIC_F32 : out_f32 += acc  via red.global.add (split-K weight gradients)
 * bias / row_scale may be NULL.  n % 32 == 0, k % 8 == 0, ld* % 8 == 0 required.
 */
#define DEVIAS_EPI_STORE_BF16 0
#define DEVIAS_EPI_STORE_F32 1
#define DEVIAS_EPI_GELU_BF16 2
#define DEVIAS_EPI_DGELU_BF16 3
#define DEVIAS_EPI_RESID_F32 4
#define DEVIAS_EPI_ATOMIC_F32 5
int devias_gemm_bf16(const void* a, int64_t lda, int a_mn_major, const void* b, int64_t ldb, int b_mn_major, int m, int n,
                     int k, int epilogue, void* out, int64_t ldo, void* out2, int64_t ldo2, const float* bias,
                     const void* aux, int64_t ldaux, int aux_row_mod, const float* row_scale, int rows_per_scale,
                     int split_k, void* stream);

/* ---- fused softmax attention of the encoder (head_dim 64) ------------------------------------------
 * qkv: packed bf16 [batch*seq, 3*heads*64] exactly as produced by the qkv GEMM (q | k | v column blocks, head-major inside
 * each).  out: bf16 [batch*seq, heads*64].  out = softmax(scale * q k^T) v per (clip, head); replaces
 * model/modeling_slot.py:102-112 (q*scale, q@k^T, softmax, attn@v, transpose/reshape) without materialising the
 * [12, seq, seq] probabilities.  lse2: fp32 [batch, heads, seq_pad] (seq_pad = seq rounded up to 128), log2-domain
 * log-sum-exp kept for the backward (may be NULL for inference).
 * Backward: dqkv bf16 [batch*seq, 3*heads*64] from dout.  Caller-provided scratch: aug_ws bf16 [batch*heads*seq_pad*16]
 * (the log-sum-exp / row-term operand blocks the kernel's extra MMA k-step reads) and dq_ws fp32 [batch*seq*heads*64]
 * (zeroed by the call, accumulated with red.global.add, converted into dqkv at the end). */
int devias_flash_attn_fwd(const void* qkv, void* out, float* lse2, int batch, int seq, int heads, int head_dim, float scale,
                          void* stream);
int devias_flash_attn_bwd(const void* qkv, const void* out, const void* dout, const float* lse2, void* dqkv, void* aug_ws,
                          float* dq_ws, int batch, int seq, int heads, int head_dim, float scale, void* stream);

/* ---- streaming slot attention (folded form; devias_b200/slot_attention.py) ----------------------------
 * One pass over the context tokens of every clip, replacing per layer: LayerNorm(context) + to_k + to_v + q k^T + slot-axis
 * softmax + token-axis renormalisation + attn v of agg_block/attention.py:32-40,120-141.
 *   tokens fp32 [batch, n_tokens, 768]; g fp32 [batch, 4*S, 768]; G, c0 fp32 [batch, 4*S]   (sh = head*S + slot)
 *   U [batch, 4*S, 768], m, A [batch, 4*S] are ACCUMULATED (+=; pass zero-filled buffers)
 *   attn [batch, 4*S, n_tokens] (= the reference's sim_distill in '(b h) s n' order) or NULL
 *   mu, rstd [batch, n_tokens]: LayerNorm statistics of the tokens, written when non-NULL (needed by the backward). */
int devias_slot_stream_fwd(const float* tokens, const float* g, const float* G, const float* c0, float* U, float* m, float* A,
                           float* attn, float* mu, float* rstd, int batch, int n_tokens, int dim, int num_slots, float eps,
                           void* stream);

/* Same contract for BF16 context tokens [batch, n_tokens, 768] (BASELINE config 5 "fp32 and bf16"): both contractions run on
 * tcgen05 straight off the TMA-landed token tile (g and the softmax weights are rounded to bf16, accumulation in fp32, LayerNorm
 * statistics / softmax in fp32); S in {2, 4, 8}.  Replaces agg_block/attention.py:32-40,120-141 when the model runs in bf16. */
int devias_slot_stream_fwd_bf16(const void* tokens, const float* g, const float* G, const float* c0, float* U, float* m, float* A,
                                float* attn, float* mu, float* rstd, int batch, int n_tokens, int dim, int num_slots, float eps,
                                void* stream);

/* Backward of devias_slot_stream_fwd.  attn / mu / rstd are the forward's outputs; dU, dm, dA (and optionally dattn) the
 * upstream gradients.  Writes dtokens [batch, n_tokens, 768] (accumulate_dtokens != 0: +=, used to sum the layers of the
 * aggregation block in place) and ACCUMULATES dg [batch, 4*S, 768], dG, dc0 [batch, 4*S].  S in {2, 4, 8} (S = 8 runs as two
 * passes over two heads each). */
int devias_slot_stream_bwd(const float* tokens, const float* mu, const float* rstd, const float* g, const float* G,
                           const float* attn, const float* dU, const float* dm, const float* dA, const float* dattn,
                           float* dtokens, int accumulate_dtokens, float* dg, float* dG, float* dc0, int batch, int n_tokens,
                           int dim, int num_slots, void* stream);

/* Backward of devias_slot_stream_fwd_bf16 (same arguments as devias_slot_stream_bwd with BF16 tokens): the three contractions
 * (dots with [g; dU], dg, dtokens) run on tcgen05 off the token tile and one bf16 image of [g; dU]; per-token coefficient math,
 * accumulation and the token gradient [batch, n_tokens, 768] stay fp32. */
int devias_slot_stream_bwd_bf16(const void* tokens, const float* mu, const float* rstd, const float* g, const float* G,
                                const float* attn, const float* dU, const float* dm, const float* dA, const float* dattn,
                                float* dtokens, int accumulate_dtokens, float* dg, float* dG, float* dc0, int batch,
                                int n_tokens, int dim, int num_slots, void* stream);

/* dtype ids for entry points that accept several input element types */
#define DEVIAS_DTYPE_F32 0
#define DEVIAS_DTYPE_BF16 1
#define DEVIAS_DTYPE_F16 2

/* ---- LayerNorm (dim = 768) ---------------------------------------------------------------------
 * Forward: y = (x - mean) * rstd * gamma + beta, x fp32 [rows, dim]; y bf16 (y_is_bf16) or fp32; mean/rstd fp32 [rows]
 * (may be NULL).  Replaces nn.LayerNorm at model/modeling_slot.py:126,132 (eps 1e-6), :373 and
 * agg_block/attention.py:29-30 (eps 1e-5).
 * Backward: dx = d_resid + dLN(dy) (fp32, optional) and/or its bf16 copy; dgamma/dbeta/dx_colsum (fp32 [dim], optional)
 * are ACCUMULATED (+=).  d_resid may be NULL; dx may alias d_resid.  dx_colsum = column sums of dx, i.e. the bias
 * gradient of the linear layer that produced the residual branch (Block.forward, model/modeling_slot.py:150-151).
 * bf16_row_scale (optional, [rows / rows_per_scale]): the per-sample drop-path factor (modeling_slot.py:36-47) of the branch
 * that consumes this gradient next; it multiplies the bf16 copy and dx_colsum only (dx itself stays unscaled). */
int devias_layernorm_fwd(const float* x, const float* gamma, const float* beta, void* y, int y_is_bf16, float* mean,
                         float* rstd, int rows, int dim, float eps, void* stream);
int devias_layernorm_bwd(const void* dy, int dy_is_bf16, const float* x, const float* mean, const float* rstd,
                         const float* gamma, const float* d_resid, float* dx, void* dx_bf16, float* dgamma, float* dbeta,
                         float* dx_colsum, const float* bf16_row_scale, int rows_per_scale, int rows, int dim, void* stream);
/* out[c] += sum_r a[r, c]  (bf16 [rows, cols] with row stride lda) -- bias gradients of qkv / fc1 */
int devias_colsum_bf16(const void* a, int64_t lda, int rows, int cols, float* out, void* stream);

/* ---- operand preparation ------------------------------------------------------------------------
 * cast: fp32 master weights -> the bf16 copies the tensor-core kernels read (one launch over the flat weight arena).
 * patchify: clip [B, C, T, H, W] (f32/bf16/f16) -> bf16 [B*(T/2)*(H/16)*(W/16), C*512] tube-patch rows so that
 * Conv3d(k=s=(2,16,16)) (model/modeling_slot.py:167-176) becomes devias_gemm_bf16 with the RESID_F32 epilogue adding the
 * bias and the sin-cos table (:354-355).  row = t*196 + h*14 + w, col = c*512 + dt*256 + dy*16 + dx. */
int devias_cast_f32_bf16(const float* in, void* out, int64_t n, void* stream);
/* bf16 -> fp32 (tokens handed to the fp32 slot-attention backward when the context arrives in bf16) */
int devias_cast_bf16_f32(const void* in, float* out, int64_t n, void* stream);
/* Tube patch embedding as an implicit GEMM (model/modeling_slot.py:167-177 Conv3d(k = s = (2,16,16)) + flatten/transpose, bias, and
 * the position table of :354-355 in the epilogue): out fp32 [B*1568, 768] = im2col(clip) W^T + bias + pos[token % 1568].
 * The A operand is fetched by a 5-D TMA box straight from the NCTHW fp32 clip [B, 3, 16, 224, 224] (no patch matrix, no
 * conversion pass; kind::tf32 MMAs on the fp32 data); weight = the fp32 Conv3d weight [768, 3*2*16*16]; pos fp32 [1568, 768]. */
int devias_patch_embed_fwd(const float* clip, const float* weight, const float* bias, const float* pos, float* out, int batch,
                           int chans, int frames, int height, int width, int dim, void* stream);
/* out_bf16[r,:] = bf16(in_f32[r,:] * row_scale[r / rows_per_scale]) -- gradient entering a drop-path branch (modeling_slot.py:36-47) */
int devias_scale_rows_cast(const float* in, void* out, int rows, int cols, const float* row_scale, int rows_per_scale,
                           void* stream);
int devias_patchify(const void* clip, int clip_dtype, void* out, int batch, int chans, int frames, int height, int width,
                    void* stream);
/* ---- slot-side glue around the streaming kernel (devias_b200/slot_attention.py), rows = batch * 4 * slots, dim = 768 --------
 * fold : g = scale * qt * gamma,  G[r] = sum_c g[r, c],  c0[r] = scale * sum_c qt[r, c] * beta[c]   (qt = Wk_h^T q: the key projection
 *        and the context LayerNorm's affine folded onto the slot queries, agg_block/attention.py:35-38,121-131)
 * ctx  : cbar = (gamma * (U - m) + beta * A) / (A + eps)   (token-axis renormalised context before to_v, :132-136)
 * The backward entry points ACCUMULATE (+=) into dgamma / dbeta. */
int devias_slot_fold_fwd(const float* qt, const float* gamma, const float* beta, float scale, float* g, float* G, float* c0, int rows,
                         int dim, void* stream);
int devias_slot_fold_bwd(const float* qt, const float* gamma, const float* beta, float scale, const float* dg, const float* dG,
                         const float* dc0, float* dqt, float* dgamma, float* dbeta, int rows, int dim, void* stream);
int devias_slot_ctx_fwd(const float* U, const float* m, const float* A, const float* gamma, const float* beta, float eps, float* cbar,
                        int rows, int dim, void* stream);
int devias_slot_ctx_bwd(const float* dcbar, const float* U, const float* m, const float* A, const float* gamma, const float* beta,
                        float eps, float* dU, float* dm, float* dA, float* dgamma, float* dbeta, int rows, int dim, void* stream);

/* ---- head: slot selection (model/modeling_slot.py:396-404; model/modeling_slot_fusion.py:376-386) -------------
 * logits [batch*slots, ld] fp32 (n_action + n_scene used columns) -> per clip the slot whose softmax row has the largest
 * action-class probability / scene-class probability (int64 [batch] each; first maximum wins, as torch.argmax). */
int devias_slot_select(const float* logits, int64_t ld, int batch, int slots, int n_action, int n_scene, long long* action_idx,
                       long long* scene_idx, void* stream);

/* diagnostic: the bare TMA token stream of the slot kernels (same tensor map / ring / tile split, no arithmetic) */
int devias_debug_token_stream(const float* tokens, int batch, int n_tokens, int stages, float* scratch, void* stream);

/* ---- slot-side products (fp32, M = clips x slots rows) -----------------------------------------------
 * The aggregation block's projections and the heads act on B*S rows only (agg_block/attention.py:120-141 to_q / to_out
 * and the folded to_k / to_v, :81-82 FeedForward; model/modeling_slot.py:390-410 head and mask predictor): every product
 * is one pass over the weight matrix.  Rows of x / y / a / b are addressed by a map of four int64 (host memory)
 *   {outer, inner, ld, batch}:  offset(m, z) = (m / inner) * outer + (m % inner) * ld + z * batch     (elements)
 * and z = 0..batch-1 selects the problem (slot head); weights advance by w_batch per problem.
 *   nt    : y[m, n]  = sum_k x[m, k] w[n, k] (+ bias[n])    w [N, K] row-major; K % 4 == 0
 *   nn    : y[m, n] += sum_k x[m, k] w[k, n]                w [K, N] row-major; y is accumulated (caller pre-fills it)
 *   outer : c[i, j]  = sum_m a[m, i] b[m, j]                c [I, J] row-major, J % 4 == 0; colsum[i] = sum_m a[m, i] if
 *                                                           colsum != NULL (the bias gradient when a = dY);
 *                                                           accumulate != 0: c and colsum are added to (+=) -- weight-tied layers
 *                                                           summing into the gradient arena */
int devias_skinny_nt(const float* x, const int64_t* x_map, const float* w, int64_t w_batch, const float* bias, float* y,
                     const int64_t* y_map, int M, int N, int K, int batch, void* stream);
int devias_skinny_nn(const float* x, const int64_t* x_map, const float* w, int64_t w_batch, float* y, const int64_t* y_map,
                     int M, int N, int K, int batch, void* stream);
int devias_skinny_outer(const float* a, const int64_t* a_map, const float* b, const int64_t* b_map, float* c, int64_t c_batch,
                        float* colsum, int64_t colsum_batch, int M, int I, int J, int batch, int accumulate, void* stream);

/* ---- optimizer update over the flat parameter arena --------------------------------------------------------------
 * AdamW exactly as torch.optim.AdamW (the optimizer utils/optim_factory.py:94-178 builds; stepped at
 * engine/engine_for_slot.py:147-166), applied in ONE pass over flat fp32 arenas {param, grad, exp_avg, exp_avg_sq} of n
 * elements (n % 8 == 0): the pass also refreshes the bf16 copies the tensor-core kernels read (param_bf16, may be NULL) and
 * zero-fills the gradient (zero_grad != 0), replacing the multi-tensor optimizer launches, the weight cast and the memset.
 * Parameters occupy segments of 8-element granules: seg_start[0..n_seg] (device int32, ascending, seg_start[n_seg] = n / 8),
 * seg_group[s] = hyper-parameter group of segment s.  All hyper-parameters are read from DEVICE memory (so a captured CUDA
 * graph follows the learning-rate / weight-decay schedule written every iteration at engine/engine_for_slot.py:91-97):
 *   hyper[0] = 1 - beta1^t, [1] = 1 - beta2^t, [2] = beta1, [3] = beta2, [4] = eps, [5] = max_norm (<= 0: off),
 *   [6] = gradient pre-scale, [7] unused, then per group g: [8 + 2g] = lr, [9 + 2g] = weight_decay.
 * grad_sumsq (device fp32 scalar, may be NULL): sum of squares of the whole gradient arena (devias_sumsq_f32); with
 * max_norm > 0 the gradient is scaled by min(1, max_norm / (|pre-scale| * sqrt(sumsq) + 1e-6)) = torch clip_grad_norm_.
 * grad_bf16 (may be NULL): when given, the gradient VALUES are read from this bf16 arena of n elements (the buffer a bf16 gradient
 * all-reduce left its result in) instead of `grad`; `grad` is still the arena that gets zero-filled. */
int devias_sumsq_f32(const float* x, int64_t n, float* out, void* stream);
int devias_adamw_arena(float* param, float* grad, float* exp_avg, float* exp_avg_sq, void* param_bf16,
                       const int32_t* seg_start, const int32_t* seg_group, int n_seg, const float* hyper,
                       const float* grad_sumsq, int64_t n, int zero_grad, const void* grad_bf16, void* stream);

/* ---- the training objective ('matching' branch of utils/loss/train_loss.py:85-187) in one launch per direction ----------------
 * Per clip: softmax of the S slot rows of slots_head [B*S, width = n_action + scene classes], assignment of distinct (action, scene)
 * slots minimising -p[i, target] - p[j, scene_target] (:112-122; scene_target = n_action + argmax teacher), then
 *   out6[0] action  = CE(head[i], target)                                   out6[1] scene = w_scene * KL(teacher_full || head[j]) / width
 *   out6[2] cosine  = mean_{i != j} <slots_i, slots_j> (normalised)                          (scene_ce != 0: CE(head[j], scene_target))
 *   out6[3] maskpred = w_maskpred * BCEwithLogits(maskp[i], fg)             out6[4] distill = w_distill * MSE(mean_heads attn[i], fgf)
 *   out6[5] total; all summed over the batch / batch.  slot_idx int64 [B, 2] = (i, j).  teacher_full = [var x n_action | teacher]
 * with var_scalar a DEVICE scalar = min over the whole batch of the teacher logits - 1 (:103).
 * attn fp32 [B*heads, S, n_tokens], maskp [B*S, n_patches], slots [B*S, dim], teacher [B, width - n_action], fg [B, n_patches],
 * fgf [B, n_tokens], target int64 [B].  The backward writes the FULL gradients of out6[5] * grad_total[0] (zeros included). */
int devias_train_loss_fwd(const float* head, const float* attn, const float* maskp, const float* slots, const int64_t* target,
                          const float* teacher, const float* var_scalar, const float* fg, const float* fgf, int batch,
                          int slots_per_clip, int width, int n_action, int heads, int n_tokens, int n_patches, int dim, int scene_ce,
                          float w_scene, float w_maskpred, float w_distill, float* out6, int64_t* slot_idx, void* stream);
int devias_train_loss_bwd(const float* head, const float* attn, const float* maskp, const float* slots, const int64_t* target,
                          const float* teacher, const float* var_scalar, const float* fg, const float* fgf, int batch,
                          int slots_per_clip, int width, int n_action, int heads, int n_tokens, int n_patches, int dim, int scene_ce,
                          float w_scene, float w_maskpred, float w_distill, const float* grad_total, float* dhead, float* dattn,
                          float* dmaskp, float* dslots, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* DEVIAS_B200_H_ */
