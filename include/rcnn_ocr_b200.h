/*
 * rcnn_ocr_b200.h -- C ABI of the B200-native RCNN-OCR sequence-recognition hot path.
 *
 * The upstream reference (sherstpasha/RCNN-OCR) is pure Python/PyTorch and has no FFI
 * layer of its own; its boundary for this path is the Python
 * nn.Module / function surface.  This header is the boundary a binding for that surface
 * calls: plain pointers and sizes, no torch types.  Each entry point names the reference
 * interface it replaces (path:line under the reference tree).
 *
 * Conventions
 *   - every pointer is DEVICE memory owned by the caller (workspaces included);
 *   - every call is asynchronous on `stream` (a cudaStream_t passed as void*), never
 *     allocates, never synchronises, never throws;
 *   - return value 0 = OK; non-zero = RCNN_ERR_* (argument errors) or 1000 + cudaError_t;
 *     rcnn_last_error() returns a thread-local message for the last failure;
 *   - kernels are compiled for sm_100a only; there is no CPU fallback.
 */
#ifndef RCNN_OCR_B200_H
#define RCNN_OCR_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RCNN_OK 0
#define RCNN_ERR_ARG 1          /* invalid argument / unsupported shape              */
#define RCNN_ERR_WORKSPACE 2    /* workspace missing or too small                    */
#define RCNN_ERR_DEVICE 3       /* current device is not compute capability 10.x     */
#define RCNN_ERR_CUDA_BASE 1000 /* + cudaError_t                                     */

/* element types */
#define RCNN_F32 0
#define RCNN_BF16 1
#define RCNN_F16 2

/* CTC reductions (torch.nn.CTCLoss `reduction`) */
#define RCNN_REDUCE_NONE 0
#define RCNN_REDUCE_MEAN 1
#define RCNN_REDUCE_SUM 2

typedef void *rcnn_stream_t; /* cudaStream_t */

int rcnn_version(void);
const char *rcnn_last_error(void);
/* 0 when the current CUDA device can run these kernels (sm_100), else RCNN_ERR_DEVICE. */
int rcnn_device_check(void);

/* ---------------------------------------------------------------------------------------
 * K4  greedy CTC decode: argmax over classes, collapse repeats, strip blank.
 * Replaces training/utils.py:122-150 (ctc_greedy_decoder: `logits.argmax(dim=2)` + the
 * per-frame host loop) and the decode step of inference.py:167-180.
 *   logits    [B,T,C] addressed as logits[b*stride_b + t*stride_t + c] (element strides;
 *             class stride is 1), dtype RCNN_F32 or RCNN_BF16
 *   ids_out   [B,T] int32: the collapsed label ids, left-packed, padded with -1
 *   len_out   [B]   int32: number of ids per sequence
 *   conf_out  [B]   float or NULL: mean over non-blank frames of max softmax probability
 *             (the CTC analogue of inference.py:183-189; 0 when no such frame)
 * argmax follows torch: first maximal index wins, NaN is maximal.
 * ------------------------------------------------------------------------------------- */
int rcnn_ctc_greedy(const void *logits, int dtype, int B, int T, int C,
                    int64_t stride_b, int64_t stride_t, int blank,
                    int32_t *ids_out, int32_t *len_out, float *conf_out,
                    rcnn_stream_t stream);

/* ---------------------------------------------------------------------------------------
 * K3  CTC loss forward-backward, fused with log_softmax and its backward.
 * The reference has no CTC code; this is the nn.CTCLoss(blank,
 * reduction, zero_infinity) call the north_star places at training/train.py:289 and
 * :503-505 (and :557-559 for validation).
 *   x            [T,N,C] float32 addressed as x[t*stride_t + n*stride_n + c]
 *   from_logits  1: x are raw logits, log_softmax is applied inside and grad_out is the
 *                   gradient w.r.t. the logits;
 *                0: x are log-probabilities (nn.CTCLoss input) and grad_out follows ATen's
 *                   backward convention exp(x) - occupancy
 *   targets      int64; padded [N, tgt_stride] when tgt_stride > 0, else 1-D concatenated
 *                (offsets are derived on the device from target_lengths)
 *   input_lengths, target_lengths   int64 [N]
 *   max_target_len   host-side upper bound on target_lengths (sizes the label lattice)
 *   nll_out      [N] float32: per-sample negative log likelihood (inf if no alignment and
 *                !zero_infinity; 0 if zero_infinity)
 *   loss_out     [1] float32: the reduced loss (mean: mean_n nll_n / max(tgt_len_n,1);
 *                sum); untouched for RCNN_REDUCE_NONE
 *   grad_out     NULL (forward only) or float32 addressed like x through
 *                (gstride_t, gstride_n): gradient of the REDUCED loss (of sum_n nll_n for
 *                RCNN_REDUCE_NONE).  Frames t >= input_length get 0.
 *   workspace    rcnn_ctc_workspace_bytes(T, N, C, max_target_len) bytes
 * ------------------------------------------------------------------------------------- */
size_t rcnn_ctc_workspace_bytes(int T, int N, int C, int max_target_len);
int rcnn_ctc_loss(const float *x, int from_logits, int T, int N, int C,
                  int64_t stride_t, int64_t stride_n,
                  const int64_t *targets, int64_t tgt_stride,
                  const int64_t *input_lengths, const int64_t *target_lengths,
                  int max_target_len, int blank, int reduction, int zero_infinity,
                  float *nll_out, float *loss_out,
                  float *grad_out, int64_t gstride_t, int64_t gstride_n,
                  void *workspace, size_t workspace_bytes, rcnn_stream_t stream);

/* grad[t,n,:] *= scale[per_sample ? n : 0] (scale read on the device; rows whose factor is
 * exactly 1.0f are not touched).  Used to apply autograd's upstream gradient. */
int rcnn_ctc_scale_grad(float *grad, int T, int N, int C, int64_t gstride_t, int64_t gstride_n,
                        const float *scale, int per_sample, rcnn_stream_t stream);

/* ---------------------------------------------------------------------------------------
 * K1  dense bf16 GEMM on tcgen05/TMEM fed by TMA:  D[M,N] = A[M,K] * B[N,K]^T (+ bias[N]).
 * Replaces the library GEMMs behind the reference's encoder block: the input projection
 * W_ih x_t for all timesteps inside nn.LSTM (model/model.py:154-156,161) and
 * nn.Linear(2H -> out) (model/model.py:157,162); also used for the CTC head and, in the
 * backward pass, for dX and dW.
 *   A [M,K], B [N,K]  bf16 row-major, leading dimensions lda/ldb in elements (multiples of 8,
 *                     16-byte aligned bases)
 *   D [M,N]           row-major, ldd in elements, out_dtype RCNN_F32, RCNN_BF16 or RCNN_F16
 *   bias [N]          float32 or NULL
 * ------------------------------------------------------------------------------------- */
int rcnn_gemm_bf16(const void *A, int64_t lda, const void *B, int64_t ldb, void *D, int64_t ldd,
                   int out_dtype, const float *bias, int M, int N, int K, rcnn_stream_t stream);

/* Weight-gradient shape of K1:  D[M,N] (+)= A[K,M]^T * B[K,N], fp32 output.  A and B are row-major
 * bf16 with the CONTRACTION index as the row (MN-major UMMA operands, no transposed copies);
 * split-K with fp32 red.add.  accumulate == 0 zeroes D first.  Used for dW = dY^T X (autograd of
 * nn.Linear / nn.LSTM weights in the reference, model/model.py:154-157). */
int rcnn_gemm_bf16_atb(const void *A, int64_t lda, const void *B, int64_t ldb, float *D, int64_t ldd,
                       int M, int N, int K, int accumulate, rcnn_stream_t stream);
/* `groups` independent problems of the same shape in one launch: group g reads columns
 * [g*a_gcols, +M) of A and [g*b_gcols, +N) of B and writes D + g*d_goff (the two directions'
 * dW_hh = dG_d^T h_prev_d are one launch). */
int rcnn_gemm_bf16_atb_grouped(const void *A, int64_t lda, int a_gcols, const void *B, int64_t ldb, int b_gcols,
                               float *D, int64_t ldd, int64_t d_goff, int groups, int M, int N, int K,
                               int accumulate, rcnn_stream_t stream);

/* ---------------------------------------------------------------------------------------
 * K2  BidirectionalLSTM recurrence (model/model.py:151-163: nn.LSTM(bidirectional=True,
 * batch_first=True), gate order i,f,g,o, h0 = c0 = 0, final state discarded).
 *
 * rcnn_lstm_pack_weights: one launch converts the eight nn.LSTM parameter tensors of a block
 *   (weight_ih_l0, weight_hh_l0, bias_ih_l0, bias_hh_l0 and their _reverse twins, float32,
 *   torch layout) into the bf16 views the kernels read.  `packed` holds, in this order
 *   (P = packed gate-row order p(c,j,g) = c*128 + j*4 + g  <->  torch row g*H + 32c + j):
 *     wih_p  bf16 [2*4H, I]   rows in P order, direction-major   (B operand of the xp GEMM)
 *     bias_p f32  [2*4H]      b_ih + b_hh in P order
 *     whh_p  bf16 [2*4H, H]   rows in P order                    (resident operand, forward)
 *     whh_pt bf16 [2, H, 4H]  per-direction transpose of whh_p   (resident operand, backward)
 *     wih_pt bf16 [I, 2*4H]   transpose of wih_p                 (B operand of the dX GEMM)
 * rcnn_lstm_forward: the T dependent steps of both directions in one persistent kernel.
 *   xp        f16  [B*T, 2*4H]  x W_ih^T + b for every (b,t) (row b*T+t), columns in P order
 *                               (fp16: 11-bit mantissa keeps the pre-activation error ~5e-4 and
 *                               halves the bytes every step streams into the SM)
 *   whh_p     the whh_p view of `packed`
 *   hcat      bf16 [B, T, 2H]   output: h_t of the forward (cols [0,H)) and reverse ([H,2H))
 *                               direction; also the buffer h_{t-1} is re-read from
 *   gates_save f16 [2, T, B, 4H] activated gates in P order, c_save f32 [2, T, B, H]: both NULL
 *                               for inference, both set when a backward pass will follow
 *   H must be 64, 128, 256 or 512 (a group of H/32 CTAs owns one (direction, 64-sequence tile) work item at a
 *   time; cooperative launch, the groups synchronise per step through release / acquire counters).
 * ------------------------------------------------------------------------------------- */
size_t rcnn_lstm_packed_bytes(int I, int H);
int rcnn_lstm_pack_weights(const float *w_ih_f, const float *w_hh_f, const float *b_ih_f, const float *b_hh_f,
                           const float *w_ih_r, const float *w_hh_r, const float *b_ih_r, const float *b_hh_r,
                           int I, int H, void *packed, rcnn_stream_t stream);
/* parts: 1 = the views the forward pass reads (wih_p, bias_p, whh_p), 2 = the transposed views of the backward
 * pass (whh_pt, wih_pt), 3 = both: lets a caller convert the backward views on another stream while the forward
 * recurrence already runs. */
int rcnn_lstm_pack_weights_parts(const float *w_ih_f, const float *w_hh_f, const float *b_ih_f, const float *b_hh_f,
                                 const float *w_ih_r, const float *w_hh_r, const float *b_ih_r, const float *b_hh_r,
                                 int I, int H, void *packed, int parts, rcnn_stream_t stream);
int rcnn_lstm_forward(const void *xp, const void *whh_p, int B, int T, int H, void *hcat,
                      void *gates_save, float *c_save, rcnn_stream_t stream);

/* rcnn_lstm_forward_fused: variant of rcnn_lstm_forward with the input projection fused in: takes the
 * block input x (bf16 [B, T, I] contiguous, I a multiple of 64 and <= 512) and the wih_p / bias_p / whh_p views
 * of `packed` instead of xp; the W_ih slice stays resident in shared memory, `W_ih x_t` of step t+1 is multiplied
 * while step t waits for its exchange.  Same hcat output and the same saved activations (gates_save, c_save:
 * both NULL for inference) as rcnn_lstm_forward. */
int rcnn_lstm_forward_fused(const void *x, const void *wih_p, const float *bias_p, const void *whh_p, int B, int T,
                            int I, int H, void *hcat, void *gates_save, float *c_save, rcnn_stream_t stream);

/* rcnn_lstm_backward: BPTT through the recurrence of both directions (autograd of nn.LSTM in the
 * reference, model/model.py:161).
 *   whh_pt      the whh_pt view of `packed`
 *   gates_save, c_save   as written by rcnn_lstm_forward
 *   dhcat       f32 [B, T, 2H]  gradient w.r.t. hcat
 *   dG          bf16 [B, T, 2*4H] out: gradient w.r.t. the gate pre-activations (= w.r.t. xp),
 *               columns in P order.
 *   workspace   rcnn_lstm_backward_workspace_bytes(B,T,H) bytes of device memory (need not be zeroed): the
 *               bf16 partial sums of dG W_hh the CTAs of a group exchange every step
 *   db_p        f32 [2*4H] out or NULL: column sums of dG (= d b_ih = d b_hh in P order), accumulated in
 *               fp32 inside the kernel from the unrounded values.  dX = dG wih_p, dW_ih_p = dG^T X, dW_hh_p = dG^T H_prev,
 *               db_p = column sums of dG are then GEMMs / reductions (rcnn_gemm_bf16,
 *               rcnn_gemm_bf16_atb, rcnn_colsum_bf16, rcnn_lstm_hprev) and rcnn_lstm_unpack_grads scatters the
 *               P-ordered results back to torch's parameter layout. */
size_t rcnn_lstm_backward_workspace_bytes(int B, int T, int H);
int rcnn_lstm_backward(const void *whh_pt, const void *gates_save, const float *c_save, const float *dhcat,
                       int B, int T, int H, void *dG, float *db_p, void *workspace, size_t workspace_bytes,
                       rcnn_stream_t stream);
/* rcnn_lstm_plan: how the recurrent kernels lay a batch over the GPU -- `*ngroups` groups of H/32 CTAs, each working
 * on `*nslot` work items (direction, 64 sequences) at a time: 1 while every item gets a group of its own, 2 (two
 * interleaved dependency chains per half, one item's step inside the other's exchange latency) once items would queue
 * behind each other (the backward kernel; the forward kernel only with RCNN_FWD_SLOTS=2, it measures no gain there).
 * which: 0 = rcnn_lstm_forward_fused, 1 = rcnn_lstm_backward.  Reporting / tests only. */
int rcnn_lstm_plan(int which, int B, int H, int *nslot, int *ngroups);
/* out[col] = sum over rows of src[row*ld + col]  (bf16 [rows, cols] -> f32 [cols]) */
int rcnn_colsum_bf16(const void *src, int64_t ld, int64_t rows, int cols, float *out, rcnn_stream_t stream);
/* out bf16 [B, T, 2H]: out[b, t, dir*H+u] = hcat[b, t-1 (dir 0) / t+1 (dir 1), dir*H+u], 0 at the
 * direction's first step: the h that multiplied W_hh when gates_t were formed (B operand of the
 * dW_hh GEMM). */
int rcnn_lstm_hprev(const void *hcat, void *out, int B, int T, int H, rcnn_stream_t stream);
/* Both weight gradients of a block in one call and in nn.LSTM's row order (gate-major), no h_prev copy and no
 * unpack pass:  dwih[d] (4H x I) (+)= sum_{b,t} dG_d[b,t]^T x[b,t],  dwhh[d] (4H x H) (+)= sum dG_d[b,t]^T h_d[b,t-/+1]
 * (h_{t-1} for d = 0, h_{t+1} for d = 1, zero outside the sequence -- read from hcat through a shifted tensor map).
 * dG [B,T,8H] bf16 (packed gate order, as rcnn_lstm_backward writes it), x [B,T,I] bf16, hcat [B,T,2H] bf16;
 * dwih [2,4H,I], dwhh [2,4H,H] fp32.  accumulate == 0 zeroes the outputs first.  H in {256, 512}, I >= 256 and a
 * multiple of 32 (smaller blocks use rcnn_gemm_bf16_atb* + rcnn_lstm_hprev + rcnn_lstm_unpack_grads). */
int rcnn_lstm_weight_grads(const void *dG, const void *x, const void *hcat, int B, int T, int I, int H,
                           float *dwih, float *dwhh, int accumulate, rcnn_stream_t stream);
int rcnn_lstm_unpack_grads(const float *dwih_p, const float *dwhh_p, const float *db_p, int I, int H,
                           float *dw_ih_f, float *dw_hh_f, float *db_ih_f, float *db_hh_f,
                           float *dw_ih_r, float *dw_hh_r, float *db_ih_r, float *db_hh_r,
                           rcnn_stream_t stream);

/* Layout helpers used by the host side of the block (fp32 strided -> bf16 contiguous; 2-D
 * bf16 transpose with output row stride ldo >= R, so that odd R still gives 16-byte rows). */
int rcnn_cast_bf16_3d(const float *src, int64_t sb, int64_t st, int64_t sc, void *dst, int B, int T, int C,
                      rcnn_stream_t stream);
/* fp32 [rows, cols] (row stride ld_src) -> bf16 (row stride ld_dst >= cols; pad columns zeroed) */
int rcnn_cast_bf16_2d(const float *src, int64_t ld_src, void *dst, int64_t ld_dst, int64_t rows, int cols,
                      rcnn_stream_t stream);
int rcnn_transpose_bf16(const void *src, int64_t ld, void *dst, int64_t ldo, int R, int C, rcnn_stream_t stream);

/* ---------------------------------------------------------------------------------------
 * K5  batched edit distance for the validation metrics (training/metrics.py:5-32 as used at
 * training/train.py:582-598 and evaluate_dataset.py:104-119: CER = Levenshtein / len(reference)
 * on characters, WER on words, accuracy = exact match).
 *   hyp_ids   int32 [N, hyp_stride]  class ids left by rcnn_ctc_greedy (-1 padded), hyp_len int32 [N]
 *   ref_ids   int64 flat class ids (CTC targets), ref_off int64 [N] start of pair n, ref_len int64 [N]
 *   cp_off    int32 [C+1], cp int32: class k expands to the Unicode code points cp[cp_off[k]..cp_off[k+1])
 *             (class 0 = blank = empty; class k = itos[k-1], training/utils.py:146)
 *   words     0: distance over code points; 1: over words (split on U+0020, empty words dropped)
 *   dist_out, nref_out, nhyp_out  int32 [N]: distance and the two symbol counts; dist = -1 when a
 *             sequence expands to more than 320 symbols
 * ------------------------------------------------------------------------------------- */
int rcnn_edit_distance(const int32_t *hyp_ids, int64_t hyp_stride, const int32_t *hyp_len,
                       const int64_t *ref_ids, const int64_t *ref_off, const int64_t *ref_len, int N,
                       const int32_t *cp_off, const int32_t *cp, int C, int words,
                       int32_t *dist_out, int32_t *nref_out, int32_t *nhyp_out, rcnn_stream_t stream);

/* ---------------------------------------------------------------------------------------
 * K6  per-step kernels of the attention decoder (model/model.py:23-148, inference path); the
 * matrix products of a step (h2h, the LSTMCell gates, generator) go through rcnn_gemm_bf16 and
 * i2h(batch_H) is computed once.
 * rcnn_attn_score_context (model/model.py:35-45): e = score(tanh(proj_H + proj_h)), alpha =
 *   softmax over the T encoder frames, context = alpha^T enc.
 *     projH f32 [B,T,H], projh f32 [B,H], v f32 [H] (score.weight), enc f32 [B,T,C] addressed as
 *     enc[b*stride_b + t*stride_t + c]; alpha_out f32 [B,T] or NULL; the context is written as
 *     bf16 into xcat[b*ldx + 0..C) -- the first half of the LSTMCell GEMM operand [context | h].
 * rcnn_attn_cell (nn.LSTMCell at model/model.py:46, gate order i,f,g,o): gates f32 [B,4H] = the GEMM
 *   over [context | h_{t-1}], embT f32 [V,4H] = columns C.. of rnn.weight_ih transposed (the one-hot
 *   input of :44 is a row gather by y int64 [B]); c f32 [B,H] updated in place; h_t is written as
 *   bf16 into xcat[b*ldx + C + j] and (if hid_out != NULL) as f32 into hid_out[b*hid_ld + j].
 * rcnn_attn_argmax (model/model.py:104-108): logits[:, blank] = -1e4 (blank < 0: no mask), the
 *   masked row is copied to probs[b*probs_ld + k] (if probs != NULL), y[b] = argmax (first maximum).
 * ------------------------------------------------------------------------------------- */
int rcnn_attn_score_context(const float *projH, const float *projh, const float *v, const float *enc,
                            int64_t enc_stride_b, int64_t enc_stride_t, int B, int T, int H, int C,
                            float *alpha_out, void *xcat, int64_t ldx, rcnn_stream_t stream);
int rcnn_attn_cell(const float *gates, const float *embT, const int64_t *y, int B, int H, int V, float *c,
                   void *xcat, int64_t ldx, int C, float *hid_out, int64_t hid_ld, rcnn_stream_t stream);
int rcnn_attn_argmax(const float *logits, int B, int V, int blank, float *probs, int64_t probs_ld,
                     int64_t *y, rcnn_stream_t stream);
/* the same two with explicit row pitches of proj_h / logits (elements): the greedy decode forms h2h(h_t) for the next step and
 * generator(h_t) for this one in ONE product over concatenated weights and hands each kernel its column block */
int rcnn_attn_score_context_ld(const float *projH, const float *projh, int64_t projh_ld, const float *v, const float *enc,
                               int64_t enc_stride_b, int64_t enc_stride_t, int B, int T, int H, int C,
                               float *alpha_out, void *xcat, int64_t ldx, rcnn_stream_t stream);
int rcnn_attn_argmax_ld(const float *logits, int64_t logits_ld, int B, int V, int blank, float *probs, int64_t probs_ld,
                        int64_t *y, rcnn_stream_t stream);
/* rcnn_attn_score_context_ld with proj_H [B,T,H] and enc [B,T,C] held as bf16 (the decode loop's only large reads, halved;
 * both pass through bf16 tensor-core operands anyway).  H, C, enc_stride_* multiples of 8 elements, arrays 16-byte aligned
 * (RCNN_ERR_ARG otherwise: there is no slow path). */
int rcnn_attn_score_context_bf16(const void *projH, const float *projh, int64_t projh_ld, const float *v, const void *enc,
                                 int64_t enc_stride_b, int64_t enc_stride_t, int B, int T, int H, int C,
                                 float *alpha_out, void *xcat, int64_t ldx, rcnn_stream_t stream);
/* the same, preceded (if prev_logits != NULL) by rcnn_attn_argmax of the PREVIOUS step's logits for the same sequences:
 * masked row -> prev_probs (may be NULL), argmax -> y.  The greedy loop then needs four launches per step. */
int rcnn_attn_step_bf16(const void *projH, const float *projh, int64_t projh_ld, const float *v, const void *enc,
                        int64_t enc_stride_b, int64_t enc_stride_t, int B, int T, int H, int C, float *alpha_out,
                        void *xcat, int64_t ldx, const float *prev_logits, int64_t prev_ld, int V, int blank,
                        float *prev_probs, int64_t probs_ld, int64_t *y, rcnn_stream_t stream);

/* The decoder's gate product and LSTMCell step in one launch (model/model.py:43-45: rnn(concat([context, onehot]), hidden)):
 * [c, h_t] = LSTMCell pointwise of  xcat[B, K] @ wcat_il[4H, K]^T + bias_il + embT_il[y]  with c [B, H] updated in place and h_t
 * written as bf16 to h_out (row pitch h_ld; must not alias xcat) and, if hid_out != NULL, as f32.  The "_il" arrays are
 * gate-interleaved along 4H: row / element 4u + g is gate g (torch order i, f, g, o) of hidden unit u.  H % 8 == 0. */
int rcnn_attn_gates_cell(const void *xcat, int64_t ldx, const void *wcat_il, int64_t ldw, const float *bias_il,
                         const float *embT_il, const int64_t *y, int B, int H, int K, int V, float *c, void *h_out,
                         int64_t h_ld, float *hid_out, int64_t hid_ld, rcnn_stream_t stream);

/* Attention._greedy_decode (model/model.py:89-108) as one host call: `steps` times (rcnn_attn_step_bf16, rcnn_attn_gates_cell,
 * rcnn_gemm_bf16 over comb_w = [W_h2h ; W_generator] padded to comb_rows >= H + V rows) and a final rcnn_attn_argmax_ld.
 * The caller provides: projH [B,T,H] bf16 (i2h(batch_H), hoisted), enc [B,T,C] bf16, the gate-interleaved weights of
 * rcnn_attn_gates_cell, y [B] = <SOS>, xcat0 / xcat1 [B, C+H] bf16 and c [B,H] f32 zeroed, hg [B, comb_rows] f32 with columns [0, H)
 * = the h2h bias (h_0 = 0).  Writes probs [B, steps, V] (blank-masked logits) and leaves the last tokens in y.  chain != 0 launches
 * the loop as a programmatic-dependent chain (rcnn_chain_launches). */
int rcnn_attn_greedy_decode(const void *projH, const float *v, const void *enc, int64_t enc_stride_b, int64_t enc_stride_t,
                            const void *wcat_il, const float *bcat_il, const float *embT_il, const void *comb_w,
                            const float *comb_b, int comb_rows, int B, int T, int H, int C, int V, int steps, int blank,
                            int64_t *y, void *xcat0, void *xcat1, float *c, float *hg, float *probs, int chain,
                            rcnn_stream_t stream);

/* ---------------------------------------------------------------------------------------
 * Attention decoder, teacher-forced TRAINING pass (model/model.py:110-148 under autograd; the reference's live loss path,
 * training/train.py:499-505).  Forward step t: rcnn_gemm_bf16 (h2h) -> rcnn_attn_step_train -> rcnn_attn_gates_cell_train;
 * backward step t: rcnn_attn_cell_bwd -> rcnn_gemm_bf16 (dcontext) -> rcnn_attn_step_bwd -> rcnn_gemm_bf16 (dh_{t-1});
 * after the loop rcnn_attn_dprojH once and the weight gradients as products over all (step, sequence) rows.
 *
 * rcnn_attn_step_train: rcnn_attn_score_context_bf16 that also returns alpha (before dropout; required) and multiplies it by
 *   alpha_scale [B,T] (F.dropout(alpha), model/model.py:40: 0 or 1/(1-p), drawn by the caller; NULL = no dropout) for the context.
 * rcnn_attn_gates_cell_train: rcnn_attn_gates_cell with c_{t-1} read from c_in, c_t written to c, and the gate activations
 *   (sigmoid i, sigmoid f, tanh g, sigmoid o; gate-interleaved [B,4H] f32) kept in gates_act (may be NULL).
 * rcnn_attn_cell_bwd: dh = dh_a (+ dh_b if not NULL), dc [B,H] holds dc_t on entry and dc_{t-1} on return; writes the gradients
 *   of the gate pre-activations as bf16, gate-interleaved, to dg (row pitch ldg >= 4H).  c_prev NULL = zeros (t = 0).
 * rcnn_attn_step_bwd: from dcontext [B, dctx_ld], alpha, alpha_scale, enc / proj_H (bf16), proj_h, v: de [B,T] (gradient of the
 *   scores), d proj_h as bf16 (row pitch dprojh_ld) and dv_acc[b, :] += sum_t de[t] tanh(.) (one row per sequence; the caller
 *   sums the rows at the end).
 * rcnn_attn_dprojH: dprojH[b,t,j] (bf16) = v[j] sum_s de_all[s,b,t] (1 - tanh^2(projH[b,t,j] + projh_all[s,b,j])).
 * ------------------------------------------------------------------------------------- */
int rcnn_attn_step_train(const void *projH, const float *projh, int64_t projh_ld, const float *v, const void *enc,
                         int64_t enc_stride_b, int64_t enc_stride_t, int B, int T, int H, int C, float *alpha_out,
                         const float *alpha_scale, void *xcat, int64_t ldx, rcnn_stream_t stream);
int rcnn_attn_gates_cell_train(const void *xcat, int64_t ldx, const void *wcat_il, int64_t ldw, const float *bias_il,
                               const float *embT_il, const int64_t *y, int B, int H, int K, int V, const float *c_in, float *c,
                               void *h_out, int64_t h_ld, float *hid_out, int64_t hid_ld, float *gates_act, rcnn_stream_t stream);
int rcnn_attn_cell_bwd(const float *gates_act, const float *c_prev, const float *c_t, const float *dh_a, int64_t dh_a_ld,
                       const float *dh_b, int64_t dh_b_ld, float *dc, int B, int H, void *dg, int64_t ldg, rcnn_stream_t stream);
int rcnn_attn_step_bwd(const float *dctx, int64_t dctx_ld, const float *alpha, const float *alpha_scale, const void *enc,
                       int64_t enc_stride_b, int64_t enc_stride_t, const void *projH, const float *projh, int64_t projh_ld,
                       const float *v, int B, int T, int H, int C, float *de_out, void *dprojh, int64_t dprojh_ld, float *dv_acc,
                       rcnn_stream_t stream);
/* The two step loops of the training pass as single host calls (the same launches in the same order).  Shapes: tokens [S,B] int64,
 * alpha_scale [S,B,T] or NULL, xcat_all [S+1, B, C+H] bf16 (zeroed; row block t = [context_t | h_{t-1}]), c_all [S+1, B, H] f32
 * (zeroed; block t = c_{t-1}), gates_all [S,B,4H], alpha_all [S,B,T], projh_all [S,B,H], out_hid [B,S,H]; backward: d_out [B,S,H],
 * b1 [C,4H] = (W_ih[:, :C], gate-interleaved)^T, b2 [H,5H] = [W_hh (gate-interleaved)^T | W_h2h^T], dg_all [S,B,5H] bf16 (zeroed),
 * dctx_all [S,B,C], de_all [S,B,T], dv_acc [B,H] and dc [B,H] zeroed, dh [B,H] scratch. */
int rcnn_attn_train_forward(const void *projH, const float *v, const void *enc, int64_t enc_stride_b, int64_t enc_stride_t,
                            const void *h2h_w, const float *h2h_b, const void *wcat_il, const float *bcat_il, const float *embT_il,
                            const int64_t *tokens, const float *alpha_scale, int B, int T, int H, int C, int V, int S,
                            void *xcat_all, float *c_all, float *gates_all, float *alpha_all, float *projh_all, float *out_hid,
                            rcnn_stream_t stream);
int rcnn_attn_train_backward(const float *d_out, const float *gates_all, const float *c_all, const void *b1, const void *b2,
                             const float *alpha_all, const float *alpha_scale, const void *enc, int64_t enc_stride_b,
                             int64_t enc_stride_t, const void *projH, const float *projh_all, const float *v, int B, int T, int H,
                             int C, int S, void *dg_all, float *dctx_all, float *de_all, float *dv_acc, float *dc, float *dh,
                             rcnn_stream_t stream);
int rcnn_attn_dprojH(const float *de_all, const float *projh_all, const void *projH, const float *v, int S, int B, int T, int H,
                     void *dprojH, rcnn_stream_t stream);

/* ---------------------------------------------------------------------------------------
 * Backbone, inference: the squeeze-and-excitation tail of an SE-ResNet block (model/seresnet31.py:
 * SELayer, then the block's residual add and ReLU) in two launches, and the stem's max pooling.  Tensors channels_last: y / skip /
 * out [B, HW, C] with C contiguous, dtype RCNN_F32 or RCNN_BF16; w1 [Cr, C] f32, w2t [Cr, C] f32 = the second product's weight [C, Cr] TRANSPOSED; gate [B, C] f32; ybias / sbias [C]
 * f32 or NULL: the biases of the convolutions that produced y / skip (folded BatchNorm shifts), added here instead of in a pass
 * of their own.
 *   rcnn_se_gate:  gate = sigmoid(w2t^T relu(w1 (mean_over_HW(y) + ybias))).  workspace: rcnn_se_gate_workspace_bytes(B, C) bytes,
 *                  ZEROED by the caller before the first use and left zeroed (partial sums + one counter per image).
 *   rcnn_se_apply: out = relu((y + ybias) * gate + skip + sbias)      (C a multiple of 8 (bf16) / 4 (f32); out may alias y or skip)
 *   rcnn_maxpool2x2_nhwc: 2 x 2 max pooling, stride 2 (nn.MaxPool2d(2, 2)); H, W even; NaN propagates as in torch.
 * ------------------------------------------------------------------------------------- */
size_t rcnn_se_gate_workspace_bytes(int B, int C);
int rcnn_se_gate(const void *y, int dtype, int B, int HW, int C, const float *w1, const float *w2t, int Cr, const float *ybias,
                 float *gate, void *workspace, rcnn_stream_t stream);
int rcnn_se_apply(const void *y, const void *skip, const float *gate, const float *ybias, const float *sbias, int dtype, int B,
                  int HW, int C, void *out, rcnn_stream_t stream);
int rcnn_maxpool2x2_nhwc(const void *x, int dtype, int B, int H, int W, int C, void *out, rcnn_stream_t stream);

/* ---------------------------------------------------------------------------------------
 * Per-kernel device timing for the roofline report (bench.py): when enabled, every launch
 * of a dominant kernel is bracketed by cudaEventRecord on the launching stream.
 * rcnn_prof_read synchronises the recorded events and returns the summed duration.
 * ------------------------------------------------------------------------------------- */
#define RCNN_K_DECODE 0
#define RCNN_K_CTC 1
#define RCNN_K_GEMM 2
#define RCNN_K_LSTM_FWD 3
#define RCNN_K_LSTM_BWD 4
#define RCNN_K_GEMM_ATB 5   /* gemm_atb_kernel (weight-gradient shape); RCNN_K_GEMM is gemm_tn_kernel */
#define RCNN_K_COUNT 8
/* Debug aid: when buf != NULL the recurrent kernels record clock64() marks of cluster 0 / CTA 0 per
 * timestep into buf[step*8 + k] (int64); NULL switches it off. */
int rcnn_debug_timeline(void *buf);
/* Debug aid: device pointer to one uint32 that the recurrent kernels increment whenever a validating warp had to
 * re-fetch exchange packets that the optimistic TMA fetch overtook (flag-in-data exchange); NULL switches it off. */
int rcnn_debug_refetch_counter(void *counter);
/* The persistent GEMM kernels (K1) size their grids to the SM count minus `n` (default 0; 0 <= n <= 64).  The data-parallel
 * reducer reserves 16: NCCL's all-reduce kernels then run beside a weight-gradient or input-gradient GEMM instead of waiting
 * for the one CTA per SM it would hold until its last tile (rcnn-ocr_b200/dist.py). */
int rcnn_reserve_sms(int n);
/* on != 0: the GEMM (rcnn_gemm_bf16, rcnn_attn_gates_cell) and attention step (rcnn_attn_step_bf16 / _score_context_bf16)
 * kernels launched from now on carry cudaLaunchAttributeProgrammaticStreamSerialization: each may start its CTAs while its
 * predecessor in the stream drains and waits (griddepcontrol.wait) before touching anything that predecessor produced.  Meant
 * for a chain of short dependent kernels -- the decoder's step loop sets it around the loop.  Process-wide; default off. */
int rcnn_chain_launches(int on);
/* number of kernels this library has launched in this process (bench.py's gpu_launches) */
unsigned long long rcnn_launch_count(void);
int rcnn_prof_enable(int on);
int rcnn_prof_reset(void);
int rcnn_prof_read(int kernel, double *total_ms, int *launches);

/* ---------------------------------------------------------------------------------------
 * K7  host -> device input step: ResizeAndPadA + Normalize(0.5, 0.5) + HWC -> CHW + batch
 *     (data/transforms.py:62-120,179; inference.py:93-124,159-164; data/dataset.py:147-156)
 *   pixels   u8, the decoded images of the batch packed into one buffer (any sizes, row pitch per image)
 *   desc     i64 [N, 6] per image: byte offset, height, width, row pitch in bytes, channels (1 grey, 3, 4 = alpha
 *            dropped), bgr (1: channel order B,G,R as cv2.imread returns it)
 *   out      f32 (out_dtype 0) or bf16 (1) [N, 3, img_h, img_w]: each image resized with its aspect ratio kept
 *            (scale = min(img_h / h, img_w / w), OpenCV INTER_AREA arithmetic when shrinking, INTER_LINEAR when
 *            enlarging), placed on a white canvas (align_h / align_v: 0 left / top, 1 center, 2 right / bottom; the
 *            reference uses left / center), then (v - 127.5) / 127.5.
 * ------------------------------------------------------------------------------------- */
int rcnn_preprocess_lines(const void *pixels, const int64_t *desc, int N, int img_h, int img_w, int align_h,
                          int align_v, void *out, int out_dtype, rcnn_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* RCNN_OCR_B200_H */
