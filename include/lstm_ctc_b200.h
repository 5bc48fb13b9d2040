/*
 * lstm_ctc_b200.h -- C ABI of liblstm_ctc_b200.so (sm_100a kernels for the BiLSTM / mixture
 * output / CTC training hot path of mobvoi/lstm_ctc).
 *
 * The reference has no FFI of its own: every entry below replaces a TensorFlow-1.8 op call
 * site of the reference's graph construction (cited per function, paths relative to
 * /root/reference).  A maintainer binds them from Python with ctypes.
 *
 * Conventions: plain pointers + sizes, all pointers are DEVICE pointers unless a name ends in
 * _host; `stream` is a cudaStream_t passed as void*; nothing allocates -- callers provide
 * workspaces sized by the matching *_workspace_bytes query; every function returns an int
 * status (0 = ok, <0 = error class, see lcb_status_string) and never throws.  Calls are
 * asynchronous on `stream` and re-entrant across streams.
 */
#ifndef LSTM_CTC_B200_H_
#define LSTM_CTC_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum {
    LCB_OK = 0,
    LCB_ERR_NULL_POINTER = -1,
    LCB_ERR_BAD_SHAPE = -2,
    LCB_ERR_UNSUPPORTED = -3,
    LCB_ERR_WORKSPACE_TOO_SMALL = -4,
    LCB_ERR_CUDA = -5,
    LCB_ERR_INVALID_LABEL = -6,   /* TF: InvalidArgument, label not in [0, num_classes-1) */
    LCB_ERR_MISALIGNED = -7,
    LCB_ERR_DEVICE_TIMEOUT = -8
};

/* ---- library ------------------------------------------------------------------------- */
int lcb_version(void);
const char* lcb_status_string(int status);
/* device-side error word (barrier time-outs etc.); returns 0 if clean.  reset!=0 clears it.
 * Synchronises the device. */
int lcb_device_error(int reset);
/* number of kernels this library has launched in this process; reset != 0 zeroes the counter. */
long long lcb_launch_count(int reset);
/* kernels launched on the library's behalf by a replayed CUDA graph (the host wrapper captured n of them once and
 * adds n per replay, so the counter keeps meaning "kernels of this library that ran"). */
void lcb_launch_count_add(long long n);

/* ---- CTC loss + gradient ---------------------------------------------------------------
 * replaces: tf.nn.ctc_loss(labels, inputs, sequence_length,
 *                          ignore_longer_outputs_than_inputs=True)   nnet/graph.py:109-114
 *           and the transpose feeding it                               nnet/graph.py:72
 *           and the dense(-1 padded) -> sparse label conversion        nnet/graph.py:74-104
 * logits [B,T,V] f32 batch-major; labels [B,Lmax] int64, -1 padded; seq_len [B] int32.
 * blank = V-1.  loss [B]; grad [B,T,V] = d loss_b / d logits (0 past seq_len).
 * Skipped utterance (seq_len==0 or #labels > seq_len): loss 0, grad 0.
 * No valid alignment: loss +inf, grad = softmax. */
size_t lcb_ctc_workspace_bytes(int B, int T, int V, int Lmax);
int lcb_ctc_loss_grad_f32(const float* logits, const int64_t* labels, int Lmax, const int32_t* seq_len,
                          int B, int T, int V, float* loss, float* grad,
                          void* workspace, size_t workspace_bytes, void* stream);
/* Same call with the lattice layout stated: -1 chosen by batch size (what lcb_ctc_loss_grad_f32 does: the alpha and beta
 * sweeps of an utterance as the two CTAs of a cluster when 2 B <= SMs, otherwise as two warp groups of one CTA), 0 one CTA
 * per utterance, 1 two.  Results agree up to the order of the gradient's atomic adds.  (More than 1023 labels per utterance --
 * 8 or 16 lattice states per thread -- always run as two CTAs.) */
int lcb_ctc_loss_grad_f32_layout(const float* logits, const int64_t* labels, int Lmax, const int32_t* seq_len,
                                 int B, int T, int V, float* loss, float* grad,
                                 void* workspace, size_t workspace_bytes, int lattice_layout, void* stream);
/* 0, or LCB_ERR_INVALID_LABEL if the last call on this workspace saw an out-of-range label.
 * Synchronises `stream`. */
int lcb_ctc_status(const void* workspace, void* stream);

/* ---- 16-bit GEMM, fp32 accumulate (tcgen05 + TMA) ---------------------------------------
 * replaces: the matmul inside tf.contrib.rnn.LSTMCell hoisted over all frames
 *           (nnet/bilstm.py:129-136,171-188), the LSTM projection, tf.nn.xw_plus_b
 *           (nnet/bilstm.py:249, nnet/moe.py:42,59) and their tf.gradients counterparts
 *           (nnet/graph.py:190-191).
 * C[M,N] (+)= op(A)[M,K] * op(B)[K,N] + bias[N]
 *   a_layout 0: A stored row-major [M,K] (ld = lda)      1: A stored row-major [K,M]
 *   b_layout 0: B stored row-major [N,K] (ld = ldb)      1: B stored row-major [K,N]
 *   dtype codes: 0 = fp32 (C only), 1 = bf16, 2 = fp16.  A and B must share one 16-bit type (the
 *   hardware rejects mixed kind::f16 operands): forward GEMMs are fp16 x fp16, gradient GEMMs bf16 x bf16.
 * lda/ldb multiples of 8 elements, A/B base pointers 16-byte aligned.
 * accumulate != 0 adds into the existing fp32 C (c_dtype must be 0).
 * max_ctas caps the persistent grid of THIS launch (0 = one CTA per SM of the device): a GEMM that runs on a side
 * stream next to a cluster kernel owning part of the SMs is launched with as many CTAs as there are free SMs. */
int lcb_gemm16(int M, int N, int K,
               const void* A, int lda, int a_layout, int a_dtype,
               const void* B, int ldb, int b_layout, int b_dtype,
               void* C, int ldc, int c_dtype,
               const float* bias, int accumulate, int max_ctas, void* stream);
/* Same, with inverted dropout fused into the epilogue: C = dropout(op(A)*op(B) + bias), where element (row, col) of C uses
 * element mask_base + row*ldc + col of the counter-based mask stream of lcb_dropout16 / lcb_dropout_mask (seed).  Replaces
 * DropoutWrapper(output_keep_prob) on the layer output h = m*W_proj (nnet/bilstm.py:128,137) and, with the same seed, the
 * mask on its gradient.  keep_prob == 1: identical to lcb_gemm16.  Needs a 16-byte aligned C and (mask_base | ldc) % 4 == 0. */
int lcb_gemm16_dropout(int M, int N, int K, const void* A, int lda, int a_layout, int a_dtype,
                       const void* B, int ldb, int b_layout, int b_dtype,
                       void* C, int ldc, int c_dtype, const float* bias, int accumulate,
                       float keep_prob, unsigned long long seed, unsigned long long mask_base, int max_ctas, void* stream);
/* Same, writing the fp16 result a second time as bf16 into C_bf16 (same shape and pitch; NULL = lcb_gemm16_dropout; c_dtype must be 2):
 * the layer output h = dropout(m*W_proj) feeds the next layer's forward GEMM as fp16 and its weight-gradient GEMM as bf16
 * (tcgen05 kind::f16 cannot mix the two), and the second store from the staged tile replaces a conversion pass over [N, 2P]. */
int lcb_gemm16_twin(int M, int N, int K, const void* A, int lda, int a_layout, int a_dtype,
                    const void* B, int ldb, int b_layout, int b_dtype,
                    void* C, int ldc, int c_dtype, void* C_bf16, const float* bias, int accumulate,
                    float keep_prob, unsigned long long seed, unsigned long long mask_base, int max_ctas, void* stream);
/* number of SMs of the current device (what max_ctas = 0 means; grids of every kernel are sized from it). */
int lcb_device_sm_count(void);
/* bf16 x bf16 shorthand of the above (max_ctas = 0). */
int lcb_gemm_bf16(int M, int N, int K,
                  const void* A, int lda, int a_layout,
                  const void* B, int ldb, int b_layout,
                  void* C, int ldc, int c_dtype,
                  const float* bias, int accumulate, void* stream);
/* Same contracts on plain CUDA cores -- slow on-device CHECKERS for tests, not a product path. */
int lcb_gemm16_simt_check(int M, int N, int K,
                          const void* A, int lda, int a_layout, int a_dtype,
                          const void* B, int ldb, int b_layout, int b_dtype,
                          void* C, int ldc, int c_dtype,
                          const float* bias, int accumulate, void* stream);
int lcb_gemm_bf16_simt_check(int M, int N, int K,
                             const void* A, int lda, int a_layout,
                             const void* B, int ldb, int b_layout,
                             void* C, int ldc, int c_dtype,
                             const float* bias, int accumulate, void* stream);

/* ---- LSTM recurrence, both directions of one BiLSTM layer (cluster-persistent, DSMEM) ----
 * replaces: tf.nn.dynamic_rnn(DropoutWrapper(LSTMCell(num_units, num_proj, use_peepholes,
 *           forget_bias=5.0)), sequence_length=...) for "fd{i}" AND "bd{i}"  nnet/bilstm.py:127-188
 *           and tf.reverse_sequence                                          nnet/bilstm.py:112,190,203
 * Hp = hidden size padded to a multiple of 64 (<= 512).  Packed gate column of (unit, gate) =
 * (unit/8)*32 + gate*8 + unit%8, gates (i,j,f,o); direction d owns columns [d*4Hp, (d+1)*4Hp).  Time-major rows n = t*B + b.
 *   G      [T*B, 8Hp] f32   x_t*W_x + bias (from lcb_gemm16)
 *   WfoldT [8Hp, Hp]  fp16  (W_proj*W_h)^T: rows = packed gate columns of both directions, cols = unit.
 *                           Each CTA keeps its 128 rows resident in tensor memory for the whole sequence.
 *   peep   [2,3,Hp]   f32   (w_f, w_i, w_o) per direction, or NULL (use_peepholes = False)
 *   lens   [B] int32        sequence_length
 *   Mout   [T*B, 2Hp] fp16  m_t = o*tanh(c) per direction; rows with t >= lens[b] are 0 (dynamic_rnn
 *                           zero output); h = Mout * W_proj is a bulk lcb_gemm16 afterwards
 *   gates  [T*B, 2Hp] 8 B   saved (i, tanh j, f, o) as 4 x fp16, and
 *   cst    [T*B, 2Hp] f32   saved cell state, for BPTT -- both NULL for inference
 *   cfin, mfin [B,2,Hp] f32 final states (both or neither) -- `encoder` of nnet/bilstm.py:206-208 */
int lcb_lstm_rec_config(int Hp, int* units_per_cta_div32, int* cluster_size);
/* clusters of the forward (which=0) / BPTT (which=1) kernel the device keeps resident at once (<0: error) */
int lcb_lstm_rec_max_clusters(int Hp, int which);
/* SMs (one CTA each) a forward (which=0) / BPTT (which=1) launch over B utterances occupies -- what a GEMM overlapped with it
 * on another stream must leave free: its max_ctas = lcb_device_sm_count() - lcb_lstm_rec_grid(...). */
int lcb_lstm_rec_grid(int B, int Hp, int num_dirs, int which);
/* debug probe: the next lcb_lstm_rec_fwd launches write steps*16 clock64 samples of CTA 0 into buf (NULL: off). */
int lcb_debug_rec_profile(long long* buf, int steps);
/* debug: forward cluster layout (0 = automatic: as many clusters as stay resident, the surplus 16-utterance groups paired into
 * the first clusters; 1 = two groups in every cluster, the round-1 layout).  Results are identical, only the timing differs. */
int lcb_debug_fwd_layout(int mode);
/* debug / measurements: smallest number of CTC lattice states per thread the plan may choose (2 | 4 | 8 | 16; default 2).
 * Changes lcb_ctc_workspace_bytes as well. */
int lcb_debug_ctc_min_spt(int spt);
/* workspace (required): device scratch of lcb_lstm_rec_workspace_bytes(B, Hp) bytes (16-byte aligned, caller-owned, one per
 * concurrently running launch): the per-step exchange of m_t goes  shared memory -> bulk store -> this L2-resident
 * scratch -> ONE multicast bulk load into all CTAs of the cluster.
 * num_dirs: 2 = BiLSTM layer, clusters alternate between the "fd" and "bd" direction and both run concurrently;
 *           1 = uni-directional layer (nnet/lstm.py:236-260, dynamic_rnn scope "drnn{i}"): only direction-0 clusters are
 *               launched; the direction-1 column halves of G / Mout / gates / cst / dG / dbias / dpeep are neither read nor
 *               written (tensor shapes stay those of the two-direction layout). */
size_t lcb_lstm_rec_workspace_bytes(int B, int Hp);
int lcb_lstm_rec_fwd(const float* G, const void* WfoldT, const float* peep, const int32_t* lens,
                     void* Mout, void* gates, float* cst, float* cfin, float* mfin,
                     int T, int B, int Hp, int num_dirs, float forget_bias,
                     void* workspace, size_t workspace_bytes, void* stream);
/* Same, for scan steps [s_begin, s_end) only (scan step s is frame s of the forward direction and frame T-1-s of the
 * backward direction).  A launch with s_begin > 0 resumes from the saved cst / Mout rows of scan step s_begin-1 (so gates
 * and cst must be given); launches over consecutive ranges in stream order equal one launch over [0, T).  This lets the
 * caller start the recurrence when only the first frames' pre-activations G exist and compute the rest beside it. */
int lcb_lstm_rec_fwd_range(const float* G, const void* WfoldT, const float* peep, const int32_t* lens,
                           void* Mout, void* gates, float* cst, float* cfin, float* mfin,
                           int T, int B, int Hp, int num_dirs, float forget_bias, int s_begin, int s_end,
                           void* workspace, size_t workspace_bytes, void* stream);
/* *dst = value on `stream` (one thread): advances the ready_steps counter below behind a projected chunk of G. */
int lcb_store_i32(int32_t* dst, int32_t value, void* stream);
/* Same, with the utterance lengths ALSO given on the host (lens_host [B], as every batch assembler has them; NULL = as above).
 * The kernel then skips, per 16-utterance group, the scan steps in which no utterance of the group is live (behind the group's
 * longest utterance in the forward direction, before it starts in the backward direction): those steps cost no weight pass and
 * no exchange, only the zero rows of m are written.  The group maxima travel in the kernel parameters -- control flow around
 * the tcgen05.mma issue must be provably warp-uniform, which a value loaded from device memory is not.  Results are identical
 * to lcb_lstm_rec_fwd_range except that gates / cst rows of the skipped steps (never read by BPTT) are left unwritten.
 * ready_steps (device word, NULL = everything is there): the number of leading scan steps whose G rows exist -- frames [0, n) of
 * the forward and [T-n, T) of the backward direction.  The caller projects G chunk by chunk on ANOTHER stream, on the SMs this
 * launch leaves idle (lcb_lstm_rec_grid), and advances the word with lcb_store_i32 behind every chunk; the kernel's prefetch warp
 * waits for it (bounded: LCB_WAIT_TIMEOUT_NS, then lcb_device_error).  One launch covers the scan; the chunks must already be
 * enqueued when it is launched, and the two streams must be able to run concurrently (not under a serialising profiler). */
int lcb_lstm_rec_fwd_range_hl(const float* G, const void* WfoldT, const float* peep, const int32_t* lens, const int32_t* lens_host,
                              const int32_t* ready_steps,
                              void* Mout, void* gates, float* cst, float* cfin, float* mfin,
                              int T, int B, int Hp, int num_dirs, float forget_bias, int s_begin, int s_end,
                              void* workspace, size_t workspace_bytes, void* stream);
/* The same launch publishing its progress (the forward mirror of lcb_lstm_rec_bwd_range_pg): word (cluster, sub-group, CTA) of
 * `progress` -- lcb_lstm_rec_fwd_progress_words(B, Hp, num_dirs) int32 words, zeroed by the caller, NULL = none -- counts the
 * leading scan steps whose Mout rows that CTA has written (rows [0, n) of the forward, [T-n, T) of the backward direction's column
 * half); advanced every 16 steps, set to s_end when the sub-group is done.  lcb_wait_progress(progress, words, n) on another
 * stream then releases the output projection h = m*W_proj (nnet/bilstm.py:128) of the finished frames -- and the half of the
 * next layer's input projection that reads them -- beside the running recurrence.  The launch never waits for its readers.
 * g_dtype (dtype codes of lcb_gemm16): 0 = G is fp32, 2 = G is fp16 -- lcb_gemm16 then writes half the bytes and this kernel
 * reloads half of them; the forget bias and the recurrent product are added in fp32 either way.
 * Mout_bf16 (NULL = none): [T*B, 2Hp] bf16 twin of Mout, written beside it -- the weight-gradient GEMMs take bf16 operands
 * (tcgen05 kind::f16 cannot mix fp16 x bf16), and a conversion pass over Mout costs a read and a launch per layer. */
int lcb_lstm_rec_fwd_progress_words(int B, int Hp, int num_dirs);
int lcb_lstm_rec_fwd_range_pg(const void* G, int g_dtype, const void* WfoldT, const float* peep, const int32_t* lens, const int32_t* lens_host,
                              const int32_t* ready_steps,
                              void* Mout, void* Mout_bf16, void* gates, float* cst, float* cfin, float* mfin,
                              int T, int B, int Hp, int num_dirs, float forget_bias, int s_begin, int s_end,
                              int32_t* progress, void* workspace, size_t workspace_bytes, void* stream);
/* BPTT of the above (replaces tf.gradients through the while_loop, nnet/graph.py:190-191).
 *   dM    [T*B, 2Hp] f32   d loss / d m_t arriving from the output projection
 *   Wfold [2*Hp, 4Hp] bf16 W' = W_proj*W_h per direction: rows = units, cols = packed gate columns
 *   dG    [T*B, 8Hp] bf16  d loss / d z_t (packed columns; 0 where t >= lens[b])
 *   dbias [2*4Hp] f32 +=,  dpeep [2,3,Hp] f32 += (NULL iff peep NULL) */
int lcb_lstm_rec_bwd(const float* dM, const void* gates, const float* cst, const void* Wfold, const float* peep,
                     const int32_t* lens, void* dG, float* dbias, float* dpeep,
                     int T, int B, int Hp, int num_dirs, void* workspace, size_t workspace_bytes, void* stream);
/* The same over scan steps [s_begin, s_end) only (scan step s visits frame T-1-s in the forward, s in the backward direction --
 * the reverse of the forward pass).  A launch with s_end < T leaves, per cell, the recurrent part of d loss / d m of the next
 * step and the carried d loss / d c in `carry` ([B,2,Hp,2] f32); a launch with s_begin > 0 resumes from them.  Launches over
 * consecutive ranges in stream order equal one launch over [0, T) bit for bit (dbias / dpeep accumulate), so the caller can
 * start BPTT when only the last frames' dM exist and compute the rest beside it.  LCB_ERR_UNSUPPORTED for a partial range
 * unless lcb_lstm_rec_bwd_can_split(Hp). */
int lcb_lstm_rec_bwd_range(const float* dM, const void* gates, const float* cst, const void* Wfold, const float* peep,
                           const int32_t* lens, void* dG, float* dbias, float* dpeep,
                           int T, int B, int Hp, int num_dirs, int s_begin, int s_end, float* carry,
                           void* workspace, size_t workspace_bytes, void* stream);
/* Same, publishing its progress: word (cluster, sub-group, CTA) of `progress` (lcb_lstm_rec_bwd_progress_words words, zeroed by the
 * caller before the launch; Hp = 512 only) = number of leading scan steps whose dG rows that CTA has written.  A caller that wants
 * the rows of frames final in both directions after scan step s -- [T-s, s) -- while the launch is still running enqueues
 * lcb_wait_progress(progress, words, s, other_stream) in front of the GEMMs that read them: ONE BPTT launch per layer instead of
 * one per release point.  The launch itself never waits on anything, so a serialised execution order cannot deadlock. */
int lcb_lstm_rec_bwd_range_pg(const float* dM, const void* gates, const float* cst, const void* Wfold, const float* peep,
                              const int32_t* lens, void* dG, float* dbias, float* dpeep,
                              int T, int B, int Hp, int num_dirs, int s_begin, int s_end, float* carry, int32_t* progress,
                              void* workspace, size_t workspace_bytes, void* stream);
int lcb_lstm_rec_bwd_progress_words(int B, int Hp, int num_dirs);
int lcb_wait_progress(const int32_t* progress, int n, int target, void* stream);
int lcb_lstm_rec_bwd_can_split(int Hp);

/* ---- HBM-bound helpers -----------------------------------------------------------------
 * lcb_pack_input: pipeline tensor nnet_input [B,T,D] f32 (nnet/pipeline.py:35-61) -> time-major
 *                 fp16 [T,B,Dp] (pad columns zero, saturating); replaces the batch-major walk of dynamic_rnn.
 * lcb_cast_f32_16 (dst_dtype 1 bf16 / 2 fp16) / lcb_split_f32_bf16: operand casts (x = hi + lo split
 *                 for fp32-accurate weight folds).
 * lcb_colsum: out[c] += sum_r src[r,c] (bias gradients). */
int lcb_pack_input(const float* nnet_input, void* x0, int B, int T, int D, int Dp, void* stream);
int lcb_cast_f32_16(const float* src, void* dst, int dst_dtype, size_t n, void* stream);
int lcb_split_f32_bf16(const float* src, void* hi, void* lo, size_t n, void* stream);
int lcb_f16_to_bf16(const void* src, void* dst, size_t n, void* stream);   /* n even */
/* in-place inverted dropout on a 16-bit tensor (dtype 1 bf16 / 2 fp16): DropoutWrapper(output_keep_prob)
 * on the LSTM layer outputs (nnet/bilstm.py:128,137); the same call on the gradient is its backward.
 * lcb_dropout_mask writes the 0/1 mask of (seed, index) as bytes. */
int lcb_dropout16(void* x, int dtype, size_t n, float keep_prob, unsigned long long seed, void* stream);
int lcb_dropout_mask(unsigned char* mask, size_t n, float keep_prob, unsigned long long seed, void* stream);
int lcb_colsum(const void* src, int src_dtype, int rows, int cols, int ld, float* out, void* stream);
/* y += x on fp16 tensors, n even: the layer-0 residual  finput = finput + concat(fwd, bwd)  (nnet/bilstm.py:199-200). */
int lcb_add_f16(void* y, const void* x, size_t n, void* stream);
/* y[r,c] += mask * x[r,c] / keep_prob on 16-bit 2-D tensors (dtype 1 bf16, 2 fp16; row strides ldy, ldx in elements), mask =
 * element (mask_base + r*ldm + c) of the (seed) stream lcb_gemm16_dropout applies to the same output.  The residual connection of
 * DropoutWrapper(ResidualWrapper(LSTMCell)), out = dropout(x + cell(x))  (nnet/lstm.py:236-260): forward adds the masked layer
 * input to the projected output, backward adds the masked output gradient to the input gradient. */
int lcb_masked_add16(void* y, int ldy, const void* x, int ldx, long long rows, int cols, int dtype, float keep_prob,
                     unsigned long long seed, unsigned long long mask_base, int ldm, void* stream);
/* label-smoothing regulariser (nnet/bilstm.py:254-269) over `rows` rows of logits [rows,V]:
 *   *loss_out += weight * sum p (log p - q),  dlogits (nullable) += its gradient;  q = log(1/V) when log_prior is
 *   NULL (uniform_label_sm) else log_prior[V] (prior_label_sm, nnet/class_prior.py).  All rows, padding included. */
int lcb_label_smooth(const float* logits, float* dlogits, long long rows, int V, float weight,
                     const float* log_prior, float* loss_out, void* stream);

/* ---- output layer: mixture of tanh-bounded expert logits, or affine -------------------------
 * lcb_output_fwd replaces create_moe (nnet/moe.py:29-72) / tf.nn.xw_plus_b (nnet/bilstm.py:249)
 * and the reshape to [B,T,V] (nnet/bilstm.py:250).  One fused tcgen05 GEMM + mixture epilogue; the
 * [N,K,V] expert tensor (moe.py:60) is never materialised.
 *   X      [T*B, ldx] fp16 time-major encoder output (D2 = 2*num_projects columns used)
 *   Wall   [K*V + K, D2] fp16: row v*K+k = column k*V+v of the reference's W (moe.py:50-58),
 *          rows K*V.. = W_prior^T (moe.py:34-41).  K == 0 (affine): Wall = W^T [V, D2].
 *   bias   [K*V + K] f32 in the same order (affine: [V])
 *   logits [B,T,V] f32 batch-major.   tau = moe_temperature.   K <= 128.
 *   keep_prob / seed: dropout on the mixture weights (moe.py:46) and on tau*tanh (moe.py:61), keep_prob = 1
 *   disables it.  Masks are a pure function of (seed, n, column): lcb_mos_bwd_dz regenerates them, and
 *   lcb_dropout_mask exports them (mixture weights: seed, index n*K+k; expert logits:
 *   seed ^ 0xD1B54A32D192ED03, index n*K*V + v*K+k).
 * lcb_mos_bwd_dz: rows [n0, n0+R) of Z = X*Wall^T + bias (recomputed with lcb_gemm16, [R, ldz] f32)
 *   and dlogits [B,T,V] f32 -> dZ [R, ldz] bf16 (same column order, pad columns zero); dX, dWall and
 *   dbias then follow from lcb_gemm16 / lcb_colsum (tf.gradients, nnet/graph.py:190-191).
 * lcb_pack_dlogits: [B,T,V] f32 -> time-major [T*B, ldo] bf16 (the affine layer's dZ). */
int lcb_output_fwd(const void* X, int ldx, const void* Wall, const float* bias, float* logits,
                   int T, int B, int D2, int V, int K, float tau, float keep_prob, unsigned long long seed,
                   void* stream);
int lcb_mos_bwd_dz(const float* Z, const float* dlogits, void* dZ, int n0, int R, int ldz,
                   int T, int B, int V, int K, float tau, float keep_prob, unsigned long long seed, void* stream);
int lcb_pack_dlogits(const float* dlogits, void* out, int T, int B, int V, int ldo, void* stream);

/* ---- fused L2 + global-norm clip + optimizer on flat fp32 buffers ---------------------------
 * replaces nnet/graph.py:183-200: g += l2*w outside the no-decay ranges (LSTM biases, :186),
 * clip_by_global_norm (:190-192), and apply_gradients for opt = 0 sgd / 1 momentum / 2 adam
 * (tf.train.* semantics, :37-48).  step >= 1 is the Adam time step.  nodecay_ranges_host: n_ranges
 * [lo,hi) element ranges (HOST pointer, <= 24).  sumsq_scratch: device double.  gnorm_out: device
 * float receiving the pre-clip global norm (nullable).  No host synchronisation. */
int lcb_optimizer_step(float* w, float* g, float* s1, float* s2, long long n, int opt,
                       float lr, long long step, float beta1, float beta2, float eps, float momentum,
                       float l2, float clip_norm, const long long* nodecay_ranges_host, int n_ranges,
                       double* sumsq_scratch, float* gnorm_out, void* stream);

/* ---- validation / inference tails ---------------------------------------------------------------
 * lcb_greedy_decode replaces tf.nn.ctc_greedy_decoder(merge_repeated=True) (nnet/graph.py:138-142):
 *   out [B,T] int32 receives the collapsed label sequence of each utterance, out_len [B] its length.
 * lcb_posterior replaces tf.nn.softmax(smooth_factor*logits) (nnet/graph.py:236) and the numpy.log /
 *   class-prior subtraction of bin/nnet-forward.py:87-91 (log_prior nullable, [V]); blank_to_front != 0 also moves
 *   column V-1 to column 0 (the `select-feats` reorder of scripts/decode_ctc_lat.sh:161-163).  out must not alias logits
 *   when blank_to_front is set. */
int lcb_greedy_decode(const float* logits, const int32_t* seq_len, int32_t* out, int32_t* out_len,
                      int B, int T, int V, void* stream);
int lcb_posterior(const float* logits, float* out, long long rows, int V, float smooth_factor,
                  int apply_log, const float* log_prior, int blank_to_front, void* stream);

/* ---- data formats either side of the path ------------------------------------
 * replaces: _splice / _subsample of nnet/tfrecord.py:28-51 (edge-replicated +-context splicing, then every
 *           subsample-th frame; output length floor(len / subsample)), applied to the zero-padded minibatch ON THE DEVICE.
 *   in   [B, T, D] f32, lens [B] int32  ->  out [B, Tout, D*(1+lc+rc)] f32 (0 past the new length), lens_out [B] (nullable)
 *   subsample <= 1: no subsampling.  Tout >= T / max(subsample, 1). */
int lcb_splice_subsample(const float* in, const int32_t* lens, float* out, int32_t* lens_out,
                         int B, int T, int D, int left_context, int right_context, int subsample, int Tout, void* stream);
/* CRC-32C (Castagnoli, reflected 0x82F63B78) of a HOST buffer, continuing from `crc` (0 to start): the checksum of the
 * TFRecord framing written by tf.python_io.TFRecordWriter (nnet/tfrecord.py:132) -- masked as ((c >> 15 | c << 17) + 0xa282ead8). */
uint32_t lcb_crc32c(const void* data, size_t n, uint32_t crc);
/* tf.train.SequenceExample decoder on HOST buffers: what tf.parse_single_sequence_example yields for the FixedLenSequenceFeature
 * specs of nnet/tfrecord.py:96-106 -- feature_lists "nnet_input" (one float_list Feature per frame) -> x_out [rows, cols] f32,
 * "nnet_target" (int64_list) -> y_out [num_labels].  Two calls: with x_out = y_out = NULL it only sizes (rows, cols,
 * num_labels -- -1 when the record has no "nnet_target" list); the fill call passes the buffers, their capacities in elements, and *cols as returned by the sizing call.
 * LCB_ERR_BAD_SHAPE: malformed message or frames of unequal width. */
int lcb_parse_sequence_example(const void* buf, size_t n, float* x_out, size_t x_cap, int64_t* y_out, size_t y_cap,
                               long long* rows, long long* cols, long long* num_labels);

#ifdef __cplusplus
}
#endif
#endif /* LSTM_CTC_B200_H_ */
