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
/* 0, or LCB_ERR_INVALID_LABEL if the last call on this workspace saw an out-of-range label.
 * Synchronises `stream`. */
int lcb_ctc_status(const void* workspace, void* stream);

/* ---- bf16 GEMM, fp32 accumulate (tcgen05 + TMA) ----------------------------------------
 * replaces: the matmul inside tf.contrib.rnn.LSTMCell hoisted over all frames
 *           (nnet/bilstm.py:129-136,171-188), the LSTM projection, tf.nn.xw_plus_b
 *           (nnet/bilstm.py:249, nnet/moe.py:42,59) and their tf.gradients counterparts
 *           (nnet/graph.py:190-191).
 * C[M,N] (+)= op(A)[M,K] * op(B)[K,N] + bias[N]
 *   a_layout 0: A stored row-major [M,K] (ld = lda)      1: A stored row-major [K,M]
 *   b_layout 0: B stored row-major [N,K] (ld = ldb)      1: B stored row-major [K,N]
 *   c_dtype  0: C fp32                                    1: C bf16
 * A, B bf16; lda/ldb multiples of 8 elements, base pointers 16-byte aligned.
 * accumulate != 0 adds into the existing fp32 C (c_dtype must be 0). */
int lcb_gemm_bf16(int M, int N, int K,
                  const void* A, int lda, int a_layout,
                  const void* B, int ldb, int b_layout,
                  void* C, int ldc, int c_dtype,
                  const float* bias, int accumulate, void* stream);
/* Same contract on plain CUDA cores -- a slow on-device CHECKER for tests, not a product path. */
int lcb_gemm_bf16_simt_check(int M, int N, int K,
                             const void* A, int lda, int a_layout,
                             const void* B, int ldb, int b_layout,
                             void* C, int ldc, int c_dtype,
                             const float* bias, int accumulate, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* LSTM_CTC_B200_H_ */
