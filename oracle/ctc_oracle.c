/*
 * oracle/ctc_oracle.c  --  TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement of the CTC loss + gradient that the reference obtains from
 * TensorFlow r1.8's `tf.nn.ctc_loss` (call site: /root/reference/nnet/graph.py:109-114,
 * followed by reduce_sum at :116).  TensorFlow itself is a third-party dependency that is
 * NOT under /root/reference (pinned only in prose: tensorflow_gpu-1.8.0-cp27,
 * /root/reference/README.md:6,23), so this file restates the published algorithm of
 * tensorflow/core/util/ctc/ctc_loss_calculator.{h,cc} (Graves et al. 2006, eq. 6-16):
 *
 *   - blank index = V-1 (TF convention; /root/reference/egs/wsj/run_wsj_phn.sh:128)
 *   - y_t = softmax(logits[t,b,:]) for t < seq_len[b]
 *   - l' = [blank, l1, blank, ..., lL, blank], S = 2L+1
 *   - alpha includes the emission at t, beta does NOT (TF's convention), both in log space
 *   - loss_b = -log(alpha(S-1,T-1) + alpha(S-2,T-1))
 *   - dlogits[t,v] = y_t(v) - (1/p) * sum_{s: l'_s = v} alpha_t(s) beta_t(s);  0 for t >= seq_len[b]
 *   - ignore_longer_outputs_than_inputs=True (graph.py:113): L_b > T_b  -> loss 0, grad 0
 *   - T_b == 0 -> loss 0, grad 0
 *   - feasible length but no valid path (repeats) -> loss = +inf, grad = y
 *   - preprocess_collapse_repeated=False, ctc_merge_repeated=True (TF defaults)
 *
 * Arithmetic is double precision (TF uses float); the parity tests compare the CUDA path with
 * this oracle at the tolerance north_star states (1e-4 relative).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may load this file.
 * PARITY PINNING: there are no golden vectors in the reference tree; this oracle is pinned by
 * the two known-answer vectors of upstream TF's ctc_loss_op_test.py (tests/golden/ctc_tf_kat.json)
 * and by agreement with torch.nn.functional.ctc_loss and brute-force path enumeration.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

static inline double log_add(double a, double b) {
    if (a == -INFINITY) return b;
    if (b == -INFINITY) return a;
    return a > b ? a + log1p(exp(b - a)) : b + log1p(exp(a - b));
}

/* logits: [B,T,V] (batch-major, as create_logits_blstm returns them: nnet/bilstm.py:250)
 * labels: [B,Lmax] int64, -1 padded (nnet/pipeline.py:43); non-(-1) entries are taken in order
 *         (the dense->sparse conversion of nnet/graph.py:74-104)
 * returns 0 on success, -1 on an invalid label (TF raises InvalidArgument). */
int ctc_oracle_f64(const double* logits, const int64_t* labels, int Lmax,
                   const int32_t* seq_len, int B, int T, int V,
                   double* loss, double* grad, int nthreads)
{
    const int blank = V - 1;
    int err = 0;
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#pragma omp parallel for schedule(dynamic)
#endif
    for (int b = 0; b < B; ++b) {
        const double* x = logits + (size_t)b * T * V;
        double* g = grad + (size_t)b * T * V;
        memset(g, 0, sizeof(double) * (size_t)T * V);
        loss[b] = 0.0;
        int Tb = seq_len[b];
        if (Tb > T) Tb = T;
        int* lab = (int*)malloc(sizeof(int) * (Lmax > 0 ? Lmax : 1));
        int L = 0;
        for (int i = 0; i < Lmax; ++i) {
            int64_t v = labels[(size_t)b * Lmax + i];
            if (v == -1) continue;
            if (v < 0 || v >= blank) { err = -1; }
            lab[L++] = (int)v;
        }
        if (err || Tb <= 0 || L > Tb) { free(lab); continue; }
        const int S = 2 * L + 1;
        double* lp = (double*)malloc(sizeof(double) * (size_t)Tb * V);   /* log softmax */
        double* al = (double*)malloc(sizeof(double) * (size_t)Tb * S);
        double* be = (double*)malloc(sizeof(double) * (size_t)Tb * S);
        for (int t = 0; t < Tb; ++t) {
            double m = -INFINITY;
            for (int v = 0; v < V; ++v) if (x[(size_t)t * V + v] > m) m = x[(size_t)t * V + v];
            double s = 0.0;
            for (int v = 0; v < V; ++v) s += exp(x[(size_t)t * V + v] - m);
            double lse = m + log(s);
            for (int v = 0; v < V; ++v) lp[(size_t)t * V + v] = x[(size_t)t * V + v] - lse;
        }
#define LPRIME(s) (((s) & 1) ? lab[(s) >> 1] : blank)
        for (int s = 0; s < S; ++s) al[s] = -INFINITY;
        al[0] = lp[blank];
        if (S > 1) al[1] = lp[lab[0]];
        for (int t = 1; t < Tb; ++t) {
            for (int s = 0; s < S; ++s) {
                double a = al[(size_t)(t - 1) * S + s];
                if (s >= 1) a = log_add(a, al[(size_t)(t - 1) * S + s - 1]);
                if (s >= 2 && (s & 1) && LPRIME(s) != LPRIME(s - 2))
                    a = log_add(a, al[(size_t)(t - 1) * S + s - 2]);
                al[(size_t)t * S + s] = (a == -INFINITY) ? a : a + lp[(size_t)t * V + LPRIME(s)];
            }
        }
        for (int s = 0; s < S; ++s) be[(size_t)(Tb - 1) * S + s] = -INFINITY;
        be[(size_t)(Tb - 1) * S + S - 1] = 0.0;
        if (S > 1) be[(size_t)(Tb - 1) * S + S - 2] = 0.0;
        for (int t = Tb - 2; t >= 0; --t) {
            for (int s = 0; s < S; ++s) {
                double a = be[(size_t)(t + 1) * S + s];
                if (a != -INFINITY) a += lp[(size_t)(t + 1) * V + LPRIME(s)];
                if (s + 1 < S) {
                    double c = be[(size_t)(t + 1) * S + s + 1];
                    if (c != -INFINITY) a = log_add(a, c + lp[(size_t)(t + 1) * V + LPRIME(s + 1)]);
                }
                if (s + 2 < S && (s & 1) && LPRIME(s) != LPRIME(s + 2)) {
                    double c = be[(size_t)(t + 1) * S + s + 2];
                    if (c != -INFINITY) a = log_add(a, c + lp[(size_t)(t + 1) * V + LPRIME(s + 2)]);
                }
                be[(size_t)t * S + s] = a;
            }
        }
        double logp = al[(size_t)(Tb - 1) * S + S - 1];
        if (S > 1) logp = log_add(logp, al[(size_t)(Tb - 1) * S + S - 2]);
        loss[b] = -logp;
        for (int t = 0; t < Tb; ++t)
            for (int v = 0; v < V; ++v) g[(size_t)t * V + v] = exp(lp[(size_t)t * V + v]);
        if (logp != -INFINITY) {
            for (int t = 0; t < Tb; ++t)
                for (int s = 0; s < S; ++s) {
                    double ab = al[(size_t)t * S + s] + be[(size_t)t * S + s];
                    if (ab != -INFINITY) g[(size_t)t * V + LPRIME(s)] -= exp(ab - logp);
                }
        }
#undef LPRIME
        free(lp); free(al); free(be); free(lab);
    }
    return err;
}

/* float32 entry with the same algorithm (double internally); used as the timed CPU baseline
 * ("port" of the TF CPU kernel, batch sharded over host threads like TF's intra-op pool). */
int ctc_oracle_f32(const float* logits, const int64_t* labels, int Lmax,
                   const int32_t* seq_len, int B, int T, int V,
                   float* loss, float* grad, int nthreads)
{
    size_t n = (size_t)B * T * V;
    double* x = (double*)malloc(sizeof(double) * n);
    double* g = (double*)malloc(sizeof(double) * n);
    double* l = (double*)malloc(sizeof(double) * (size_t)B);
    for (size_t i = 0; i < n; ++i) x[i] = logits[i];
    int rc = ctc_oracle_f64(x, labels, Lmax, seq_len, B, T, V, l, g, nthreads);
    for (size_t i = 0; i < n; ++i) grad[i] = (float)g[i];
    for (int b = 0; b < B; ++b) loss[b] = (float)l[b];
    free(x); free(g); free(l);
    return rc;
}
