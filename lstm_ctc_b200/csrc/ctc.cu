// ctc.cu -- CTC loss + d(loss)/d(logits) for sm_100a (K4 in DESIGN.md).
//
// Replaces the tf.nn.ctc_loss call of /root/reference/nnet/graph.py:109-114 (CPU-only op in
// TF 1.8) and the [B,T,V]->[T,B,V] transpose before it (graph.py:72): logits are consumed
// batch-major exactly as create_logits_blstm returns them.
//
// Three launches, one stream:
//   ctc_prep      : per utterance, compact the -1-padded dense labels (graph.py:74-104), validate
//                   them, decide "skipped" (T_b==0 or L_b>T_b: ignore_longer_outputs_than_inputs).
//   ctc_softmax   : HBM-bound streaming pass, one warp (or CTA) per frame row: log-sum-exp,
//                   grad[b,t,:] = softmax (0 past seq_len / skipped utts), and the compact
//                   per-frame log-probs the lattice needs: lpb[b,t] (blank) and lpl[b,t,j] (label j).
//                   Reads logits once, writes grad once: the 8*T*B*V algorithmic bytes.
//   ctc_alpha_beta: latency-bound lattice pass, TWO CTAs (1..32 warps each) per utterance running
//                   CONCURRENTLY: CTA 2b sweeps alpha forward in time, CTA 2b+1 sweeps beta backward
//                   (the two recursions are independent; only gamma needs both), so the serial chain
//                   is T steps instead of 2T.  SPT consecutive lattice states per thread in
//                   registers; log-space values are kept in fp64 while exp/log run in fp32 on
//                   max-subtracted differences, so the absolute error per step is ~1e-7 regardless
//                   of |alpha| (plain fp32 log-space, which is what TF does, loses 1e-4..1e-3 at
//                   T~3000).  Both sweeps spill their lattice rows to the workspace.
//   ctc_gamma     : one warp per frame row: gamma = exp(alpha + beta - log p), grad[b,t,l'_s] -= gamma
//                   with red.global.add (blank contributions pre-summed per warp).
#include "ptx.cuh"
#include "lstm_ctc_b200.h"

namespace lcb {

constexpr double CTC_NEG = -1.0e30;       // log(0) sentinel (finite: no inf-inf NaNs)
constexpr double CTC_ZERO_THRESH = -1.0e29;

struct CtcMeta {   // per utterance, in workspace
    int L;         // number of labels (non -1 entries)
    int Tb;        // min(seq_len, T)
    int skip;      // 1 -> loss 0, grad 0
    int pad;
};

// --------------------------------------------------------------------------------------------
__global__ void ctc_prep_kernel(const int64_t* __restrict__ labels, int Lmax, const int32_t* __restrict__ seq_len,
                                int B, int T, int V, CtcMeta* __restrict__ meta, int* __restrict__ lab_out, int LABP,
                                int* __restrict__ status)
{
    int b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    int lane = threadIdx.x & 31;
    if (b >= B) return;
    const int blank = V - 1;
    int L = 0;
    bool bad = false;
    for (int base = 0; base < Lmax; base += 32) {
        int i = base + lane;
        long long v = (i < Lmax) ? labels[(size_t)b * Lmax + i] : -1;
        bool keep = (v != -1);
        if (keep && (v < 0 || v >= blank)) bad = true;
        unsigned m = __ballot_sync(0xffffffffu, keep);
        if (keep) lab_out[(size_t)b * LABP + L + __popc(m & ((1u << lane) - 1))] = (int)v;
        L += __popc(m);
    }
    bad = __any_sync(0xffffffffu, bad);
    if (lane == 0) {
        int Tb = seq_len[b];
        Tb = Tb > T ? T : Tb;
        CtcMeta mt;
        mt.L = L; mt.Tb = Tb < 0 ? 0 : Tb;
        mt.skip = (Tb <= 0 || L > Tb || bad) ? 1 : 0;
        mt.pad = 0;
        meta[b] = mt;
        if (bad) atomicCAS(status, 0, LCB_ERR_INVALID_LABEL);
    }
}

// --------------------------------------------------------------------------------------------
// softmax / gather pass.  GROUP threads cooperate on one row; each thread caches NCH chunks of VEC
// consecutive floats.  Rows up to GROUP*VEC*NCH elements.
template <int VEC> struct VecT;
template <> struct VecT<1> { using type = float; };
template <> struct VecT<2> { using type = float2; };
template <> struct VecT<4> { using type = float4; };

template <int GROUP>
__device__ __forceinline__ float group_max(float v, float* red, int tid_in_group) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    if constexpr (GROUP > 32) {
        __syncthreads();
        if ((tid_in_group & 31) == 0) red[tid_in_group >> 5] = v;
        __syncthreads();
        v = red[0];
#pragma unroll
        for (int w = 1; w < GROUP / 32; ++w) v = fmaxf(v, red[w]);
    }
    return v;
}
template <int GROUP>
__device__ __forceinline__ float group_sum(float v, float* red, int tid_in_group) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if constexpr (GROUP > 32) {
        __syncthreads();
        if ((tid_in_group & 31) == 0) red[tid_in_group >> 5] = v;
        __syncthreads();
        v = red[0];
#pragma unroll
        for (int w = 1; w < GROUP / 32; ++w) v += red[w];
    }
    return v;
}

template <int GROUP, int VEC, int NCH>
__global__ void __launch_bounds__(256)
ctc_softmax_kernel(const float* __restrict__ logits, float* __restrict__ grad, int B, int T, int V,
                   const CtcMeta* __restrict__ meta, const int* __restrict__ lab, int LABP,
                   float* __restrict__ lpb, float* __restrict__ lpl, int LPP)
{
    using VT = typename VecT<VEC>::type;
    __shared__ float red[8];
    constexpr int GROUPS_PER_CTA = 256 / GROUP;
    const int g_in_cta = threadIdx.x / GROUP;
    const int tig = threadIdx.x % GROUP;
    const long long nrows = (long long)B * T;
    const int blank = V - 1;
    const float LOG2E = 1.4426950408889634f;

    for (long long row = (long long)blockIdx.x * GROUPS_PER_CTA + g_in_cta; row < nrows;
         row += (long long)gridDim.x * GROUPS_PER_CTA) {
        const int b = (int)(row / T);
        const int t = (int)(row % T);
        const CtcMeta mt = meta[b];
        const float* x = logits + (size_t)row * V;
        float* g = grad + (size_t)row * V;
        const bool live = (!mt.skip) && (t < mt.Tb);     // uniform across the group
        if (!live) {
#pragma unroll
            for (int c = 0; c < NCH; ++c) {
                int e = (c * GROUP + tig) * VEC;
                if (e < V) {
                    if constexpr (VEC == 4) *reinterpret_cast<float4*>(g + e) = make_float4(0.f, 0.f, 0.f, 0.f);
                    else if constexpr (VEC == 2) *reinterpret_cast<float2*>(g + e) = make_float2(0.f, 0.f);
                    else g[e] = 0.f;
                }
            }
            if constexpr (GROUP > 32) { /* keep barrier counts uniform: nothing to do, no barriers taken */ }
            continue;
        }
        float v[NCH][VEC];
        float mx = -INFINITY;
#pragma unroll
        for (int c = 0; c < NCH; ++c) {
            int e = (c * GROUP + tig) * VEC;
            if (e < V) {
                VT tmp = __ldg(reinterpret_cast<const VT*>(x + e));
                const float* tp = reinterpret_cast<const float*>(&tmp);
#pragma unroll
                for (int k = 0; k < VEC; ++k) { v[c][k] = tp[k]; mx = fmaxf(mx, tp[k]); }
            } else {
#pragma unroll
                for (int k = 0; k < VEC; ++k) v[c][k] = -INFINITY;
            }
        }
        mx = group_max<GROUP>(mx, red, tig);
        float s = 0.f;
        const float mxl = mx * LOG2E;
#pragma unroll
        for (int c = 0; c < NCH; ++c)
#pragma unroll
            for (int k = 0; k < VEC; ++k) { v[c][k] = exp2f(fmaf(v[c][k], LOG2E, -mxl)); s += v[c][k]; }
        s = group_sum<GROUP>(s, red, tig);
        const float inv = 1.0f / s;
        const float lse = mx + logf(s);
#pragma unroll
        for (int c = 0; c < NCH; ++c) {
            int e = (c * GROUP + tig) * VEC;
            if (e < V) {
                if constexpr (VEC == 4) *reinterpret_cast<float4*>(g + e) = make_float4(v[c][0] * inv, v[c][1] * inv, v[c][2] * inv, v[c][3] * inv);
                else if constexpr (VEC == 2) *reinterpret_cast<float2*>(g + e) = make_float2(v[c][0] * inv, v[c][1] * inv);
                else g[e] = v[c][0] * inv;
            }
        }
        // compact log-probs for the lattice (re-reads hit L1/L2: the row was just streamed)
        if (tig == 0) lpb[row] = x[blank] - lse;
        const int* lb = lab + (size_t)b * LABP;
        float* lo = lpl + (size_t)row * LPP;
        for (int j = tig; j < mt.L; j += GROUP) lo[j] = x[lb[j]] - lse;
    }
}

// Fallback for rows longer than the register cache: CTA per row, three passes over global
// (passes 2 and 3 hit L2).
__global__ void __launch_bounds__(256)
ctc_softmax_bigrow_kernel(const float* __restrict__ logits, float* __restrict__ grad, int B, int T, int V,
                          const CtcMeta* __restrict__ meta, const int* __restrict__ lab, int LABP,
                          float* __restrict__ lpb, float* __restrict__ lpl, int LPP)
{
    __shared__ float red[8];
    const long long nrows = (long long)B * T;
    const int blank = V - 1;
    for (long long row = blockIdx.x; row < nrows; row += gridDim.x) {
        const int b = (int)(row / T), t = (int)(row % T);
        const CtcMeta mt = meta[b];
        const float* x = logits + (size_t)row * V;
        float* g = grad + (size_t)row * V;
        const bool live = (!mt.skip) && (t < mt.Tb);
        if (!live) { for (int e = threadIdx.x; e < V; e += 256) g[e] = 0.f; continue; }
        float mx = -INFINITY;
        for (int e = threadIdx.x; e < V; e += 256) mx = fmaxf(mx, x[e]);
        mx = group_max<256>(mx, red, threadIdx.x);
        float s = 0.f;
        for (int e = threadIdx.x; e < V; e += 256) s += __expf(x[e] - mx);
        s = group_sum<256>(s, red, threadIdx.x);
        const float lse = mx + logf(s);
        for (int e = threadIdx.x; e < V; e += 256) g[e] = __expf(x[e] - lse);
        if (threadIdx.x == 0) lpb[row] = x[blank] - lse;
        const int* lb = lab + (size_t)b * LABP;
        float* lo = lpl + (size_t)row * LPP;
        for (int j = threadIdx.x; j < mt.L; j += 256) lo[j] = x[lb[j]] - lse;
        __syncthreads();
    }
}

// --------------------------------------------------------------------------------------------
// lattice pass
__device__ __forceinline__ double lse3(double a0, double a1, double a2) {
    double m = fmax(a0, fmax(a1, a2));
    float s = __expf((float)(a0 - m)) + __expf((float)(a1 - m)) + __expf((float)(a2 - m));
    return m + (double)__logf(s);
}
__device__ __forceinline__ double lse2(double a0, double a1) {
    double m = fmax(a0, a1);
    float s = __expf((float)(a0 - m)) + __expf((float)(a1 - m));
    return m + (double)__logf(s);
}

template <int NW>
__device__ __forceinline__ void cta_sync() {
    if constexpr (NW == 1) __syncwarp(); else __syncthreads();
}

// smem layout per CTA (dynamic):
//   nb   : double2[2][NT]                      neighbour exchange, parity double-buffered
//   lpL  : float [2][TC][NT][SPT/2]            per-thread-private staged label log-probs
//   lpB  : float [2][TC]  (each thread reads the broadcast copy; staged by thread 0..TC-1)
template <int NW, int SPT>
__global__ void __launch_bounds__(NW * 32)
ctc_alpha_beta_kernel(const CtcMeta* __restrict__ meta, const int* __restrict__ lab, int LABP,
                      const float* __restrict__ lpb, const float* __restrict__ lpl, int LPP,
                      double* __restrict__ alpha_ws, double* __restrict__ beta_ws, double* __restrict__ logp_ws,
                      float* __restrict__ loss, int T, int V, int TC)
{
    constexpr int NT = NW * 32;
    constexpr int HL = SPT / 2;          // label slots per thread
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double2* nb = reinterpret_cast<double2*>(smem_raw);                      // [2][NT]
    float* lpL = reinterpret_cast<float*>(nb + 2 * NT);                      // [2][TC][NT][HL]
    float* lpB = lpL + (size_t)2 * TC * NT * HL;                             // [2][TC] (padded to 4)
    __shared__ double s_fin[2];

    const int b = blockIdx.x >> 1;
    const int which = blockIdx.x & 1;                 // 0: alpha sweep (+ loss), 1: beta sweep
    const int tid = threadIdx.x;
    const CtcMeta mt = meta[b];
    if (mt.skip) { if (tid == 0 && which == 0) { loss[b] = 0.f; logp_ws[b] = CTC_NEG; } return; }
    const int Tb = mt.Tb, L = mt.L, S = 2 * L + 1;
    const int s0 = tid * SPT;
    const int* lb = lab + (size_t)b * LABP;
    const float* lpb_b = lpb + (size_t)b * T;
    const float* lpl_b = lpl + (size_t)b * T * LPP;
    double* aw = (which ? beta_ws : alpha_ws) + (size_t)b * T * (NT * SPT);

    // per-state constants: odd local index i <-> label slot i/2 (s0 is even since SPT is even)
    int labv[HL];
    bool skipf[HL];     // alpha: s-2 -> s allowed for odd state 2*j+1 (j = s0/2 + k): lab[j] != lab[j-1]
    bool skipb[HL];     // beta : s -> s+2 allowed: lab[j+1] != lab[j]
#pragma unroll
    for (int k = 0; k < HL; ++k) {
        int j = s0 / 2 + k;
        int l0 = (j < L) ? lb[j] : -1;
        labv[k] = l0;
        skipf[k] = (j < L) && (j >= 1) && (lb[j - 1] != l0);
        skipb[k] = (j + 1 < L) && (lb[j + 1] != l0);
    }

    // ---- staging helpers (per-thread-private cp.async; no CTA barrier needed for visibility) ----
    auto stage_lp = [&](int buf, int t_first, int nfr) {      // frames t_first .. t_first+nfr-1
        for (int f = 0; f < nfr; ++f) {
            const float* src = lpl_b + (size_t)(t_first + f) * LPP + s0 / 2;
            float* dst = lpL + (((size_t)buf * TC + f) * NT + tid) * HL;
            if (s0 / 2 < LPP) cp_async<HL * 4>(dst, src);
        }
        if (tid < nfr) cp_async<4>(lpB + buf * ((TC + 3) & ~3) + tid, lpb_b + t_first + tid);
    };

    if (which == 0) {
    double a[SPT];
    // =============================== alpha sweep ===============================
    {
        const int nchunks = (Tb + TC - 1) / TC;
        stage_lp(0, 0, min(TC, Tb));
        cp_async_commit();
        int par = 0;
        for (int c = 0; c < nchunks; ++c) {
            const int t_first = c * TC;
            const int nfr = min(TC, Tb - t_first);
            if (c + 1 < nchunks) stage_lp((c + 1) & 1, t_first + TC, min(TC, Tb - t_first - TC));
            cp_async_commit();
            cp_async_wait<1>();
            cta_sync<NW>();          // lpB is staged by other threads
            const int buf = c & 1;
            for (int f = 0; f < nfr; ++f) {
                const int t = t_first + f;
                const float lb_t = lpB[buf * ((TC + 3) & ~3) + f];
                float ll[HL];
#pragma unroll
                for (int k = 0; k < HL; ++k) {
                    ll[k] = lpL[(((size_t)buf * TC + f) * NT + tid) * HL + k];
                    if (s0 / 2 + k >= L) ll[k] = 0.f;       // slots past L hold uninitialised workspace
                }
                if (t == 0) {
#pragma unroll
                    for (int i = 0; i < SPT; ++i) a[i] = CTC_NEG;
                    if (s0 == 0) { a[0] = (double)lb_t; if (L > 0) a[1] = (double)ll[0]; }
                } else {
                    // neighbour values (states s0-1, s0-2) from thread tid-1
                    nb[par * NT + tid] = make_double2(a[SPT - 1], a[SPT - 2]);
                    cta_sync<NW>();
                    double p1 = CTC_NEG, p2 = CTC_NEG;
                    if (tid > 0) { double2 q = nb[par * NT + tid - 1]; p1 = q.x; p2 = q.y; }
                    par ^= 1;
                    double na[SPT];
#pragma unroll
                    for (int i = 0; i < SPT; ++i) {
                        const double am1 = (i >= 1) ? a[i - 1] : p1;
                        const double am2 = (i >= 2) ? a[i - 2] : ((i == 1) ? p1 : p2);
                        double r;
                        if (i & 1) {     // label state
                            const int k = i >> 1;
                            r = lse3(a[i], am1, skipf[k] ? am2 : CTC_NEG) + (double)ll[k];
                        } else {         // blank state
                            r = lse2(a[i], am1) + (double)lb_t;
                        }
                        na[i] = (s0 + i < S) ? r : CTC_NEG;
                    }
#pragma unroll
                    for (int i = 0; i < SPT; ++i) a[i] = na[i];
                }
                // spill alpha_t
                double* dst = aw + (size_t)t * (NT * SPT) + s0;
#pragma unroll
                for (int q = 0; q < SPT / 2; ++q) *reinterpret_cast<double2*>(dst + 2 * q) = make_double2(a[2 * q], a[2 * q + 1]);
            }
            cta_sync<NW>();          // all reads of this lpB buffer done before it is restaged
        }
        cp_async_wait<0>();
    }
    // ---- log p = lse(alpha_{T-1}(S-1), alpha_{T-1}(S-2)) ----
#pragma unroll
    for (int i = 0; i < SPT; ++i) {
        if (s0 + i == S - 1) s_fin[0] = a[i];
        if (s0 + i == S - 2) s_fin[1] = a[i];
    }
    if (S == 1 && tid == 0) s_fin[1] = CTC_NEG;
    __syncthreads();
    if (tid == 0) {
        double lp = lse2(s_fin[0], s_fin[1]);
        logp_ws[b] = lp;                            // <= CTC_ZERO_THRESH: no valid path, grad stays = softmax (TF behaviour)
        loss[b] = (lp <= CTC_ZERO_THRESH) ? INFINITY : (float)(-lp);
    }
    return;
    }

    // =============================== beta sweep ===============================
    // e[i] = beta_t(s) + lp_t(l'_s)  ("beta with emission");  beta_{t-1}(s) = lse(e(s), e(s+1), skip ? e(s+2))
    {
        double be[SPT];
#pragma unroll
        for (int i = 0; i < SPT; ++i) be[i] = (s0 + i == S - 1 || s0 + i == S - 2) ? 0.0 : CTC_NEG;
        const int nchunks = (Tb + TC - 1) / TC;
        // chunk c covers frames [Tb - (c+1)*TC, Tb - c*TC) clipped at 0, processed descending
        auto chunk_first = [&](int c) { int f = Tb - (c + 1) * TC; return f < 0 ? 0 : f; };
        auto chunk_n = [&](int c) { return (Tb - c * TC) - chunk_first(c); };
        stage_lp(0, chunk_first(0), chunk_n(0));
        cp_async_commit();
        int par = 0;
        for (int c = 0; c < nchunks; ++c) {
            const int t_first = chunk_first(c);
            const int nfr = chunk_n(c);
            if (c + 1 < nchunks) stage_lp((c + 1) & 1, chunk_first(c + 1), chunk_n(c + 1));
            cp_async_commit();
            cp_async_wait<1>();
            cta_sync<NW>();
            const int buf = c & 1;
            for (int f = nfr - 1; f >= 0; --f) {
                const int t = t_first + f;
                const float lb_t = lpB[buf * ((TC + 3) & ~3) + f];
                float ll[HL];
#pragma unroll
                for (int k = 0; k < HL; ++k) {
                    ll[k] = lpL[(((size_t)buf * TC + f) * NT + tid) * HL + k];
                    if (s0 / 2 + k >= L) ll[k] = 0.f;       // slots past L hold uninitialised workspace
                }
                {   // spill beta_t (without the emission at t, as TF defines it)
                    double* dst = aw + (size_t)t * (NT * SPT) + s0;
#pragma unroll
                    for (int q = 0; q < SPT / 2; ++q) *reinterpret_cast<double2*>(dst + 2 * q) = make_double2(be[2 * q], be[2 * q + 1]);
                }
                if (t == 0) break;
                // e = beta_t + lp_t ; then beta_{t-1}
                double e[SPT];
#pragma unroll
                for (int i = 0; i < SPT; ++i) e[i] = (s0 + i < S) ? be[i] + (double)((i & 1) ? ll[i >> 1] : lb_t) : CTC_NEG;
                nb[par * NT + tid] = make_double2(e[0], e[1]);
                cta_sync<NW>();
                double n1 = CTC_NEG, n2 = CTC_NEG;        // e(s0+SPT), e(s0+SPT+1) from thread tid+1
                if (tid + 1 < NT) { double2 q = nb[par * NT + tid + 1]; n1 = q.x; n2 = q.y; }
                par ^= 1;
#pragma unroll
                for (int i = 0; i < SPT; ++i) {
                    const double ep1 = (i + 1 < SPT) ? e[i + 1] : n1;
                    const double ep2 = (i + 2 < SPT) ? e[i + 2] : ((i + 2 == SPT) ? n1 : n2);
                    double r;
                    if (i & 1) r = lse3(e[i], ep1, skipb[i >> 1] ? ep2 : CTC_NEG);
                    else r = lse2(e[i], ep1);
                    be[i] = (s0 + i < S) ? r : CTC_NEG;
                }
            }
            cta_sync<NW>();
        }
        cp_async_wait<0>();
    }
}

// --------------------------------------------------------------------------------------------
// gamma_t(s) = exp(alpha_t(s) + beta_t(s) - log p);  grad[b,t,l'_s] -= gamma_t(s).  One warp per frame row.
__global__ void __launch_bounds__(256)
ctc_gamma_kernel(const CtcMeta* __restrict__ meta, const int* __restrict__ lab, int LABP,
                 const double* __restrict__ alpha_ws, const double* __restrict__ beta_ws, const double* __restrict__ logp_ws,
                 float* __restrict__ grad, int B, int T, int V, int NS)
{
    const int lane = threadIdx.x & 31;
    const long long rows = (long long)B * T;
    const int blank = V - 1;
    for (long long row = (long long)blockIdx.x * 8 + (threadIdx.x >> 5); row < rows; row += (long long)gridDim.x * 8) {
        const int b = (int)(row / T), t = (int)(row - (long long)b * T);
        const CtcMeta mt = meta[b];
        if (mt.skip || t >= mt.Tb) continue;
        const double logp = logp_ws[b];
        if (logp <= CTC_ZERO_THRESH) continue;        // no valid path: grad stays = softmax
        const int S = 2 * mt.L + 1;
        const double* aw = alpha_ws + (size_t)row * NS;
        const double* bw = beta_ws + (size_t)row * NS;
        const int* lb = lab + (size_t)b * LABP;
        float* grow = grad + (size_t)row * V;
        float gblank = 0.f;
        for (int s = lane; s < S; s += 32) {
            const double ex = aw[s] + bw[s] - logp;
            const float gm = (ex < -80.0) ? 0.f : __expf((float)ex);
            if (s & 1) { if (gm != 0.f) atomicAdd(grow + lb[s >> 1], -gm); }
            else gblank += gm;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) gblank += __shfl_xor_sync(0xffffffffu, gblank, o);
        if (lane == 0 && gblank != 0.f) atomicAdd(grow + blank, -gblank);
    }
}

// --------------------------------------------------------------------------------------------
struct CtcPlan {
    int NW, SPT, TC;
    int LABP, LPP;
    size_t off_meta, off_lab, off_lpb, off_lpl, off_alpha, off_beta, off_logp, off_status, total;
    size_t smem;
};

static bool ctc_make_plan(int B, int T, int V, int Lmax, CtcPlan& p) {
    const int S = 2 * Lmax + 1;
    struct Cfg { int nw, spt; };
    // fewest states per thread first: the sweep is a serial chain of T steps whose length grows with the per-thread
    // instruction count (measured 4.4 cycles per issued instruction at one warp per scheduler), a CTA barrier costs less
    // (idle threads still execute the sweep: at L = 300 the 16 x 2 plan kept 41 % of its warps busy with padding states and the
    // kernel is XU-pipe bound there -- MUFU + f64<->f32 conversions -- so the warp count follows S closely)
    const Cfg cfgs[] = {{1, 2}, {2, 2}, {4, 2}, {6, 2}, {8, 2}, {10, 2}, {12, 2}, {16, 2}, {16, 4}, {32, 4}, {32, 8}};
    bool ok = false;
    for (const Cfg& c : cfgs)
        if (S <= c.nw * 32 * c.spt) { p.NW = c.nw; p.SPT = c.spt; ok = true; break; }
    if (!ok) return false;
    const int NT = p.NW * 32;
    p.LABP = (Lmax + 3) & ~3; if (p.LABP == 0) p.LABP = 4;
    p.LPP = NT * p.SPT / 2;                                  // one private slot group per thread
    // staging budget: ~64 KB per CTA
    const size_t per_frame = (size_t)NT * (p.SPT / 2) * 4;
    int tc = (int)((32 * 1024) / (2 * per_frame));
    if (tc < 1) tc = 1;
    if (tc > 32) tc = 32;
    p.TC = tc;
    p.smem = (size_t)2 * NT * sizeof(double2) + (size_t)2 * tc * NT * (p.SPT / 2) * 4 + (size_t)2 * ((tc + 3) & ~3) * 4 + 16;
    auto al = [](size_t x) { return (x + 255) & ~(size_t)255; };
    size_t o = 0;
    p.off_status = o; o = al(o + 16);
    p.off_meta = o; o = al(o + sizeof(CtcMeta) * (size_t)B);
    p.off_lab = o; o = al(o + sizeof(int) * (size_t)B * p.LABP);
    p.off_lpb = o; o = al(o + sizeof(float) * (size_t)B * T);
    p.off_lpl = o; o = al(o + sizeof(float) * (size_t)B * T * p.LPP);
    p.off_alpha = o; o = al(o + sizeof(double) * (size_t)B * T * NT * p.SPT);
    p.off_beta = o; o = al(o + sizeof(double) * (size_t)B * T * NT * p.SPT);
    p.off_logp = o; o = al(o + sizeof(double) * (size_t)B);
    p.total = o;
    return true;
}

template <int GROUP, int VEC, int NCH>
static void launch_softmax(const float* logits, float* grad, int B, int T, int V, const CtcMeta* meta, const int* lab,
                           int LABP, float* lpb, float* lpl, int LPP, cudaStream_t st)
{
    const long long rows = (long long)B * T;
    const int gpc = 256 / GROUP;
    long long want = (rows + gpc - 1) / gpc;
    int grid = (int)(want < num_sms() * 8 ? want : num_sms() * 8);
    if (grid < 1) grid = 1;
    ctc_softmax_kernel<GROUP, VEC, NCH><<<grid, 256, 0, st>>>(logits, grad, B, T, V, meta, lab, LABP, lpb, lpl, LPP);
}

template <int GROUP, int VEC>
static bool dispatch_nch(int nch, const float* logits, float* grad, int B, int T, int V, const CtcMeta* meta,
                         const int* lab, int LABP, float* lpb, float* lpl, int LPP, cudaStream_t st)
{
    if (nch <= 1) launch_softmax<GROUP, VEC, 1>(logits, grad, B, T, V, meta, lab, LABP, lpb, lpl, LPP, st);
    else if (nch <= 2) launch_softmax<GROUP, VEC, 2>(logits, grad, B, T, V, meta, lab, LABP, lpb, lpl, LPP, st);
    else if (nch <= 4) launch_softmax<GROUP, VEC, 4>(logits, grad, B, T, V, meta, lab, LABP, lpb, lpl, LPP, st);
    else if (nch <= 8) launch_softmax<GROUP, VEC, 8>(logits, grad, B, T, V, meta, lab, LABP, lpb, lpl, LPP, st);
    else return false;
    return true;
}

template <int VEC>
static void dispatch_softmax(const float* logits, float* grad, int B, int T, int V, const CtcMeta* meta, const int* lab,
                             int LABP, float* lpb, float* lpl, int LPP, cudaStream_t st)
{
    const int nvec = (V + VEC - 1) / VEC;
    const int nch_warp = (nvec + 31) / 32;
    if (nch_warp <= 8) { dispatch_nch<32, VEC>(nch_warp, logits, grad, B, T, V, meta, lab, LABP, lpb, lpl, LPP, st); return; }
    const int nch_cta = (nvec + 255) / 256;
    if (nch_cta <= 8) { dispatch_nch<256, VEC>(nch_cta, logits, grad, B, T, V, meta, lab, LABP, lpb, lpl, LPP, st); return; }
    const long long rows = (long long)B * T;
    int grid = (int)(rows < num_sms() * 8 ? rows : num_sms() * 8);
    ctc_softmax_bigrow_kernel<<<grid, 256, 0, st>>>(logits, grad, B, T, V, meta, lab, LABP, lpb, lpl, LPP);
}

template <int NW, int SPT>
static cudaError_t launch_ab(const CtcPlan& p, int B, int T, int V, const CtcMeta* meta, const int* lab, const float* lpb,
                             const float* lpl, double* alpha, double* beta, double* logp, float* grad, float* loss, cudaStream_t st)
{
    cudaError_t e = cudaFuncSetAttribute(ctc_alpha_beta_kernel<NW, SPT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.smem);
    if (e != cudaSuccess) return e;
    ctc_alpha_beta_kernel<NW, SPT><<<2 * B, NW * 32, p.smem, st>>>(meta, lab, p.LABP, lpb, lpl, p.LPP, alpha, beta, logp, loss, T, V, p.TC);
    const long long rows = (long long)B * T;
    long long want = (rows + 7) / 8;
    const int grid = (int)(want < num_sms() * 8 ? want : num_sms() * 8);
    ctc_gamma_kernel<<<grid, 256, 0, st>>>(meta, lab, p.LABP, alpha, beta, logp, grad, B, T, V, NW * 32 * SPT);
    return cudaGetLastError();
}

}  // namespace lcb

using namespace lcb;

extern "C" size_t lcb_ctc_workspace_bytes(int B, int T, int V, int Lmax)
{
    CtcPlan p;
    if (B <= 0 || T <= 0 || V < 2 || Lmax < 0) return 0;
    if (!ctc_make_plan(B, T, V, Lmax, p)) return 0;
    return p.total;
}

extern "C" int lcb_ctc_loss_grad_f32(const float* logits, const int64_t* labels, int Lmax, const int32_t* seq_len,
                                     int B, int T, int V, float* loss, float* grad, void* workspace,
                                     size_t workspace_bytes, void* stream)
{
    if (!logits || !seq_len || !loss || !grad || !workspace) return LCB_ERR_NULL_POINTER;
    if (B <= 0 || T <= 0 || V < 2 || Lmax < 0) return LCB_ERR_BAD_SHAPE;
    if (Lmax > 0 && !labels) return LCB_ERR_NULL_POINTER;
    CtcPlan p;
    if (!ctc_make_plan(B, T, V, Lmax, p)) return LCB_ERR_UNSUPPORTED;
    if (workspace_bytes < p.total) return LCB_ERR_WORKSPACE_TOO_SMALL;
    cudaStream_t st = (cudaStream_t)stream;
    unsigned char* ws = (unsigned char*)workspace;
    int* status = (int*)(ws + p.off_status);
    CtcMeta* meta = (CtcMeta*)(ws + p.off_meta);
    int* lab = (int*)(ws + p.off_lab);
    float* lpb = (float*)(ws + p.off_lpb);
    float* lpl = (float*)(ws + p.off_lpl);
    double* alpha = (double*)(ws + p.off_alpha);
    double* beta = (double*)(ws + p.off_beta);
    double* logp = (double*)(ws + p.off_logp);

    cudaMemsetAsync(status, 0, 16, st);
    g_launches += 4; ctc_prep_kernel<<<(B + 3) / 4, 128, 0, st>>>(labels, Lmax, seq_len, B, T, V, meta, lab, p.LABP, status);
    if ((V & 3) == 0 && ((uintptr_t)logits & 15) == 0 && ((uintptr_t)grad & 15) == 0)
        dispatch_softmax<4>(logits, grad, B, T, V, meta, lab, p.LABP, lpb, lpl, p.LPP, st);
    else if ((V & 1) == 0 && ((uintptr_t)logits & 7) == 0 && ((uintptr_t)grad & 7) == 0)
        dispatch_softmax<2>(logits, grad, B, T, V, meta, lab, p.LABP, lpb, lpl, p.LPP, st);
    else
        dispatch_softmax<1>(logits, grad, B, T, V, meta, lab, p.LABP, lpb, lpl, p.LPP, st);
    cudaError_t e = cudaSuccess;
    if (p.NW == 1 && p.SPT == 2) e = launch_ab<1, 2>(p, B, T, V, meta, lab, lpb, lpl, alpha, beta, logp, grad, loss, st);
    else if (p.NW == 2 && p.SPT == 2) e = launch_ab<2, 2>(p, B, T, V, meta, lab, lpb, lpl, alpha, beta, logp, grad, loss, st);
    else if (p.NW == 6 && p.SPT == 2) e = launch_ab<6, 2>(p, B, T, V, meta, lab, lpb, lpl, alpha, beta, logp, grad, loss, st);
    else if (p.NW == 10 && p.SPT == 2) e = launch_ab<10, 2>(p, B, T, V, meta, lab, lpb, lpl, alpha, beta, logp, grad, loss, st);
    else if (p.NW == 12 && p.SPT == 2) e = launch_ab<12, 2>(p, B, T, V, meta, lab, lpb, lpl, alpha, beta, logp, grad, loss, st);
    else if (p.NW == 4 && p.SPT == 2) e = launch_ab<4, 2>(p, B, T, V, meta, lab, lpb, lpl, alpha, beta, logp, grad, loss, st);
    else if (p.NW == 8 && p.SPT == 2) e = launch_ab<8, 2>(p, B, T, V, meta, lab, lpb, lpl, alpha, beta, logp, grad, loss, st);
    else if (p.NW == 16 && p.SPT == 2) e = launch_ab<16, 2>(p, B, T, V, meta, lab, lpb, lpl, alpha, beta, logp, grad, loss, st);
    else if (p.NW == 16 && p.SPT == 4) e = launch_ab<16, 4>(p, B, T, V, meta, lab, lpb, lpl, alpha, beta, logp, grad, loss, st);
    else if (p.NW == 32 && p.SPT == 4) e = launch_ab<32, 4>(p, B, T, V, meta, lab, lpb, lpl, alpha, beta, logp, grad, loss, st);
    else e = launch_ab<32, 8>(p, B, T, V, meta, lab, lpb, lpl, alpha, beta, logp, grad, loss, st);
    if (e != cudaSuccess) return LCB_ERR_CUDA;
    e = cudaGetLastError();
    return e == cudaSuccess ? LCB_OK : LCB_ERR_CUDA;
}

// status word written by the last lcb_ctc_loss_grad_f32 on this workspace (device pointer):
// 0 ok, LCB_ERR_INVALID_LABEL if some label was outside [0, V-1).  Synchronises the stream.
extern "C" int lcb_ctc_status(const void* workspace, void* stream)
{
    int h = 0;
    cudaError_t e = cudaMemcpyAsync(&h, workspace, sizeof(int), cudaMemcpyDeviceToHost, (cudaStream_t)stream);
    if (e != cudaSuccess) return LCB_ERR_CUDA;
    e = cudaStreamSynchronize((cudaStream_t)stream);
    if (e != cudaSuccess) return LCB_ERR_CUDA;
    return h;
}
