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
//                   grad[b,t,:] = softmax (0 past seq_len / skipped utts), plus lse[b,t] and the blank's
//                   log-prob lpb[b,t].  Reads logits once, writes grad once: the 8*T*B*V algorithmic bytes.
//   ctc_lattice   : ONE CTA per utterance runs the alpha sweep (forward in time) and the beta sweep (backward)
//                   concurrently on two warp groups, in a linear-domain mantissa/exponent representation, with
//                   the gradient fused in (see the kernel): loss[b] and grad[b,t,l'_s] -= gamma_t(s).
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
                   const CtcMeta* __restrict__ meta, float* __restrict__ lpb, float* __restrict__ lse_out)
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
        // what the lattice needs of this row besides the label logits it gathers itself: log-sum-exp and the blank's log-prob
        if (tig == 0) { lpb[row] = x[blank] - lse; lse_out[row] = lse; }
    }
}

// Fallback for rows longer than the register cache: CTA per row, three passes over global
// (passes 2 and 3 hit L2).
__global__ void __launch_bounds__(256)
ctc_softmax_bigrow_kernel(const float* __restrict__ logits, float* __restrict__ grad, int B, int T, int V,
                          const CtcMeta* __restrict__ meta, float* __restrict__ lpb, float* __restrict__ lse_out)
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
        if (threadIdx.x == 0) { lpb[row] = x[blank] - lse; lse_out[row] = lse; }
        __syncthreads();
    }
}

// --------------------------------------------------------------------------------------------
// lattice pass: alpha AND beta sweeps of one utterance in ONE CTA, gamma fused, linear domain
//
// Numbers.  Lattice values are kept as m * 2^e with m an fp32 in [1, 2) (or exactly 0) and e an int32: a linear-domain
// recursion (two adds and one multiply per state and frame, no ex2 / lg2 / fp64 on the serial chain) with the range of the log
// domain.  Adding aligns the mantissas to the largest exponent (integer ops + fmul), the product with the frame's emission
// 2^(lp * log2 e) adds exponents; every result is renormalised by reading its exponent field.  Relative error ~1e-7 per
// operation whatever the magnitude (plain fp32 log space, which is what TF runs, loses 1e-4..1e-3 at T ~ 3000).
//
// Mapping.  The CTA has an alpha group and a beta group of NW warps each; a thread owns SPT consecutive lattice states
// (blank, label, blank, label ...).  Inside a warp the neighbour states travel by shuffle; between warps through a small
// shared-memory ring with release / acquire counters, so the warps of a group run SKEWED (warp w one frame behind warp w-1)
// and no CTA-wide barrier sits on the T-step chain.  Emissions are gathered straight from the logits row (x[lab] - lse, the
// row's lse and blank log-prob come from the softmax pass) and prefetched one chunk of frames ahead.
//
// Meet in the middle.  alpha runs t = 0 .. Tb-1, beta runs t = Tb-1 .. 0, concurrently.  For a frame of the first half alpha
// arrives first and spills alpha_t; beta, arriving later, reads it and forms gamma_t = alpha_t beta_t / p on the fly.  For the
// second half the roles swap.  p = sum_s alpha_mid(s) beta_mid(s) is reduced once, where the sweeps cross (two CTA barriers
// per utterance).  Each sweep therefore spills only HALF of its rows (8 bytes per state), nobody re-reads the lattice in a
// separate pass, and grad[b,t,l'_s] -= gamma leaves the sweeps as red.global.add (blank states pre-summed per warp).
constexpr int ME_ZERO = -(1 << 28);          // exponent of an exact zero
constexpr int CTC_RING = 16;                 // depth of the inter-warp exchange ring (frames a warp may run ahead of its reader)
constexpr int CTC_CH = 4;                    // frames per prefetch chunk

__device__ __forceinline__ float ctc_ex2(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
// 2^d, flushed to 0 below 2^-126 (d <= 127)
__device__ __forceinline__ float pow2_int(int d) { d += 127; d = d < 0 ? 0 : d; return __int_as_float(d << 23); }
// v >= 0 with exponent offset E  ->  (m in [1,2) | 0, e)
__device__ __forceinline__ void me_norm(float v, int E, float& m, int& e) {
    const int bits = __float_as_int(v);
    const bool z = bits < 0x00800000;                       // zero or denormal
    m = z ? 0.f : __int_as_float((bits & 0x007fffff) | 0x3f800000);
    e = z ? ME_ZERO : E + ((bits >> 23) - 127);
}
// exp(lp) as (m, e)
__device__ __forceinline__ void me_exp(float lp, float& m, int& e) {
    const float x2 = fmaxf(lp * 1.4426950408889634f, -1.0e6f);
    const float xi = floorf(x2);
    m = ctc_ex2(x2 - xi);
    e = (int)xi;
}
__device__ __forceinline__ void me_add(float& m, int& e, float m2, int e2) {      // (m,e) += (m2,e2), result NOT normalised
    const int E = e > e2 ? e : e2;
    m = m * pow2_int(e - E) + m2 * pow2_int(e2 - E);
    e = E;
}
__device__ __forceinline__ int ld_acquire_s32(const int* p) {
    int v; asm volatile("ld.acquire.cta.shared.b32 %0, [%1];" : "=r"(v) : "r"(smem_u32(p)) : "memory"); return v;
}
__device__ __forceinline__ void st_release_s32(int* p, int v) {
    asm volatile("st.release.cta.shared.b32 [%0], %1;" ::"r"(smem_u32(p)), "r"(v) : "memory");
}
// bounded spin on a shared-memory counter (records a device error instead of hanging)
__device__ __forceinline__ int spin_until_gt(const int* p, int n) {
    int v = ld_acquire_s32(p);
    uint32_t spins = 0;
    while (v <= n) {
        if (++spins > (1u << 26)) { dev_set_error(DEV_ERR_MBAR_TIMEOUT); break; }
        v = ld_acquire_s32(p);
    }
    return v;
}

struct CtcXch {                 // one per warp boundary and group
    int prog;                   // steps published by the writer warp
    int cons;                   // steps consumed by the reader warp
    int pad[2];
    float m[2][CTC_RING];
    int e[2][CTC_RING];
};

template <int SPT, int MAXT>
__global__ void __launch_bounds__(MAXT)
ctc_lattice_kernel(const float* __restrict__ logits, const CtcMeta* __restrict__ meta, const int* __restrict__ lab, int LABP,
                   const float* __restrict__ lse, const float* __restrict__ lpb, int2* __restrict__ spill,
                   float* __restrict__ loss, float* __restrict__ grad, int T, int V, int NW)
{
    constexpr int HL = SPT / 2;
    constexpr int CH = CTC_CH;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    CtcXch* xch = reinterpret_cast<CtcXch*>(smem_raw);                       // [2 groups][NW]  (entry w: written by warp w of the group)
    float* red_m = reinterpret_cast<float*>(xch + 2 * NW);                   // [NW]
    int* red_e = reinterpret_cast<int*>(red_m + NW);                         // [NW]
    float* p_sh = reinterpret_cast<float*>(red_e + NW);                      // [4]: m_p, e_p (bits), 1/m_p, no-path flag (bits)

    const int b = blockIdx.x;
    const int tid = threadIdx.x, lane = tid & 31;
    const int NTG = NW * 32;                                                  // threads per group
    const int grp = tid >= NTG ? 1 : 0;                                       // 0: alpha sweep, 1: beta sweep
    const int gt = tid - grp * NTG;                                           // thread index inside the group
    const int w = gt >> 5;                                                    // warp inside the group
    const CtcMeta mt = meta[b];
    if (mt.skip) { if (tid == 0) loss[b] = 0.f; return; }                     // (the softmax pass zeroed the gradient rows)
    const int Tb = mt.Tb, L = mt.L, S = 2 * L + 1;
    const int mid = Tb >> 1;
    const int s0 = gt * SPT;
    const int NSP = NTG * SPT;                                                // padded states per frame in the spill
    const int* lb = lab + (size_t)b * LABP;
    const float* x_b = logits + (size_t)b * T * V;
    const float* lse_b = lse + (size_t)b * T;
    const float* lpb_b = lpb + (size_t)b * T;
    int2* sp_b = spill + (size_t)b * T * NSP + s0;
    float* g_b = grad + (size_t)b * T * V;
    const int blank = V - 1;

    for (int i = tid; i < 2 * NW; i += blockDim.x) { xch[i].prog = 0; xch[i].cons = 0; }
    // per-state constants
    int labv[HL];
    bool skipf[HL], skipb[HL], valid[SPT];
#pragma unroll
    for (int k = 0; k < HL; ++k) {
        const int j = s0 / 2 + k;
        const int l0 = (j < L) ? lb[j] : 0;
        labv[k] = l0;
        skipf[k] = (j < L) && (j >= 1) && (lb[j - 1] != l0);      // alpha: s-2 -> s allowed into label state 2j+1
        skipb[k] = (j + 1 < L) && (lb[j + 1] != l0);              // beta : s -> s+2 allowed out of label state 2j+1
    }
#pragma unroll
    for (int i = 0; i < SPT; ++i) valid[i] = (s0 + i < S);
    __syncthreads();

    float am[SPT]; int ae[SPT];          // this group's lattice row (alpha_t, or beta_t without the emission at t)
#pragma unroll
    for (int i = 0; i < SPT; ++i) { am[i] = 0.f; ae[i] = ME_ZERO; }
    float pinv = 0.f; int pe = 0; bool nopath = false;
    int cons_cache = 0;                  // writer side: last seen consumer count of the boundary this warp writes
    CtcXch* xw = &xch[grp * NW + w];     // ring this warp WRITES (read by its successor: alpha w+1, beta w-1)
    CtcXch* xr = grp == 0 ? (w > 0 ? &xch[w - 1] : nullptr) : (w + 1 < NW ? &xch[NW + w + 1] : nullptr);   // ring it READS
    const bool has_succ = grp == 0 ? (w + 1 < NW) : (w > 0);

    auto load_row = [&](int t, float (&om)[SPT], int (&oe)[SPT]) {           // the other sweep's spilled row of frame t
#pragma unroll
        for (int q = 0; q < SPT / 2; ++q) {
            const int4 v = __ldcg(reinterpret_cast<const int4*>(sp_b + (size_t)t * NSP + 2 * q));
            om[2 * q] = __int_as_float(v.x); oe[2 * q] = v.y; om[2 * q + 1] = __int_as_float(v.z); oe[2 * q + 1] = v.w;
        }
    };
    auto store_row = [&](int t) {
#pragma unroll
        for (int q = 0; q < SPT / 2; ++q)
            __stcg(reinterpret_cast<int4*>(sp_b + (size_t)t * NSP + 2 * q),
                   make_int4(__float_as_int(am[2 * q]), ae[2 * q], __float_as_int(am[2 * q + 1]), ae[2 * q + 1]));
    };
    // gamma_t(s) = alpha_t(s) beta_t(s) / p for this thread's states, subtracted from the gradient row of frame t
    auto gamma_row = [&](int t, const float (&om)[SPT], const int (&oe)[SPT]) {
        float gblank = 0.f;
        float* grow = g_b + (size_t)t * V;
#pragma unroll
        for (int i = 0; i < SPT; ++i) {
            const float g = (am[i] * om[i]) * pinv * pow2_int(ae[i] + oe[i] - pe);
            if (i & 1) { if (g > 9.0e-13f) atomicAdd(grow + labv[i >> 1], -g); }
            else gblank += g;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) gblank += __shfl_xor_sync(0xffffffffu, gblank, o);
        if (lane == 0 && gblank > 9.0e-13f) atomicAdd(grow + blank, -gblank);
    };
    auto read_p = [&]() { pe = __float_as_int(p_sh[1]); pinv = p_sh[2]; nopath = __float_as_int(p_sh[3]) != 0; };

    // register prefetch, CH frames ahead: slot c serves the frames with (step % CH) == c -- consumed, then refilled at once
    float xq[CH][HL];                    // logit of this thread's label(s)
    float oqm[CH][SPT]; int oqe[CH][SPT];   // the other sweep's spilled row (second phase; SPT == 2 only, else loaded at use)
    float lse_r = 0.f, lpb_r = 0.f;      // frame-shared values of a 32-frame block, one frame per lane

    if (grp == 0) {
        // =========================================== alpha sweep: t = 0 .. Tb-1 ===========================================
#pragma unroll
        for (int c = 0; c < CH; ++c)
#pragma unroll
            for (int k = 0; k < HL; ++k) xq[c][k] = (c < Tb) ? __ldg(x_b + (size_t)c * V + labv[k]) : 0.f;
        for (int t0 = 0; t0 < Tb; t0 += CH) {
#pragma unroll
            for (int c = 0; c < CH; ++c) {
                const int t = t0 + c;
                if (t < Tb) {
                    if ((t & 31) == 0) {
                        const int tt = t + lane;
                        lse_r = tt < Tb ? __ldg(lse_b + tt) : 0.f;
                        lpb_r = tt < Tb ? __ldg(lpb_b + tt) : 0.f;
                    }
                    const float lse_t = __shfl_sync(0xffffffffu, lse_r, t & 31);
                    const float lpb_t = __shfl_sync(0xffffffffu, lpb_r, t & 31);
                    if (t == mid) asm volatile("bar.sync 0;" ::: "memory");  // B1: every spill of phase 1 (both sweeps) is visible
                    float ybm; int ybe;
                    me_exp(lpb_t, ybm, ybe);
                    float ylm[HL]; int yle[HL];
#pragma unroll
                    for (int k = 0; k < HL; ++k) me_exp(xq[c][k] - lse_t, ylm[k], yle[k]);
                    if (t == 0) {
                        if (gt == 0) { am[0] = ybm; ae[0] = ybe; if (L > 0) { am[1] = ylm[0]; ae[1] = yle[0]; } }
                    } else {
                        // neighbour: the last (label) state of the previous thread, frame t-1
                        float nm = __shfl_up_sync(0xffffffffu, am[SPT - 1], 1);
                        int ne = __shfl_up_sync(0xffffffffu, ae[SPT - 1], 1);
                        if (lane == 0) {
                            nm = 0.f; ne = ME_ZERO;
                            if (xr) {
                                spin_until_gt(&xr->prog, t - 1);              // frame t-1 published
                                nm = xr->m[0][(t - 1) % CTC_RING]; ne = xr->e[0][(t - 1) % CTC_RING];
                                st_release_s32(&xr->cons, t);
                            }
                        }
                        float nmv[SPT]; int nev[SPT];
#pragma unroll
                        for (int i = 0; i < SPT; ++i) {
                            const float m1 = (i == 0) ? nm : am[i - 1];
                            const int e1 = (i == 0) ? ne : ae[i - 1];
                            if (i & 1) {
                                const int k = i >> 1;
                                float m2 = (i == 1) ? nm : am[i >= 2 ? i - 2 : 0];
                                int e2 = (i == 1) ? ne : ae[i >= 2 ? i - 2 : 0];
                                if (!skipf[k]) { m2 = 0.f; e2 = ME_ZERO; }
                                const int E = max(ae[i], max(e1, e2));
                                const float sm = (am[i] * pow2_int(ae[i] - E) + m1 * pow2_int(e1 - E) + m2 * pow2_int(e2 - E)) * ylm[k];
                                me_norm(sm, E + yle[k], nmv[i], nev[i]);
                            } else {
                                const int E = max(ae[i], e1);
                                const float sm = (am[i] * pow2_int(ae[i] - E) + m1 * pow2_int(e1 - E)) * ybm;
                                me_norm(sm, E + ybe, nmv[i], nev[i]);
                            }
                            if (!valid[i]) { nmv[i] = 0.f; nev[i] = ME_ZERO; }
                        }
#pragma unroll
                        for (int i = 0; i < SPT; ++i) { am[i] = nmv[i]; ae[i] = nev[i]; }
                    }
                    // publish this warp's last state of frame t to the next warp
                    if (has_succ && lane == 31) {
                        if (t - cons_cache >= CTC_RING) cons_cache = spin_until_gt(&xw->cons, t - CTC_RING);
                        xw->m[0][t % CTC_RING] = am[SPT - 1]; xw->e[0][t % CTC_RING] = ae[SPT - 1];
                        st_release_s32(&xw->prog, t + 1);
                    }
                    if (t < mid) {
                        store_row(t);
                    } else {
                        float om[SPT]; int oe[SPT];
                        if (SPT == 2 && t >= mid + CH) {
#pragma unroll
                            for (int i = 0; i < SPT; ++i) { om[i] = oqm[c][i]; oe[i] = oqe[c][i]; }
                        } else {
                            load_row(t, om, oe);
                        }
                        if (t == mid) {
                            // p = sum_s alpha_mid(s) beta_mid(s): thread sum -> warp shuffle reduce -> one thread combines the warps
                            float qm = 0.f; int qe = ME_ZERO;
#pragma unroll
                            for (int i = 0; i < SPT; ++i) me_add(qm, qe, am[i] * om[i], (am[i] == 0.f || om[i] == 0.f) ? ME_ZERO : ae[i] + oe[i]);
#pragma unroll
                            for (int o = 16; o > 0; o >>= 1) {
                                const float m2 = __shfl_xor_sync(0xffffffffu, qm, o);
                                const int e2 = __shfl_xor_sync(0xffffffffu, qe, o);
                                me_add(qm, qe, m2, e2);
                            }
                            if (lane == 0) { red_m[w] = qm; red_e[w] = qe; }
                            asm volatile("bar.sync 1, %0;" ::"r"(NTG) : "memory");          // the alpha group only
                            if (gt == 0) {
                                float tm = 0.f; int te = ME_ZERO;
                                for (int k = 0; k < NW; ++k) me_add(tm, te, red_m[k], red_e[k]);
                                float fm; int fe;
                                me_norm(tm, te, fm, fe);
                                const bool none = (fm == 0.f);
                                p_sh[0] = fm; p_sh[1] = __int_as_float(fe); p_sh[2] = none ? 0.f : 1.0f / fm; p_sh[3] = __int_as_float(none ? 1 : 0);
                                // loss = -ln p; no valid alignment: +inf and the gradient stays = softmax (TF behaviour)
                                loss[b] = none ? INFINITY : (float)(-(log2((double)fm) + (double)fe) * 0.6931471805599453);
                            }
                            asm volatile("bar.sync 0;" ::: "memory");                         // B2: p is published
                            read_p();
                        }
                        if (!nopath) gamma_row(t, om, oe);
                    }
                    // refill slot c for frame t + CH
                    const int tn = t + CH;
#pragma unroll
                    for (int k = 0; k < HL; ++k) xq[c][k] = (tn < Tb) ? __ldg(x_b + (size_t)tn * V + labv[k]) : 0.f;
                    if (SPT == 2 && t >= mid && tn < Tb) load_row(tn, oqm[c], oqe[c]);       // (B1 has passed: beta's rows are final)
                }
            }
        }
        return;
    }

    // =========================================== beta sweep: step n visits frame t = Tb-1-n ===========================================
    // am/ae hold beta_t WITHOUT the emission at t (TF's definition)
#pragma unroll
    for (int i = 0; i < SPT; ++i) if (s0 + i == S - 1 || s0 + i == S - 2) { am[i] = 1.f; ae[i] = 0; }     // beta_{Tb-1}
    bool synced = false;
#pragma unroll
    for (int c = 0; c < CH; ++c)
#pragma unroll
        for (int k = 0; k < HL; ++k) xq[c][k] = (Tb - 1 - c >= 0) ? __ldg(x_b + (size_t)(Tb - 1 - c) * V + labv[k]) : 0.f;
    for (int n0 = 0; n0 < Tb; n0 += CH) {
#pragma unroll
        for (int c = 0; c < CH; ++c) {
            const int n = n0 + c;
            const int t = Tb - 1 - n;
            if (t >= 0) {
                if ((t & 31) == 31 || n == 0) {                               // frame-shared values of the 32-frame block holding t
                    const int tt = (t & ~31) + lane;
                    lse_r = tt < Tb ? __ldg(lse_b + tt) : 0.f;
                    lpb_r = tt < Tb ? __ldg(lpb_b + tt) : 0.f;
                }
                const float lse_t = __shfl_sync(0xffffffffu, lse_r, t & 31);
                const float lpb_t = __shfl_sync(0xffffffffu, lpb_r, t & 31);
                if (t >= mid) {
                    store_row(t);
                } else {
                    if (!synced) {                                            // t == mid-1: the sweeps cross here
                        asm volatile("bar.sync 0;" ::: "memory");             // B1
                        asm volatile("bar.sync 0;" ::: "memory");             // B2
                        synced = true;
                        read_p();
                    }
                    float om[SPT]; int oe[SPT];
                    if (SPT == 2 && t <= mid - 1 - CH) {
#pragma unroll
                        for (int i = 0; i < SPT; ++i) { om[i] = oqm[c][i]; oe[i] = oqe[c][i]; }
                    } else {
                        load_row(t, om, oe);
                    }
                    if (!nopath) gamma_row(t, om, oe);
                }
                if (t > 0) {
                    // e(s) = beta_t(s) y_t(l'_s); beta_{t-1}(s) = e(s) + e(s+1) + [skip] e(s+2)
                    float ybm; int ybe;
                    me_exp(lpb_t, ybm, ybe);
                    float em[SPT]; int ee[SPT];
#pragma unroll
                    for (int i = 0; i < SPT; ++i) {
                        float ym = ybm; int ye = ybe;
                        if (i & 1) me_exp(xq[c][i >> 1] - lse_t, ym, ye);
                        em[i] = am[i] * ym;                                    // in [1,4) or 0: left unnormalised, the sums align it
                        ee[i] = (am[i] == 0.f) ? ME_ZERO : ae[i] + ye;
                    }
                    // neighbours: e(s0+SPT) (blank) and e(s0+SPT+1) (label) of the next thread, same frame
                    float n1m = __shfl_down_sync(0xffffffffu, em[0], 1), n2m = __shfl_down_sync(0xffffffffu, em[1], 1);
                    int n1e = __shfl_down_sync(0xffffffffu, ee[0], 1), n2e = __shfl_down_sync(0xffffffffu, ee[1], 1);
                    if (has_succ && lane == 0) {                               // publish to the warp below
                        if (n - cons_cache >= CTC_RING) cons_cache = spin_until_gt(&xw->cons, n - CTC_RING);
                        xw->m[0][n % CTC_RING] = em[0]; xw->e[0][n % CTC_RING] = ee[0];
                        xw->m[1][n % CTC_RING] = em[1]; xw->e[1][n % CTC_RING] = ee[1];
                        st_release_s32(&xw->prog, n + 1);
                    }
                    if (lane == 31) {
                        n1m = 0.f; n1e = ME_ZERO; n2m = 0.f; n2e = ME_ZERO;
                        if (xr) {
                            spin_until_gt(&xr->prog, n);                       // step n published by the warp above
                            n1m = xr->m[0][n % CTC_RING]; n1e = xr->e[0][n % CTC_RING];
                            n2m = xr->m[1][n % CTC_RING]; n2e = xr->e[1][n % CTC_RING];
                            st_release_s32(&xr->cons, n + 1);
                        }
                    }
#pragma unroll
                    for (int i = 0; i < SPT; ++i) {
                        const float m1 = (i + 1 < SPT) ? em[i + 1 < SPT ? i + 1 : 0] : n1m;
                        const int e1 = (i + 1 < SPT) ? ee[i + 1 < SPT ? i + 1 : 0] : n1e;
                        float r; int E;
                        if (i & 1) {
                            float m2 = (i + 2 < SPT) ? em[i + 2 < SPT ? i + 2 : 0] : n2m;   // (i odd: i + 2 is the next label state)
                            int e2 = (i + 2 < SPT) ? ee[i + 2 < SPT ? i + 2 : 0] : n2e;
                            if (!skipb[i >> 1]) { m2 = 0.f; e2 = ME_ZERO; }
                            E = max(ee[i], max(e1, e2));
                            r = em[i] * pow2_int(ee[i] - E) + m1 * pow2_int(e1 - E) + m2 * pow2_int(e2 - E);
                        } else {
                            E = max(ee[i], e1);
                            r = em[i] * pow2_int(ee[i] - E) + m1 * pow2_int(e1 - E);
                        }
                        me_norm(r, E, am[i], ae[i]);
                        if (!valid[i]) { am[i] = 0.f; ae[i] = ME_ZERO; }
                    }
                }
                // refill slot c for step n + CH
                const int tn = t - CH;
#pragma unroll
                for (int k = 0; k < HL; ++k) xq[c][k] = (tn >= 0) ? __ldg(x_b + (size_t)tn * V + labv[k]) : 0.f;
                if (SPT == 2 && t <= mid - 1 && tn >= 0) load_row(tn, oqm[c], oqe[c]);       // (B1 has passed: alpha's rows are final)
            }
        }
    }
    if (!synced) {                                                            // mid == 0 (Tb == 1): no frame of the first half
        asm volatile("bar.sync 0;" ::: "memory");
        asm volatile("bar.sync 0;" ::: "memory");
    }
}

// --------------------------------------------------------------------------------------------
struct CtcPlan {
    int NW, SPT;                 // warps per sweep (the CTA has 2*NW), lattice states per thread
    int LABP;
    size_t off_meta, off_lab, off_lpb, off_lse, off_spill, off_status, total;
    size_t smem;
};

static bool ctc_make_plan(int B, int T, int V, int Lmax, CtcPlan& p) {
    const int S = 2 * Lmax + 1;
    // fewest states per thread that fit one CTA (two sweeps x NW warps, <= 1024 threads): the frame loop of a warp is
    // issue-bound in its per-thread instruction count, and skewed warps cost no barrier
    p.SPT = 0;
    for (int spt = 2; spt <= 16; spt *= 2)
        if (S <= 512 * spt) { p.SPT = spt; break; }
    if (p.SPT == 0) return false;
    p.NW = (S + 32 * p.SPT - 1) / (32 * p.SPT);
    p.LABP = (Lmax + 3) & ~3; if (p.LABP == 0) p.LABP = 4;
    p.smem = (size_t)2 * p.NW * sizeof(CtcXch) + (size_t)p.NW * 8 + 16;
    auto al = [](size_t x) { return (x + 255) & ~(size_t)255; };
    size_t o = 0;
    p.off_status = o; o = al(o + 16);
    p.off_meta = o; o = al(o + sizeof(CtcMeta) * (size_t)B);
    p.off_lab = o; o = al(o + sizeof(int) * (size_t)B * p.LABP);
    p.off_lpb = o; o = al(o + sizeof(float) * (size_t)B * T);
    p.off_lse = o; o = al(o + sizeof(float) * (size_t)B * T);
    p.off_spill = o; o = al(o + sizeof(int2) * (size_t)B * T * p.NW * 32 * p.SPT);
    p.total = o;
    return true;
}

template <int GROUP, int VEC, int NCH>
static void launch_softmax(const float* logits, float* grad, int B, int T, int V, const CtcMeta* meta,
                           float* lpb, float* lse, cudaStream_t st)
{
    const long long rows = (long long)B * T;
    const int gpc = 256 / GROUP;
    long long want = (rows + gpc - 1) / gpc;
    int grid = (int)(want < num_sms() * 8 ? want : num_sms() * 8);
    if (grid < 1) grid = 1;
    ctc_softmax_kernel<GROUP, VEC, NCH><<<grid, 256, 0, st>>>(logits, grad, B, T, V, meta, lpb, lse);
}

template <int GROUP, int VEC>
static bool dispatch_nch(int nch, const float* logits, float* grad, int B, int T, int V, const CtcMeta* meta,
                         float* lpb, float* lse, cudaStream_t st)
{
    if (nch <= 1) launch_softmax<GROUP, VEC, 1>(logits, grad, B, T, V, meta, lpb, lse, st);
    else if (nch <= 2) launch_softmax<GROUP, VEC, 2>(logits, grad, B, T, V, meta, lpb, lse, st);
    else if (nch <= 4) launch_softmax<GROUP, VEC, 4>(logits, grad, B, T, V, meta, lpb, lse, st);
    else if (nch <= 8) launch_softmax<GROUP, VEC, 8>(logits, grad, B, T, V, meta, lpb, lse, st);
    else return false;
    return true;
}

template <int VEC>
static void dispatch_softmax(const float* logits, float* grad, int B, int T, int V, const CtcMeta* meta,
                             float* lpb, float* lse, cudaStream_t st)
{
    const int nvec = (V + VEC - 1) / VEC;
    const int nch_warp = (nvec + 31) / 32;
    if (nch_warp <= 8) { dispatch_nch<32, VEC>(nch_warp, logits, grad, B, T, V, meta, lpb, lse, st); return; }
    const int nch_cta = (nvec + 255) / 256;
    if (nch_cta <= 8) { dispatch_nch<256, VEC>(nch_cta, logits, grad, B, T, V, meta, lpb, lse, st); return; }
    const long long rows = (long long)B * T;
    int grid = (int)(rows < num_sms() * 8 ? rows : num_sms() * 8);
    ctc_softmax_bigrow_kernel<<<grid, 256, 0, st>>>(logits, grad, B, T, V, meta, lpb, lse);
}

template <int SPT, int MAXT>
static cudaError_t launch_lattice(const CtcPlan& p, int B, int T, int V, const float* logits, const CtcMeta* meta, const int* lab,
                                  const float* lse, const float* lpb, int2* spill, float* grad, float* loss, cudaStream_t st)
{
    ctc_lattice_kernel<SPT, MAXT><<<B, 2 * p.NW * 32, p.smem, st>>>(logits, meta, lab, p.LABP, lse, lpb, spill, loss, grad, T, V, p.NW);
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
    float* lse = (float*)(ws + p.off_lse);
    int2* spill = (int2*)(ws + p.off_spill);

    cudaMemsetAsync(status, 0, 16, st);
    g_launches += 3; ctc_prep_kernel<<<(B + 3) / 4, 128, 0, st>>>(labels, Lmax, seq_len, B, T, V, meta, lab, p.LABP, status);
    if ((V & 3) == 0 && ((uintptr_t)logits & 15) == 0 && ((uintptr_t)grad & 15) == 0)
        dispatch_softmax<4>(logits, grad, B, T, V, meta, lpb, lse, st);
    else if ((V & 1) == 0 && ((uintptr_t)logits & 7) == 0 && ((uintptr_t)grad & 7) == 0)
        dispatch_softmax<2>(logits, grad, B, T, V, meta, lpb, lse, st);
    else
        dispatch_softmax<1>(logits, grad, B, T, V, meta, lpb, lse, st);
    cudaError_t e = cudaSuccess;
    const int nt = 2 * p.NW * 32;
#define LCB_CTC_LAUNCH(SPT_, MAXT_) e = launch_lattice<SPT_, MAXT_>(p, B, T, V, logits, meta, lab, lse, lpb, spill, grad, loss, st)
    if (p.SPT == 2) { if (nt <= 256) LCB_CTC_LAUNCH(2, 256); else if (nt <= 512) LCB_CTC_LAUNCH(2, 512); else LCB_CTC_LAUNCH(2, 1024); }
    else if (p.SPT == 4) LCB_CTC_LAUNCH(4, 1024);
    else if (p.SPT == 8) LCB_CTC_LAUNCH(8, 1024);
    else LCB_CTC_LAUNCH(16, 1024);
#undef LCB_CTC_LAUNCH
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
