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
//                   grad[b,t,:] = softmax (0 past seq_len / skipped utts), plus per row the log-sum-exp and the
//                   blank's emission.  Reads logits once, writes grad once: the 8*T*B*V algorithmic bytes.
//   ctc_lattice   : ONE CTA per utterance runs the alpha sweep (forward in time) and the beta sweep (backward)
//                   concurrently on two warp groups, in a linear-domain mantissa/exponent representation, with
//                   the gradient fused in (see the kernel): loss[b] and grad[b,t,l'_s] -= gamma_t(s).
#include <type_traits>
#include "ptx.cuh"
#include "lstm_ctc_b200.h"

namespace lcb {

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
// wide-range positive reals: value = m * 2^e, m an fp32 in [1, 2), e an int32.  ZERO is "exponent ME_ZERO": every sum aligns its
// terms to the largest exponent, so a term 2^28 binades below contributes an exact 0 and no special case is needed.
constexpr int ME_ZERO = -(1 << 28);
__device__ __forceinline__ float ctc_ex2(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
// 2^d, flushed to 0 below 2^-126 (d <= 127)
__device__ __forceinline__ float pow2_int(int d) { d += 127; d = d < 0 ? 0 : d; return __int_as_float(d << 23); }
// v >= 2^-126 with exponent offset E  ->  (m in [1,2), e)
__device__ __forceinline__ void me_norm(float v, int E, float& m, int& e) {
    const int bits = __float_as_int(v);
    m = __int_as_float((bits & 0x007fffff) | 0x3f800000);
    e = E + ((bits >> 23) - 127);
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
                   const CtcMeta* __restrict__ meta, float4* __restrict__ frame)
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
        // what the lattice needs of this row besides the label logits it gathers itself: the log-sum-exp and the blank's
        // emission exp(x_blank - lse), already split into mantissa and exponent
        if (tig == 0) {
            float ym; int ye;
            me_exp(x[blank] - lse, ym, ye);
            frame[row] = make_float4(lse, ym, __int_as_float(ye), 0.f);
        }
    }
}

// Fallback for rows longer than the register cache: CTA per row, three passes over global
// (passes 2 and 3 hit L2).
__global__ void __launch_bounds__(256)
ctc_softmax_bigrow_kernel(const float* __restrict__ logits, float* __restrict__ grad, int B, int T, int V,
                          const CtcMeta* __restrict__ meta, float4* __restrict__ frame)
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
        if (threadIdx.x == 0) {
            float ym; int ye;
            me_exp(x[blank] - lse, ym, ye);
            frame[row] = make_float4(lse, ym, __int_as_float(ye), 0.f);
        }
        __syncthreads();
    }
}

// --------------------------------------------------------------------------------------------
// lattice pass: alpha AND beta sweeps of one utterance in ONE CTA, gamma fused, linear domain
//
// Numbers.  Lattice values are kept as m * 2^e with m an fp32 in [1, 2) and e an int32 (me_* helpers above): a linear-domain
// recursion (two adds and one multiply per state and frame, no ex2 / lg2 / fp64 on the serial chain) with the range of the log
// domain.  Relative error ~1e-7 per operation whatever the magnitude (plain fp32 log space, which is what TF runs, loses
// 1e-4..1e-3 at T ~ 3000).
//
// Mapping.  The CTA has an alpha group and a beta group of NW warps each; a thread owns SPT consecutive lattice states
// (blank, label, blank, label ...).  Inside a warp the neighbour states travel by shuffle; between warps through a small
// shared-memory ring whose slots carry value and sequence number in ONE 16-byte word, so the warps of a group run SKEWED
// (warp w about one frame behind warp w-1) and neither a CTA-wide barrier nor a memory fence sits on the T-step chain.
// Everything a frame needs from global memory -- the thread's label logit(s), the row's (lse, blank emission) from the softmax
// pass, and in the second phase the other sweep's spilled row -- is staged PF-1 frames ahead with cp.async into per-thread
// shared-memory slots.  (Register prefetch does not work here: the compiler folds all outstanding loads onto the same few
// scoreboards, so waiting for the oldest waits for the newest -- measured 5000 cycles per frame instead of ~500.)
//
// Meet in the middle.  alpha runs t = 0 .. Tb-1, beta runs t = Tb-1 .. 0, concurrently.  For a frame of the first half alpha
// arrives first and spills alpha_t; beta, arriving later, reads it and forms gamma_t = alpha_t beta_t / p on the fly.  For the
// second half the roles swap.  p = sum_s alpha_mid(s) beta_mid(s) is reduced once, where the sweeps cross (two CTA barriers
// per utterance).  Each sweep therefore spills only HALF of its rows (8 bytes per state), nobody re-reads the lattice in a
// separate pass, and grad[b,t,l'_s] -= gamma leaves the sweeps as red.global.add (blank states pre-summed per warp).
constexpr int CTC_RING = 16;                 // depth of the inter-warp exchange ring (frames a warp may run ahead of its reader)
constexpr int CTC_PF = 8;                    // cp.async stages: frames staged ahead per thread (power of two)

// ---- inter-warp exchange without fences ----
// A ring slot is ONE 16-byte shared-memory word {mantissa, exponent, sequence number, -}: written by a single st.shared.v4 and
// polled by a single ld.volatile.shared.v4, so value and "ready" flag arrive together and no release / acquire fence is needed
// (a membar.cta per frame would also wait for the sweep's global spill stores and gradient reds in flight).  The reader's
// progress counter is a plain volatile word: its load of the slot precedes its store of the counter in program order on the
// same in-order shared-memory pipe.
__device__ __forceinline__ void ring_put(uint32_t addr, float m, int e, int seq) {
    asm volatile("st.volatile.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(__float_as_int(m)), "r"(e), "r"(seq), "r"(0) : "memory");
}
__device__ __forceinline__ void ring_get(uint32_t addr, int seq, float& m, int& e) {
    int a, b, c, d;
    uint32_t spins = 0;
    do {
        asm volatile("ld.volatile.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(a), "=r"(b), "=r"(c), "=r"(d) : "r"(addr) : "memory");
        if (c != seq && ++spins > (1u << 26)) { dev_set_error(DEV_ERR_MBAR_TIMEOUT); break; }      // never hang the device
    } while (c != seq);
    m = __int_as_float(a); e = b;
}
__device__ __forceinline__ int ld_volatile_s32(uint32_t addr) {
    int v; asm volatile("ld.volatile.shared.b32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory"); return v;
}
__device__ __forceinline__ void st_volatile_s32(uint32_t addr, int v) {
    asm volatile("st.volatile.shared.b32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}
// writer back-pressure: wait until the reader has consumed more than n steps
__device__ __forceinline__ int wait_consumed_gt(uint32_t addr, int n) {
    int v = ld_volatile_s32(addr);
    uint32_t spins = 0;
    while (v <= n) {
        if (++spins > (1u << 26)) { dev_set_error(DEV_ERR_MBAR_TIMEOUT); break; }
        v = ld_volatile_s32(addr);
    }
    return v;
}
__device__ __forceinline__ void cp_async4_s(uint32_t dst, const void* src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async16_s(uint32_t dst, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
// staged data: plain (non-volatile) shared loads -- ordered after cp.async.wait_group by that statement's memory clobber, otherwise
// free to be scheduled among the arithmetic
__device__ __forceinline__ float lds_f32(uint32_t addr) { float v; asm("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr) : "memory"); return v; }
__device__ __forceinline__ int4 lds_v4(uint32_t addr) {
    int4 v; asm("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr) : "memory"); return v;
}

struct __align__(16) CtcXch {   // one per warp and group: the ring this warp WRITES for its successor
    int4 slot[CTC_RING][2];     // [step % RING][value]: {m bits, e, seq = step + 1, 0}
    int cons;                   // steps consumed by the reader warp
    int pad[3];
};

template <int SPT> struct CtcLat {
    static constexpr int HL = SPT / 2;
    static constexpr int PF = SPT == 16 ? CTC_PF / 2 : CTC_PF;   // cp.async stages (the shared-memory budget halves it at SPT = 16)
    static constexpr bool ROWPF = (SPT == 2);                // the other sweep's spilled rows are staged too (16 B per thread and frame)
    __host__ __device__ static size_t off_stage(int NW) { return ((size_t)2 * NW * sizeof(CtcXch) + (size_t)NW * 8 + 16 + 15) & ~(size_t)15; }
    __host__ __device__ static size_t off_frame(int NW) { return off_stage(NW) + (size_t)PF * 2 * NW * 32 * HL * 4; }      // [PF][2 NW] float4
    __host__ __device__ static size_t off_rows(int NW) { return off_frame(NW) + (size_t)PF * 2 * NW * 16; }                // [PF][NT] int4
    static size_t smem_bytes(int NW) { return off_rows(NW) + (ROWPF ? (size_t)PF * 2 * NW * 32 * 16 : 0); }
};

//
// SPLIT (few utterances: 2 B <= SMs).  The call is then bound by the per-frame chain of one warp, and the alpha and beta warps of an utterance
// compete for the same four schedulers.  The two sweeps run as the two CTAs of a cluster instead (rank 0: alpha, rank 1: beta), on two SMs:
// the spilled rows travel through L2 as before (st.cg / ld.cg / cp.async.cg), the two CTA barriers become barrier.cluster (release / acquire
// at cluster scope orders the spills), and p reaches the beta CTA as three remote shared-memory stores before the second one.
template <int SPT, int MAXT, bool SPLIT>
__global__ void __launch_bounds__(MAXT)
ctc_lattice_kernel(const float* __restrict__ logits, const CtcMeta* __restrict__ meta, const int* __restrict__ lab, int LABP,
                   const float4* __restrict__ frame, int2* __restrict__ spill,
                   float* __restrict__ loss, float* __restrict__ grad, int T, int V, int NW)
{
    using C = CtcLat<SPT>;
    constexpr int HL = C::HL, PF = C::PF;
    constexpr bool ROWPF = C::ROWPF;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int NT = (SPLIT ? 1 : 2) * NW * 32;                                 // threads of the CTA
    const int NTG = NW * 32;                                                  // threads per group
    CtcXch* xch = reinterpret_cast<CtcXch*>(smem_raw);                       // [2 groups][NW]  (entry w: written by warp w of the group)
    float* red_m = reinterpret_cast<float*>(xch + 2 * NW);                   // [NW]
    int* red_e = reinterpret_cast<int*>(red_m + NW);                         // [NW]
    float* p_sh = reinterpret_cast<float*>(red_e + NW);                      // [4]: -, e_p (bits), 1/m_p, no-path flag (bits)

    const int b = SPLIT ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
    const int tid = threadIdx.x, lane = tid & 31;
    const int wg = __shfl_sync(0xffffffffu, tid >> 5, 0);                     // warp of the CTA (provably warp-uniform)
    const int grp = SPLIT ? (int)(blockIdx.x & 1) : (wg >= NW ? 1 : 0);       // 0: alpha sweep, 1: beta sweep (SPLIT: the CTA's rank in its pair)
    const int w = SPLIT ? wg : wg - grp * NW;                                 // warp inside the group
    const int gt = SPLIT ? tid : tid - grp * NTG;                             // thread index inside the group
    auto sync_both = [&]() {                                                  // both sweeps of the utterance
        if constexpr (SPLIT) cluster_sync_all(); else asm volatile("bar.sync 0;" ::: "memory");
    };
    const CtcMeta mt = meta[b];
    if (mt.skip) { if (tid == 0) loss[b] = 0.f; return; }                     // (the softmax pass zeroed the gradient rows)
    const int Tb = __shfl_sync(0xffffffffu, mt.Tb, 0), L = __shfl_sync(0xffffffffu, mt.L, 0), S = 2 * L + 1;
    const int mid = Tb >> 1;
    const int s0 = gt * SPT;
    const int NSP = NTG * SPT;                                                // padded states per frame in the spill
    const int* lb = lab + (size_t)b * LABP;
    const float* x_b = logits + (size_t)b * T * V;
    const float4* fr_b = frame + (size_t)b * T;
    int2* sp_b = spill + (size_t)b * T * NSP + s0;
    float* g_b = grad + (size_t)b * T * V;
    const int blank = V - 1;
    // staging slots (shared-window byte addresses; slot k of a ring sits k * stride further)
    const uint32_t xs = smem_u32(smem_raw + C::off_stage(NW)) + (uint32_t)tid * HL * 4, XS = (uint32_t)NT * HL * 4;
    const uint32_t fs = smem_u32(smem_raw + C::off_frame(NW)) + (uint32_t)wg * 16, FS = (uint32_t)2 * NW * 16;
    const uint32_t rs = smem_u32(smem_raw + C::off_rows(NW)) + (uint32_t)tid * 16, RS = (uint32_t)NT * 16;

    for (int i = tid; i < 2 * NW * CTC_RING * 2; i += blockDim.x)
        reinterpret_cast<int4*>(xch)[(i / (CTC_RING * 2)) * (sizeof(CtcXch) / 16) + (i % (CTC_RING * 2))] = make_int4(0, 0, 0, 0);
    for (int i = tid; i < 2 * NW; i += blockDim.x) xch[i].cons = 0;
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
    for (int i = 0; i < SPT; ++i) { am[i] = 1.f; ae[i] = ME_ZERO; }
    float pinv = 0.f; int pe = 0; bool nopath = false;
    int cons_cache = 0;                  // writer side: last seen consumer count of the ring this warp writes
    const uint32_t xw = smem_u32(&xch[grp * NW + w]);                         // ring this warp WRITES (successor: alpha w+1, beta w-1)
    const bool has_pred = grp == 0 ? (w > 0) : (w + 1 < NW);
    const uint32_t xr = smem_u32(&xch[grp * NW + (grp == 0 ? (w > 0 ? w - 1 : 0) : (w + 1 < NW ? w + 1 : w))]);  // ring it READS
    const bool has_succ = grp == 0 ? (w + 1 < NW) : (w > 0);
    constexpr uint32_t CONS_OFF = CTC_RING * 2 * 16;
    const int dir = grp == 0 ? 1 : -1;                                        // frames per step
    const int t_first = grp == 0 ? 0 : Tb - 1;

    // running global pointers (advanced by one frame per step in the sweep's direction)
    const float* xsrc[HL];                                                    // label logits of the frame being STAGED
#pragma unroll
    for (int k = 0; k < HL; ++k) xsrc[k] = x_b + (size_t)t_first * V + labv[k];
    const float4* fsrc = fr_b + t_first;                                      // (lse, blank emission) of the frame being staged
    const int2* rsrc = sp_b + (size_t)t_first * NSP;                          // other sweep's row of the frame being staged
    int2* srow = sp_b + (size_t)t_first * NSP;                                // spill row of the CURRENT frame
    float* grow = g_b + (size_t)t_first * V;                                  // gradient row of the current frame
    const ptrdiff_t dV = (ptrdiff_t)dir * V, dN = (ptrdiff_t)dir * NSP;
    int staged = 0;                                                           // steps staged so far

    auto stage = [&](bool rows_ok) {                                          // stage the next frame of the sweep into slot staged % PF
        const bool in = staged < Tb;
        const uint32_t sl = (uint32_t)(staged & (PF - 1));
        if (in) {
#pragma unroll
            for (int k = 0; k < HL; ++k) cp_async4_s(xs + sl * XS + 4 * k, xsrc[k]);
            if (lane == 0) cp_async16_s(fs + sl * FS, fsrc);
            if (ROWPF && rows_ok) cp_async16_s(rs + sl * RS, rsrc);
        }
#pragma unroll
        for (int k = 0; k < HL; ++k) xsrc[k] += dV;
        fsrc += dir; rsrc += dN;
        ++staged;
        cp_async_commit();
    };
    auto load_row = [&](float (&om)[SPT], int (&oe)[SPT]) {                   // the other sweep's spilled row of the current frame, from L2
#pragma unroll
        for (int k = 0; k < SPT / 2; ++k) {
            const int4 v = __ldcg(reinterpret_cast<const int4*>(srow + 2 * k));
            om[2 * k] = __int_as_float(v.x); oe[2 * k] = v.y; om[2 * k + 1] = __int_as_float(v.z); oe[2 * k + 1] = v.w;
        }
    };
    auto staged_row = [&](uint32_t sl, float (&om)[SPT], int (&oe)[SPT]) {    // ... or from its cp.async slot (SPT == 2)
        const int4 v = lds_v4(rs + sl * RS);
        om[0] = __int_as_float(v.x); oe[0] = v.y; om[SPT - 1] = __int_as_float(v.z); oe[SPT - 1] = v.w;
    };
    auto store_row = [&]() {
#pragma unroll
        for (int k = 0; k < SPT / 2; ++k)
            __stcg(reinterpret_cast<int4*>(srow + 2 * k),
                   make_int4(__float_as_int(am[2 * k]), ae[2 * k], __float_as_int(am[2 * k + 1]), ae[2 * k + 1]));
    };
    // gamma_t(s) = alpha_t(s) beta_t(s) / p for this thread's states, subtracted from the gradient row of the current frame
    auto gamma_row = [&](const float (&om)[SPT], const int (&oe)[SPT]) {
        float gblank = 0.f;
#pragma unroll
        for (int i = 0; i < SPT; ++i) {
            const float g = (am[i] * om[i]) * pinv * pow2_int(ae[i] + oe[i] - pe);
            if (i & 1) { if (g > 9.0e-13f) atomicAdd(grow + labv[i >> 1], -g); }
            else gblank += g;
        }
        // warp sum of the blank states' gamma in ONE instruction: gamma <= 1 and their sum over the whole lattice row is <= 1, so a
        // 2^-30 fixed-point integer add-reduce (redux.sync) is exact to 1e-9 -- instead of five dependent shuffle + add rounds
        const int gi = __reduce_add_sync(0xffffffffu, __float2int_rn(gblank * 1073741824.f));
        if (lane == 0 && gi > 0) atomicAdd(grow + blank, -(float)gi * 9.31322574615478515625e-10f);
    };
    auto read_p = [&]() { pe = __float_as_int(p_sh[1]); pinv = p_sh[2]; nopath = __float_as_int(p_sh[3]) != 0; };
    // emissions of the NEXT step, converted one iteration ahead so that their shared-memory reads and ex2 sit off the chain
    float ybm = 1.f; int ybe = 0;
    float ylm[HL]; int yle[HL];
#pragma unroll
    for (int k = 0; k < HL; ++k) { ylm[k] = 1.f; yle[k] = 0; }
    auto emissions = [&](int n) {                                             // of step n, from its staging slot
        const uint32_t sl = (uint32_t)(n & (PF - 1));
        const int4 f4 = lds_v4(fs + sl * FS);                                 // {lse, blank emission mantissa, exponent, -}
        ybm = __int_as_float(f4.y); ybe = f4.z;
#pragma unroll
        for (int k = 0; k < HL; ++k) me_exp(lds_f32(xs + sl * XS + 4 * k) - __int_as_float(f4.x), ylm[k], yle[k]);
    };
    // Inter-warp exchange, warp-uniform: EVERY lane polls the predecessor warp's slot (same address: a broadcast read), so the
    // wait loop never diverges; the boundary lane keeps the value.  Stores are predicated on the boundary lane.
    auto ring_wait_free = [&](int n) {                                        // before publishing step n
        if (has_succ && n - cons_cache >= CTC_RING) cons_cache = wait_consumed_gt(xw + CONS_OFF, n - CTC_RING);
    };

    for (int i = 0; i < PF - 1; ++i) stage(false);
    cp_async_wait<PF - 2>();
    __syncwarp();
    emissions(0);

    if (grp == 0) {
        // =========================================== alpha sweep: t = 0 .. Tb-1 ===========================================
        // one frame: recursion (the emissions of frame t are in registers), publish, stage frame t+PF-1, convert frame t+1's emissions
        auto frame = [&](int t, auto first, auto second) {
            constexpr bool FIRST = decltype(first)::value;                    // t == 0
            constexpr bool SECOND = decltype(second)::value;                  // t >= mid: gamma instead of spill
            if (!FIRST) {
                // neighbour: the last (label) state of the previous thread, frame t-1
                float nm = __shfl_up_sync(0xffffffffu, am[SPT - 1], 1);
                int ne = __shfl_up_sync(0xffffffffu, ae[SPT - 1], 1);
                if (has_pred) {                                               // (warp-uniform)
                    float rm; int re;
                    ring_get(xr + (uint32_t)(((t - 1) & (CTC_RING - 1)) * 32), t, rm, re);   // frame t-1 (seq = t), all lanes
                    if (lane == 0) { nm = rm; ne = re; st_volatile_s32(xr + CONS_OFF, t); }
                } else if (lane == 0) {
                    ne = ME_ZERO;
                }
                float nmv[SPT]; int nev[SPT];
#pragma unroll
                for (int i = 0; i < SPT; ++i) {
                    const float m1 = (i == 0) ? nm : am[i >= 1 ? i - 1 : 0];
                    const int e1 = (i == 0) ? ne : ae[i >= 1 ? i - 1 : 0];
                    if (i & 1) {
                        const int k = i >> 1;
                        const float m2 = (i == 1) ? nm : am[i >= 2 ? i - 2 : 0];
                        int e2 = (i == 1) ? ne : ae[i >= 2 ? i - 2 : 0];
                        e2 = skipf[k] ? e2 : ME_ZERO;
                        const int E = max(ae[i], max(e1, e2));
                        const float sm = (am[i] * pow2_int(ae[i] - E) + m1 * pow2_int(e1 - E) + m2 * pow2_int(e2 - E)) * ylm[k];
                        me_norm(sm, E + yle[k], nmv[i], nev[i]);
                    } else {
                        const int E = max(ae[i], e1);
                        const float sm = (am[i] * pow2_int(ae[i] - E) + m1 * pow2_int(e1 - E)) * ybm;
                        me_norm(sm, E + ybe, nmv[i], nev[i]);
                    }
                    nev[i] = valid[i] ? nev[i] : ME_ZERO;
                }
#pragma unroll
                for (int i = 0; i < SPT; ++i) { am[i] = nmv[i]; ae[i] = nev[i]; }
            } else if (gt == 0) {
                am[0] = ybm; ae[0] = ybe;
                if (L > 0) { am[1] = ylm[0]; ae[1] = yle[0]; }
            }
            // publish this warp's last state of frame t to the next warp
            ring_wait_free(t);
            if (has_succ && lane == 31) ring_put(xw + (uint32_t)((t & (CTC_RING - 1)) * 32), am[SPT - 1], ae[SPT - 1], t + 1);
            stage(SECOND);                                                    // frame t + PF - 1 (rows only after B1: beta's are final then)
            cp_async_wait<PF - 2>();                                          // frame t + 1 has landed (this thread's copies)
            __syncwarp();                                                     // ... and lane 0's frame scalars for the whole warp
            if (t + 1 < Tb) emissions(t + 1);
            if (!SECOND) {
                store_row();
            } else {
                float om[SPT]; int oe[SPT];
                if (ROWPF && t - (PF - 1) >= mid) staged_row((uint32_t)(t & (PF - 1)), om, oe); else load_row(om, oe);
                if (t == mid) {
                    // p = sum_s alpha_mid(s) beta_mid(s): thread sum -> warp shuffle reduce -> one thread combines the warps
                    float qm = 1.f; int qe = ME_ZERO;
#pragma unroll
                    for (int i = 0; i < SPT; ++i) me_add(qm, qe, am[i] * om[i], ae[i] + oe[i]);
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) {
                        const float m2 = __shfl_xor_sync(0xffffffffu, qm, o);
                        const int e2 = __shfl_xor_sync(0xffffffffu, qe, o);
                        me_add(qm, qe, m2, e2);
                    }
                    if (lane == 0) { red_m[w] = qm; red_e[w] = qe; }
                    asm volatile("bar.sync 1, %0;" ::"r"(NTG) : "memory");          // the alpha group only
                    if (gt == 0) {
                        float tm = 1.f; int te = ME_ZERO;
                        for (int k = 0; k < NW; ++k) me_add(tm, te, red_m[k], red_e[k]);
                        float fm; int fe;
                        me_norm(tm, te, fm, fe);
                        const bool none = fe < (ME_ZERO >> 1);                       // no valid alignment
                        p_sh[1] = __int_as_float(fe); p_sh[2] = none ? 0.f : 1.0f / fm; p_sh[3] = __int_as_float(none ? 1 : 0);
                        if constexpr (SPLIT) {                                        // the beta CTA's copy (visible after B2's release / acquire)
                            const uint32_t ra = mapa_shared(smem_u32(p_sh), 1);
                            st_cluster_b32(ra + 4, (uint32_t)fe); st_cluster_b32(ra + 8, __float_as_uint(none ? 0.f : 1.0f / fm));
                            st_cluster_b32(ra + 12, none ? 1u : 0u);
                        }
                        // loss = -ln p; no valid alignment: +inf and the gradient stays = softmax (TF behaviour)
                        loss[b] = none ? INFINITY : (float)(-(log2((double)fm) + (double)fe) * 0.6931471805599453);
                    }
                    sync_both();                                                      // B2: p is published
                    read_p();
                }
                if (!nopath) gamma_row(om, oe);
            }
            srow += NSP; grow += V;
        };
        using TT = std::true_type; using FF = std::false_type;
        int t = 0;
        if (mid > 0) {
            frame(0, TT{}, FF{});
            for (t = 1; t < mid; ++t) frame(t, FF{}, FF{});
        }
        sync_both();                                                           // B1: every spill of phase 1 (both sweeps) is visible
        if (mid == 0) { frame(0, TT{}, TT{}); t = 1; }
        for (; t < Tb; ++t) frame(t, FF{}, TT{});
        cp_async_wait<0>();
        return;
    }

    // =========================================== beta sweep: step n visits frame t = Tb-1-n ===========================================
    // am/ae hold beta_t WITHOUT the emission at t (TF's definition)
#pragma unroll
    for (int i = 0; i < SPT; ++i) if (s0 + i == S - 1 || s0 + i == S - 2) ae[i] = 0;                      // beta_{Tb-1} = 1
    auto bframe = [&](int n, auto second) {
        constexpr bool SECOND = decltype(second)::value;                      // t < mid: gamma instead of spill
        const int t = Tb - 1 - n;
        if (!SECOND) {
            store_row();
        } else {
            float om[SPT]; int oe[SPT];
            if (ROWPF && t + (PF - 1) <= mid - 1) staged_row((uint32_t)(n & (PF - 1)), om, oe); else load_row(om, oe);
            if (!nopath) gamma_row(om, oe);
        }
        if (t > 0) {
            // e(s) = beta_t(s) y_t(l'_s); beta_{t-1}(s) = e(s) + e(s+1) + [skip] e(s+2)
            float em[SPT]; int ee[SPT];
#pragma unroll
            for (int i = 0; i < SPT; ++i) {
                em[i] = am[i] * ((i & 1) ? ylm[i >> 1] : ybm);                 // in [1,4): left unnormalised, the sums align it
                ee[i] = ae[i] + ((i & 1) ? yle[i >> 1] : ybe);
            }
            // neighbours: e(s0+SPT) (blank) and e(s0+SPT+1) (label) of the next thread, same frame
            float n1m = __shfl_down_sync(0xffffffffu, em[0], 1), n2m = __shfl_down_sync(0xffffffffu, em[1], 1);
            int n1e = __shfl_down_sync(0xffffffffu, ee[0], 1), n2e = __shfl_down_sync(0xffffffffu, ee[1], 1);
            ring_wait_free(n);
            if (has_succ && lane == 0) {                                       // publish to the warp below
                ring_put(xw + (uint32_t)((n & (CTC_RING - 1)) * 32), em[0], ee[0], n + 1);
                ring_put(xw + (uint32_t)((n & (CTC_RING - 1)) * 32 + 16), em[1], ee[1], n + 1);
            }
            if (has_pred) {                                                    // step n of the warp above (seq = n + 1), all lanes poll
                float r1m, r2m; int r1e, r2e;
                ring_get(xr + (uint32_t)((n & (CTC_RING - 1)) * 32), n + 1, r1m, r1e);
                ring_get(xr + (uint32_t)((n & (CTC_RING - 1)) * 32 + 16), n + 1, r2m, r2e);
                if (lane == 31) { n1m = r1m; n1e = r1e; n2m = r2m; n2e = r2e; st_volatile_s32(xr + CONS_OFF, n + 1); }
            } else if (lane == 31) {
                n1e = ME_ZERO; n2e = ME_ZERO;
            }
#pragma unroll
            for (int i = 0; i < SPT; ++i) {
                const float m1 = (i + 1 < SPT) ? em[i + 1 < SPT ? i + 1 : 0] : n1m;
                const int e1 = (i + 1 < SPT) ? ee[i + 1 < SPT ? i + 1 : 0] : n1e;
                float r; int E;
                if (i & 1) {
                    const float m2 = (i + 2 < SPT) ? em[i + 2 < SPT ? i + 2 : 0] : n2m;   // (i odd: i + 2 is the next label state)
                    int e2 = (i + 2 < SPT) ? ee[i + 2 < SPT ? i + 2 : 0] : n2e;
                    e2 = skipb[i >> 1] ? e2 : ME_ZERO;
                    E = max(ee[i], max(e1, e2));
                    r = em[i] * pow2_int(ee[i] - E) + m1 * pow2_int(e1 - E) + m2 * pow2_int(e2 - E);
                } else {
                    E = max(ee[i], e1);
                    r = em[i] * pow2_int(ee[i] - E) + m1 * pow2_int(e1 - E);
                }
                me_norm(r, E, am[i], ae[i]);
                ae[i] = valid[i] ? ae[i] : ME_ZERO;
            }
        }
        stage(SECOND);                                                        // frame t - (PF-1) (rows only after B1: alpha's are final then)
        cp_async_wait<PF - 2>();
        __syncwarp();
        if (n + 1 < Tb) emissions(n + 1);
        srow -= NSP; grow -= V;
    };
    {
        using TT = std::true_type; using FF = std::false_type;
        int n = 0;
        for (; n < Tb - mid; ++n) bframe(n, FF{});                            // frames Tb-1 .. mid: spill
        sync_both();                                                          // B1
        sync_both();                                                          // B2
        read_p();
        for (; n < Tb; ++n) bframe(n, TT{});                                  // frames mid-1 .. 0: gamma
        cp_async_wait<0>();
    }
}

// --------------------------------------------------------------------------------------------
struct CtcPlan {
    int NW, SPT;                 // warps per sweep (the CTA has 2*NW), lattice states per thread
    int LABP;
    size_t off_meta, off_lab, off_frame, off_spill, off_status, total;
    size_t smem;
};

static int g_ctc_min_spt = 2;          // debug (lcb_debug_ctc_min_spt): smallest states-per-thread the plan may choose
static bool ctc_make_plan(int B, int T, int V, int Lmax, CtcPlan& p) {
    const int S = 2 * Lmax + 1;
    // fewest states per thread that fit one CTA (two sweeps x NW warps, <= 1024 threads): the frame loop of a warp is
    // issue-bound in its per-thread instruction count, and skewed warps cost no barrier
    p.SPT = 0;
    for (int spt = g_ctc_min_spt; spt <= 16; spt *= 2)
        if (S <= 512 * spt) { p.SPT = spt; break; }
    if (p.SPT == 0) return false;
    p.NW = (S + 32 * p.SPT - 1) / (32 * p.SPT);
    p.LABP = (Lmax + 3) & ~3; if (p.LABP == 0) p.LABP = 4;
    p.smem = p.SPT == 2 ? CtcLat<2>::smem_bytes(p.NW) : p.SPT == 4 ? CtcLat<4>::smem_bytes(p.NW) : p.SPT == 8 ? CtcLat<8>::smem_bytes(p.NW) : CtcLat<16>::smem_bytes(p.NW);
    auto al = [](size_t x) { return (x + 255) & ~(size_t)255; };
    size_t o = 0;
    p.off_status = o; o = al(o + 16);
    p.off_meta = o; o = al(o + sizeof(CtcMeta) * (size_t)B);
    p.off_lab = o; o = al(o + sizeof(int) * (size_t)B * p.LABP);
    p.off_frame = o; o = al(o + sizeof(float4) * (size_t)B * T);
    p.off_spill = o; o = al(o + sizeof(int2) * (size_t)B * T * p.NW * 32 * p.SPT);
    p.total = o;
    return true;
}

constexpr int g_ctc_softmax_ctas_per_sm = 8;      // resident CTAs per SM of the (persistent, grid-stride) softmax pass

template <int GROUP, int VEC, int NCH>
static void launch_softmax(const float* logits, float* grad, int B, int T, int V, const CtcMeta* meta,
                           float4* frame, cudaStream_t st)
{
    const long long rows = (long long)B * T;
    const int gpc = 256 / GROUP;
    long long want = (rows + gpc - 1) / gpc;
    int grid = (int)(want < num_sms() * g_ctc_softmax_ctas_per_sm ? want : num_sms() * g_ctc_softmax_ctas_per_sm);
    if (grid < 1) grid = 1;
    ctc_softmax_kernel<GROUP, VEC, NCH><<<grid, 256, 0, st>>>(logits, grad, B, T, V, meta, frame);
}

template <int GROUP, int VEC>
static bool dispatch_nch(int nch, const float* logits, float* grad, int B, int T, int V, const CtcMeta* meta,
                         float4* frame, cudaStream_t st)
{
    if (nch <= 1) launch_softmax<GROUP, VEC, 1>(logits, grad, B, T, V, meta, frame, st);
    else if (nch <= 2) launch_softmax<GROUP, VEC, 2>(logits, grad, B, T, V, meta, frame, st);
    else if (nch <= 4) launch_softmax<GROUP, VEC, 4>(logits, grad, B, T, V, meta, frame, st);
    else if (nch <= 8) launch_softmax<GROUP, VEC, 8>(logits, grad, B, T, V, meta, frame, st);
    else return false;
    return true;
}

template <int VEC>
static void dispatch_softmax(const float* logits, float* grad, int B, int T, int V, const CtcMeta* meta,
                             float4* frame, cudaStream_t st)
{
    const int nvec = (V + VEC - 1) / VEC;
    const int nch_warp = (nvec + 31) / 32;
    if (nch_warp <= 8) { dispatch_nch<32, VEC>(nch_warp, logits, grad, B, T, V, meta, frame, st); return; }
    const int nch_cta = (nvec + 255) / 256;
    if (nch_cta <= 8) { dispatch_nch<256, VEC>(nch_cta, logits, grad, B, T, V, meta, frame, st); return; }
    const long long rows = (long long)B * T;
    int grid = (int)(rows < num_sms() * g_ctc_softmax_ctas_per_sm ? rows : num_sms() * g_ctc_softmax_ctas_per_sm);
    ctc_softmax_bigrow_kernel<<<grid, 256, 0, st>>>(logits, grad, B, T, V, meta, frame);
}

template <int SPT, int MAXT, bool SPLIT>
static cudaError_t launch_lattice(const CtcPlan& p, int B, int T, int V, const float* logits, const CtcMeta* meta, const int* lab,
                                  const float4* frame, int2* spill, float* grad, float* loss, cudaStream_t st)
{
    auto kern = ctc_lattice_kernel<SPT, MAXT, SPLIT>;
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.smem) != cudaSuccess)
        return cudaErrorInvalidValue;
    if constexpr (!SPLIT) {
        kern<<<B, 2 * p.NW * 32, p.smem, st>>>(logits, meta, lab, p.LABP, frame, spill, loss, grad, T, V, p.NW);
        return cudaGetLastError();
    } else {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3((unsigned)(2 * B), 1, 1);
        cfg.blockDim = dim3((unsigned)(p.NW * 32), 1, 1);
        cfg.dynamicSmemBytes = p.smem;
        cfg.stream = st;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        return cudaLaunchKernelEx(&cfg, kern, logits, meta, lab, p.LABP, frame, (int2*)spill, loss, grad, T, V, p.NW);
    }
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

// lattice_layout: -1 chosen by batch size (the two sweeps of an utterance on two SMs when 2 B <= SMs, else one CTA per utterance),
// 0 one CTA per utterance, 1 two.  Same results up to the order of the gradient's atomic adds; 0 / 1 are for tests and A/B timing.
extern "C" int lcb_ctc_loss_grad_f32_layout(const float* logits, const int64_t* labels, int Lmax, const int32_t* seq_len,
                                            int B, int T, int V, float* loss, float* grad, void* workspace,
                                            size_t workspace_bytes, int lattice_layout, void* stream)
{
    if (lattice_layout < -1 || lattice_layout > 1) return LCB_ERR_BAD_SHAPE;
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
    float4* frame = (float4*)(ws + p.off_frame);
    int2* spill = (int2*)(ws + p.off_spill);

    cudaMemsetAsync(status, 0, 16, st);
    g_launches += 1; ctc_prep_kernel<<<(B + 3) / 4, 128, 0, st>>>(labels, Lmax, seq_len, B, T, V, meta, lab, p.LABP, status);
    const int NSP = p.NW * 32 * p.SPT;
    // the two passes over utterances [b0, b0 + nb)
    auto softmax_pass = [&](int b0, int nb, cudaStream_t s) {
        const float* x = logits + (size_t)b0 * T * V;
        float* g = grad + (size_t)b0 * T * V;
        g_launches += 1;
        if ((V & 3) == 0 && ((uintptr_t)x & 15) == 0 && ((uintptr_t)g & 15) == 0)
            dispatch_softmax<4>(x, g, nb, T, V, meta + b0, frame + (size_t)b0 * T, s);
        else if ((V & 1) == 0 && ((uintptr_t)x & 7) == 0 && ((uintptr_t)g & 7) == 0)
            dispatch_softmax<2>(x, g, nb, T, V, meta + b0, frame + (size_t)b0 * T, s);
        else
            dispatch_softmax<1>(x, g, nb, T, V, meta + b0, frame + (size_t)b0 * T, s);
    };
    auto lattice_pass = [&](int b0, int nb, cudaStream_t s) -> cudaError_t {
        const float* x = logits + (size_t)b0 * T * V;
        float* g = grad + (size_t)b0 * T * V;
        const int nt = 2 * p.NW * 32;
        cudaError_t e = cudaSuccess;
        g_launches += 1;
#define LCB_CTC_LAUNCH(SPT_, MAXT_, SPLIT_) e = launch_lattice<SPT_, MAXT_, SPLIT_>(p, nb, T, V, x, meta + b0, lab + (size_t)b0 * p.LABP, \
                                                                     frame + (size_t)b0 * T, spill + (size_t)b0 * T * NSP, g, loss + b0, s)
        // few utterances: the two sweeps of an utterance on two SMs (see the kernel)
        // (8 and 16 states per thread -- more than 1023 labels -- always run split: one sweep per CTA is <= 512 threads and may
        // use 128 registers; compiled for 1024 threads = 64 registers those instantiations spill and were observed to fault with
        // five or more warps per sweep, tools/gpu_ctc_spt_repro.py)
        const bool split = p.SPT >= 8 || (lattice_layout >= 0 ? lattice_layout != 0 : 2 * nb <= num_sms());
        if (split) {
            const int ns = p.NW * 32;
            if (p.SPT == 2) { if (ns <= 128) LCB_CTC_LAUNCH(2, 128, true); else if (ns <= 256) LCB_CTC_LAUNCH(2, 256, true); else LCB_CTC_LAUNCH(2, 512, true); }
            else if (p.SPT == 4) LCB_CTC_LAUNCH(4, 512, true);
            else if (p.SPT == 8) LCB_CTC_LAUNCH(8, 512, true);
            else LCB_CTC_LAUNCH(16, 512, true);
        }
        else if (p.SPT == 2) { if (nt <= 256) LCB_CTC_LAUNCH(2, 256, false); else if (nt <= 512) LCB_CTC_LAUNCH(2, 512, false); else LCB_CTC_LAUNCH(2, 1024, false); }
        else LCB_CTC_LAUNCH(4, 1024, false);
#undef LCB_CTC_LAUNCH
        return e;
    };
    cudaError_t e = cudaSuccess;
    // (Tried and dropped, profiles/r02_ctc_notes.txt: running the lattice of one half of the batch on a helper stream beside the
    // softmax pass of the other half at large V -- the persistent softmax CTAs and the 640-thread / 41 K-register lattice CTAs do
    // not co-reside, and a softmax grid thin enough to leave room no longer saturates HBM: 2.54 -> 2.55 / 2.74 ms at V = 5000.)
    softmax_pass(0, B, st);
    e = lattice_pass(0, B, st);
    if (e != cudaSuccess) return LCB_ERR_CUDA;
    e = cudaGetLastError();
    return e == cudaSuccess ? LCB_OK : LCB_ERR_CUDA;
}

extern "C" int lcb_ctc_loss_grad_f32(const float* logits, const int64_t* labels, int Lmax, const int32_t* seq_len,
                                     int B, int T, int V, float* loss, float* grad, void* workspace,
                                     size_t workspace_bytes, void* stream)
{
    return lcb_ctc_loss_grad_f32_layout(logits, labels, Lmax, seq_len, B, T, V, loss, grad, workspace, workspace_bytes, -1, stream);
}

// debug / measurements: smallest number of lattice states per thread the plan may choose (2, 4, 8, 16; default 2).  Affects
// lcb_ctc_workspace_bytes as well: set it before sizing the workspace.
extern "C" int lcb_debug_ctc_min_spt(int spt)
{
    if (spt != 2 && spt != 4 && spt != 8 && spt != 16) return LCB_ERR_BAD_SHAPE;
    lcb::g_ctc_min_spt = spt;
    return LCB_OK;
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
