// lstm_rec.cu -- persistent thread-block-cluster kernels for the LSTM recurrence (K2) and its BPTT (K2b).
//
// Replaces tf.nn.dynamic_rnn over DropoutWrapper(LSTMCell(num_units=H, num_proj=P, use_peepholes,
// forget_bias=5.0)) for BOTH directions of one BiLSTM layer (nnet/bilstm.py:127-137,148-158,171-188),
// the three tf.reverse_sequence copies per layer (bilstm.py:112,190,203) and the autodiff'd while-loop
// that tf.gradients builds for it (nnet/graph.py:190-191).
//
// Formulation.  TF's cell is z = [x_t, h_{t-1}] * kernel + bias, h = m * W_proj.  The x-part is hoisted
// into one GEMM over all frames (G = X * W_x + bias, gemm.cu); the recurrent part is folded,
//     h_{t-1} * W_h = m_{t-1} * (W_proj * W_h) = m_{t-1} * W',        W' : [H, 4H]
// so that one time step costs ONE on-chip exchange (of m_t) instead of two; h = M * W_proj is again a
// bulk GEMM after the loop.  Gate columns are packed gate-major inside groups of 8 units:
// col(unit, gate) = (unit/8)*32 + gate*8 + unit%8, gates (i,j,f,o)  (see packed_col()).
//
// Mapping.  One cluster of NC = Hp/32 CTAs per (direction, group of 16 utterances); both directions and all
// utterance groups run concurrently in one launch.  CTA c owns 32 units = 128 packed gate rows of W'^T and
// keeps them RESIDENT IN TENSOR MEMORY for the whole sequence (128 lanes x Hp/2 columns, fp16): the per-step
// product uses the TS form of tcgen05.mma (A from TMEM), so the weights are never re-read through the
// shared-memory port -- measured, the SS form was bound at ~40 cycles per 128x16x16 MMA by exactly that.
// Per step: two issuer warps (K halves, own accumulators; converged, one elected lane, warp-uniform operands --
// see the v2 section) issue D[128 gate rows, 16 utts] = W'^T_slice * m_{t-1}^T; 8 compute warps read their accumulator fragment with tcgen05.ld.16x256b (the packing above puts
// the four gates of a unit in ONE thread), add the prefetched G tile, apply gates / peepholes / cell update
// / length mask in registers (the fp32 cell state never leaves registers), hand the CTA's m_t slice (fp16) to an
// exchange warp that bulk-stores it to an L2-resident scratch and multicasts it into the operand buffer of every CTA
// of the cluster (completion counted on the receivers' mbarriers -- no barrier.cluster and no release fence in the
// loop), then write m_t and the saved activations to HBM off the critical path.  A loader warp keeps an 8-deep ring of G tiles in flight (one bulk copy per lane).
// The backward direction is the same scan in descending absolute time under the mask t < len[b]
// (state stays at its zero initial value until t = len[b]-1), so no reversed copies are ever made.
//
// BPTT runs the mirrored scan with W' (bf16) resident in TMEM as [units, own gate rows]: dz_t is formed in
// registers, staged locally as the MMA B operand, partial dm_{t-1} tiles are reduce-scattered to their owner
// CTAs with st.async (double-buffered reduce buffer, mbarrier tx-counted).
#include <cstring>
#include <cstdlib>
#include <cuda_fp16.h>
#include "ptx.cuh"
#include "tma_host.h"
#include "lstm_ctc_b200.h"

namespace lcb {

// debug probes: when set (lcb_debug_rec_profile), CTA 0 of cluster 0 records clock64() at phase boundaries of each
// step: [step][0..7] issuer 0, [step][8..15] compute warp 0 lane 0
__device__ long long* g_rec_prof = nullptr;
__device__ int g_rec_prof_steps = 0;
#define REC_PROBE(slot) do { if (prof && s < prof_steps) prof[(size_t)s * 16 + (slot)] = clock64(); } while (0)

#ifndef LCB_REC_NIW
#define LCB_REC_NIW 2             // issuer warps of the v2 forward / v3 BPTT kernels (measured: 1 issuer = 657-cycle passes, 2 = 526)
#endif
constexpr int REC_BG = 16;        // utterances per cluster
constexpr int REC_GCOLS = 128;    // fp32 row of a staged G tile (dense TMA box: the 4-way bank conflict of its 8 reads per thread and step is off the chain)
constexpr int REC_SG = 8;         // G prefetch ring depth
constexpr int REC_MAXGRP = 32;    // 16-utterance groups whose longest utterance travels in the kernel parameters (B <= 512; beyond: no early exit)
constexpr int REC_TMEM_ACC = 256; // first accumulator column (weights occupy [0, Hp/2) forward, [0, 64*MB) backward)

struct RecFwdParams {
    const float* G;               // [T*B][8Hp] fp32  x_t * W_x + bias, packed gate columns
    const __half* Wt;             // [8Hp][Hp] fp16   (W_proj * W_h)^T, rows = packed gate columns of both directions
    const float* peep;            // [2][3][Hp]  (w_f, w_i, w_o) or nullptr
    const int* lens;              // [B]
    __half* Mout;                 // [T*B][2Hp]   m_t, fp16 (0 where t >= len)
    __nv_bfloat16* Mbf;           // nullable: the same rows as bf16, the operand of the weight-gradient GEMMs (which cannot mix fp16 x bf16)
    uint2* gates;                 // [T*B][2Hp]   saved (i, tanh j, f, o) as 4 x fp16   (nullptr: inference)
    float* cst;                   // [T*B][2Hp]   saved cell state c_t                   (nullptr: inference)
    float* cfin;                  // [B][2][Hp] final cell state   (nullable)
    float* mfin;                  // [B][2][Hp] final m (pre-projection output) (nullable)
    int T, B, Hp, NC;
    int ndir;                     // 2: clusters alternate fd / bd direction; 1: uni-directional layer, direction 0 only
    float forget_bias;
    unsigned char* xch;           // v2: L2 exchange scratch [clusters][2][NC][slice] (nullptr: DSMEM copies)
    int s_begin, s_end;           // v2: scan steps [s_begin, s_end) of this launch (t = s forward, T-1-s backward direction);
                                  // a launch with s_begin > 0 resumes from the saved cst / Mout of step s_begin - 1
    int n_paired;                 // clusters (per direction) that hold TWO 16-utterance groups: cluster c < n_paired holds groups
                                  // 2c and 2c+1, cluster c >= n_paired holds group n_paired + c alone (fwd_layout())
    int gmax[REC_MAXGRP];         // longest utterance of each 16-utterance group (T when the caller gave no host-side lengths)
    const int* ready;             // nullable: device word = number of leading scan steps whose G rows exist (see the loader warp)
    int* progress;                // nullable: [clusters * NSG * NC] words; word (cluster, sub-group, CTA) = number of leading scan steps
                                  //     whose Mout rows this CTA has written for the sub-group's utterances (lcb_lstm_rec_fwd_range_pg)
};

struct RecBwdParams {
    const float* dM;              // [T*B][2Hp]  d loss / d m_t from the output projection
    const uint2* gates;           // [T*B][2Hp]
    const float* cst;             // [T*B][2Hp]
    const __nv_bfloat16* W;       // [2*Hp][4Hp] bf16  W' per direction: rows = units, cols = packed gate columns
    const float* peep;            // [2][3][Hp] or nullptr
    const int* lens;              // [B]
    __nv_bfloat16* dG;            // [T*B][8Hp]  d loss / d z_t  (0 where t >= len)
    float* dbias;                 // [2][4Hp]   += (packed column order)
    float* dpeep;                 // [2][3][Hp] += (nullable)
    int T, B, Hp, NC;
    int ndir;                     // 2 | 1, as in RecFwdParams
    unsigned char* xch;           // v3: L2 exchange scratch [clusters][2][NC][dz slice]
    int s_begin, s_end;           // v3: scan steps [s_begin, s_end) of this launch (t = T-1-s forward, s backward direction)
    float* carry;                 // v3: [B][2][Hp][2] (recurrent dm, carried dc) handed from one launch to the next (nullable
                                  //     when the launch covers [0, T))
    int* progress;                // v3, nullable: [clusters * NSG * 16] words; word (cluster, sub-group, CTA) = number of leading scan
                                  //     steps whose dG rows this CTA has written (lcb_lstm_rec_bwd_range_pg)
};

__device__ __forceinline__ unsigned char* align_1024(unsigned char* p) {
    return reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(p) + 1023) & ~(uintptr_t)1023);
}

// packed gate column of (unit, gate): gate-major inside each group of 8 units, so that one warp's
// tcgen05.ld.16x256b delivers all four gates of a unit to the same thread
__host__ __device__ __forceinline__ int packed_col(int unit, int gate) { return ((unit >> 3) << 5) + (gate << 3) + (unit & 7); }

// =================================================================================================
// the kernels: groups of BG utterances step in lockstep through a cluster (MMA N = BG)
//
// Per step:  operand m_{t-1} lands (multicast)  ->  32 MMAs 128 x BG x 16, A = weights from TMEM (TS form)  ->  commit->wake,
// tcgen05.ld, gate math (MUFU-bound: 5 ex2 + 2 rcp per cell with shared reciprocals)  ->  the CTA's slice of m_t is staged,
// bulk-stored to an L2-resident scratch and multicast into every CTA of the cluster.
// THE MMA ISSUE.  The first versions issued the MMAs from inside an `if (lane == 0)` region; the compiler then wraps every
// tcgen05.mma in an ELECT / 4 x R2UR / branch "waterfall" (its operands live in per-thread registers) and one thread issues one
// MMA per ~65 cycles -- the 32-MMA pass took ~1150 cycles whatever N was and was mistaken for a TMEM-bandwidth limit.  With the
// issuer warp converged, a canonical (shuffled) warp index and TMEM base, and the issuing lane selected by a predicate inside
// the asm (umma_f16_ts_elect), the operands stay in uniform registers: a pass is ~420 cycles at N = 16 and ~600 at N = 32
// (profiles/r01_recprobe_fwd_uniform_issue.txt).  Two issuer warps (own accumulators) still beat one (526 vs 657 cycles).
// =================================================================================================
__device__ __forceinline__ float ex2_approx(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float rcp_approx(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float clamp25(float x) { return fminf(fmaxf(x, -25.f), 25.f); }
// (sigmoid(a), sigmoid(b), tanh(c)) with ONE reciprocal: 1/(1+e^-a) = (1+e^-b)(1+e^-2c) / prod, ...  Arguments are clamped
// to +-25 (sigmoid(25) = 1 - 1.4e-11) so the product of the three denominators stays below 2^110.
__device__ __forceinline__ void sig_sig_tanh(float a, float b, float c, float& sa, float& sb, float& tc) {
    const float ea = 1.f + ex2_approx(-1.4426950408889634f * clamp25(a));
    const float eb = 1.f + ex2_approx(-1.4426950408889634f * clamp25(b));
    const float ec = 1.f + ex2_approx(-2.8853900817779268f * clamp25(c));
    const float ab = ea * eb;
    const float r = rcp_approx(ab * ec);
    sa = r * eb * ec;
    sb = r * ea * ec;
    tc = 2.f * (r * ab) - 1.f;
}
__device__ __forceinline__ float tanh_ex2(float c) {
    return 2.f * rcp_approx(1.f + ex2_approx(-2.8853900817779268f * clamp25(c))) - 1.f;
}
__device__ __forceinline__ void sig_tanh(float a, float c, float& sa, float& tc) {
    const float ea = 1.f + ex2_approx(-1.4426950408889634f * clamp25(a));
    const float ec = 1.f + ex2_approx(-2.8853900817779268f * clamp25(c));
    const float r = rcp_approx(ea * ec);
    sa = r * ec;
    tc = 2.f * (r * ea) - 1.f;
}

// NSG = 2 (BG = 16 only): TWO independent 16-utterance groups per cluster share the resident weights.  Each has its own
// warps, operand buffers, barriers, accumulators and exchange scratch and steps on its own; they meet only in the tensor
// pipe, where their weight passes take strict turns.  Measured (H = 512, B = 64, profiles/r01_recprobe_fwd_uniform_issue.txt):
// 2400-2460 cycles per step for both groups against ~2900 for one lockstep group of 32 (1 KB slices, half the operand
// ingress, half the gate math per thread-step; one group's exchange hides behind the other's pass and gate math); a
// 16-utterance group alone in a cluster steps in ~1970 cycles, but 8 such clusters are not co-resident on a B200 (7).
template <int BG, int NSG = 1> struct RecFwd2Cfg {
    static_assert(NSG == 1 || BG == 16, "two sub-groups only for 16-utterance groups");
    static constexpr int NUB = BG / 8;                     // utterance blocks of 8 (core-matrix rows of the MMA B operand)
    static constexpr int NCW = 4 * NUB;                    // compute warps: (utterance block, TMEM lane quarter)
    static constexpr int NIW = LCB_REC_NIW;                // MMA issuer warps (each with its own accumulator)
    static constexpr int WPS = NCW + NIW + 2;              // warps per sub-group: + loader warp + exchange warp
    static constexpr int THREADS = 32 * WPS * NSG;
    static constexpr int SLICE = 512 * NUB;                // bytes of one CTA's m_t slice: 32 units x BG utterances, fp16
    static constexpr int BARS = 1 + 2 * REC_SG + 2 + 2 + 1; // mma g[SG] gfree[SG] op[2] slice[2] acc
    __host__ __device__ static size_t op_bytes(int KB) { return (size_t)KB * 8 * NUB * 128; }    // one operand buffer (all K blocks)
    __host__ __device__ static size_t g_bytes() { return (size_t)REC_SG * BG * REC_GCOLS * 4; }
    __host__ __device__ static size_t sg_bytes(int KB) { return (2 * op_bytes(KB) + g_bytes() + 2 * SLICE + BARS * 8 + 1023) & ~(size_t)1023; }
    static size_t smem_bytes(int KB) { return 1024 + NSG * sg_bytes(KB) + 64; }
};

// GH: the hoisted pre-activations G arrive as fp16 (half the store + reload of the projections' output; the forget bias and the
// recurrent product are still added in fp32 in the accumulator) instead of fp32.
template <int BG, int NSG, bool GH = false>
__global__ void __launch_bounds__(RecFwd2Cfg<BG, NSG>::THREADS, 1)
lstm_rec_fwd2_kernel(const RecFwdParams p, const __grid_constant__ CUtensorMap tmG)
{
    using Cfg = RecFwd2Cfg<BG, NSG>;
    constexpr int SG = REC_SG, NUB = Cfg::NUB, NCW = Cfg::NCW, NIW = Cfg::NIW, SLICE = Cfg::SLICE;
    extern __shared__ unsigned char smem_raw[];
    unsigned char* smem_all = align_1024(smem_raw);
    const int Hp = p.Hp, KB = Hp >> 6, NC = p.NC, T = p.T, B = p.B;
    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;   // provably warp-uniform
    // warp order: the compute warps of all sub-groups first (warp % 4 stays the TMEM lane quarter), then per sub-group its
    // issuer(s), loader and exchange warp
    const int sg = warp < NSG * NCW ? warp / NCW : (warp - NSG * NCW) / (NIW + 2);
    const int wl = warp < NSG * NCW ? warp - sg * NCW : NCW + (warp - NSG * NCW) - sg * (NIW + 2);     // warp inside the sub-group
    int role = wl < NCW ? 0 : (wl < NCW + NIW ? 1 : wl - NCW - NIW + 2);   // 0 compute | 1 MMA issuers | 2 loader | 3 exchange
    const int rw = role == 0 ? wl : wl - NCW;
    const uint32_t OPB = (uint32_t)Cfg::op_bytes(KB);
    unsigned char* smem = smem_all + (size_t)sg * Cfg::sg_bytes(KB);

    // operand m_{t-1}: no-swizzle K-major core matrices (8 utterances x 8 units = 128 B), index (unit/8)*NUB + utt/8
    unsigned char* Bsm = smem;                                                   // [2][KB*8][NUB][128 B]
    float* Gsm = reinterpret_cast<float*>(Bsm + 2 * (size_t)OPB);               // [SG][BG][132]
    unsigned char* Msm = reinterpret_cast<unsigned char*>(Gsm + (size_t)SG * BG * REC_GCOLS);   // [2 step parities][SLICE]
    uint64_t* bars = reinterpret_cast<uint64_t*>(Msm + 2 * SLICE);
    uint64_t* mbar_mma = bars;                         //      this step's accumulator is complete
    uint64_t* mbar_g = bars + 1;                       // [SG] G tile landed (bulk-copy tx)
    uint64_t* mbar_gfree = bars + 1 + SG;              // [SG] compute warps are done with the G stage
    uint64_t* mbar_op = bars + 1 + 2 * SG;             // [2]  operand buffer: the slices of all NC CTAs have landed
    uint64_t* mbar_slice = bars + 3 + 2 * SG;          // [2]  this CTA's m_t slice is staged for the exchange warp
    uint64_t* mbar_acc = bars + 5 + 2 * SG;            //      the accumulator holds the next step's x-part (G tile): MMAs may accumulate
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem_all + (size_t)NSG * Cfg::sg_bytes(KB));
    // NSG = 2: the sub-groups' weight passes take strict turns on the tensor pipe.  mbar_turn[g] completes a phase each time the
    // OTHER sub-group has issued a pass (NIW issuer arrivals): waiting on an mbarrier suspends the issuer warp instead of
    // spinning on a shared-memory flag (the step is power-capped; idle warps should not burn issue slots)
    uint64_t* mbar_turn = reinterpret_cast<uint64_t*>(tmem_slot + 2);           // [2]

    const uint32_t cta = cluster_ctarank();
    // cluster index from blockIdx (NOT %clusterid: the output of an asm statement is not provably warp-uniform, and dir / cg feed
    // the control flow around the MMA issue through the active ranges below)
    const int cid = (int)(blockIdx.x / (unsigned)NC);
    const int dir = p.ndir == 2 ? (cid & 1) : 0, cg = p.ndir == 2 ? (cid >> 1) : cid;   // direction, utterance group of the cluster
    // utterance groups of this cluster (fwd_layout(): the first n_paired clusters of a direction hold two groups, the others one)
    const int npair = NSG == 2 ? p.n_paired : 0;
    auto group_of = [&](int g2) { return cg < npair ? 2 * cg + g2 : (g2 == 0 ? npair + cg : -1); };
    const int bg = group_of(sg);
    const int b0 = bg >= 0 ? bg * BG : B;              // first utterance of this sub-group (B: none)
    const size_t ld2 = (size_t)2 * Hp;
    // Scan steps this launch covers, and the part of them in which a group has a live utterance at all: the forward direction of
    // a group whose longest utterance has Lm frames is finished after scan step Lm - 1, its backward direction (descending t,
    // mask t < len) only starts at scan step T - Lm.  Outside that range every role of the sub-group skips the step -- no weight
    // pass, no exchange -- and the compute warps just write the zero rows of m; the roles below run the LOCAL range [S0, S0 + S).
    const int LS0 = p.s_begin, LS = p.s_end - p.s_begin;
    // (the issue loop below must stay PROVABLY warp-uniform, or the compiler wraps every tcgen05.mma in an ELECT / R2UR waterfall
    // again -- 86 instead of 29 cycles per MMA, measured)
    auto active_range = [&](int g2, int& a0, int& a1) {      // active steps of sub-group g2, as offsets into [LS0, LS0 + LS)
        const int gr = group_of(g2);
        // the group's longest utterance comes from the KERNEL PARAMETERS (host-computed, lcb_lstm_rec_fwd_range_hl): a value
        // loaded from memory -- even from a uniform address, even laundered through a shuffle -- is not provably warp-uniform,
        // and every range below feeds the control flow around the MMA issue
        // (static indices only: a dynamically indexed by-value parameter array is copied to local memory first)
        int Lm = gr < 0 ? 0 : T;
#pragma unroll
        for (int k = 0; k < REC_MAXGRP; ++k)
            if (k == gr) Lm = min(max(p.gmax[k], 0), T);
        const int lo = dir ? T - Lm : 0, hi = dir ? T : Lm;
        a0 = min(max(lo - LS0, 0), LS);
        a1 = max(min(max(hi - LS0, 0), LS), a0);
    };
    int actA0, actA1, actB0 = 0, actB1 = 0;                  // sub-group 0, sub-group 1
    active_range(0, actA0, actA1);
    if (NSG == 2) active_range(1, actB0, actB1);
    const int act0_me = sg == 0 ? actA0 : actB0, act1_me = sg == 0 ? actA1 : actB1;
    const int S0 = LS0 + act0_me, S = act1_me - act0_me;     // this sub-group runs scan steps S0 .. S0+S-1 (local index s = 0..S-1)
    const int Koth = NSG == 2 ? (sg == 0 ? actB1 - actB0 : actA1 - actA0) : 0;   // weight passes of the other sub-group (turn taking)

    if (role == 0 && rw == 0 && lane == 0) {
        mbar_init(mbar_mma, NIW);
        mbar_init(mbar_acc, NCW);
        for (int s = 0; s < SG; ++s) { mbar_init(&mbar_g[s], 1); mbar_init(&mbar_gfree[s], NCW); }
        mbar_init(&mbar_op[0], 1); mbar_init(&mbar_op[1], 1);
        mbar_init(&mbar_slice[0], NCW); mbar_init(&mbar_slice[1], NCW);
        if (sg == 0) { mbar_init(&mbar_turn[0], NIW); mbar_init(&mbar_turn[1], NIW); }
        fence_mbar_init();
    }
    if (warp == NSG * NCW) tmem_alloc<512>(tmem_slot);
    {   // operand buffer 0 := m of the step before S0 (0 for a fresh start; a padded frame's saved m is 0 as well); the
        // whole block fills the buffers of every sub-group
        const int n16 = (int)(2 * OPB / 16);
        for (int g2 = 0; g2 < NSG; ++g2) {
            uint4* bz = reinterpret_cast<uint4*>(smem_all + (size_t)g2 * Cfg::sg_bytes(KB));
            for (int i = threadIdx.x; i < n16; i += blockDim.x) bz[i] = make_uint4(0u, 0u, 0u, 0u);
        }
        __syncthreads();
        const int pieces = (Hp >> 3) * BG;               // 16-byte pieces: (unit chunk of 8, utterance)
        // (one flat, predicated loop: branches on the computed ranges around this thread-divergent loop cost the compiler its
        // proof that the issuer warps below are converged)
        for (int i = threadIdx.x; i < NSG * pieces; i += blockDim.x) {
            const int g2 = i / pieces, i2 = i - g2 * pieces;
            const int u = i2 % BG, ch = i2 / BG;
            const int S02 = LS0 + (g2 == 0 ? actA0 : actB0), n2 = g2 == 0 ? actA1 - actA0 : actB1 - actB0;
            const int tp = dir ? (T - S02) : (S02 - 1);
            const int gr2 = group_of(g2), b2 = gr2 * BG + u;
            // only a frame that was live for the utterance carries state (a row past its end is zero -- and may be one this
            // very launch is still zero-filling)
            const bool have = S02 > 0 && n2 > 0 && gr2 >= 0 && b2 < B && tp < __ldg(p.lens + (gr2 >= 0 && b2 < B ? b2 : 0));
            if (have) {
                const uint4 v = *reinterpret_cast<const uint4*>(p.Mout + ((size_t)tp * B + b2) * ld2 + (size_t)dir * Hp + ch * 8);
                *reinterpret_cast<uint4*>(smem_all + (size_t)g2 * Cfg::sg_bytes(KB) + ((size_t)ch * NUB + (u >> 3)) * 128 + (u & 7) * 16) = v;
            }
        }
    }
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);      // (warp-uniform to the compiler as well)

    // ---- W'^T slice -> tensor memory: row r of the slice lives in TMEM lane r, 16 fp16 per 8 columns ----
    if (role == 0) {                                   // (the compute warps of all sub-groups share the work)
        const int q = warp & 3;
        const __half* wrow = p.Wt + ((size_t)dir * 4 * Hp + (size_t)cta * 128 + q * 32 + lane) * Hp;
        for (int ch = sg * NUB + (wl >> 2); ch < Hp / 16; ch += NUB * NSG) {
            const uint4 v0 = __ldg(reinterpret_cast<const uint4*>(wrow + ch * 16));
            const uint4 v1 = __ldg(reinterpret_cast<const uint4*>(wrow + ch * 16 + 8));
            const uint32_t r[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
            tmem_st_32x32b_x8(tmem_base + ((uint32_t)(q * 32) << 16) + ch * 8, r);
        }
        tmem_st_wait();
    }
    tc_fence_before();
    cluster_sync_all();          // weights resident; every CTA of the cluster is initialised before any multicast traffic
    tc_fence_after();

    // (debug probes: bits 16.. of the step count select the sub-group that is sampled)
    long long* prof = (blockIdx.x == 0 && lane == 0 && sg == (g_rec_prof_steps >> 16) && ((role == 1 && rw == 0) || (role == 0 && rw == 0))) ? g_rec_prof : nullptr;
    const int prof_steps = g_rec_prof_steps & 0xffff;
    const int nvalid = (B - b0) < BG ? (B - b0) : BG;           // utterances of this group that exist
    int* const prog = p.progress ? p.progress + (size_t)(cid * NSG + sg) * NC + cta : nullptr;
    if (nvalid <= 0) role = 4;                                  // an empty second sub-group (in every CTA of the cluster alike) idles

    bool ok = true;
    if (role == 2) {
        // ============================ loader warp: G tile prefetch ============================
        // ONE TMA tensor load per tile (box = 128 packed gate columns x BG utterances of frame t).  The first version issued one
        // 512-byte bulk copy per utterance: 16 (paired: 32) operations per step in the SM's TMA queue, in a burst right when the
        // exchange warp issues the bulk store + multicast of m_t that the serial chain waits for -- measured 1.36 -> 1.27 us per
        // step (paired groups), 1.15 -> 1.09 (a group alone), profiles/r02_rec_g_tile_tma.txt
        int ready_seen = 0;
        for (int s = 0; s < S; ++s) {
            const int stage = s % SG;
            if (s >= SG) {              // wait until the compute warps drained this stage (step s - SG)
                if (lane == 0 && ok) ok = mbar_wait(&mbar_gfree[stage], (uint32_t)(((s - SG) / SG) & 1));
                ok = __shfl_sync(0xffffffffu, ok ? 1 : 0, 0) != 0;
                if (!ok) break;
            }
            const int t = dir ? (T - 1 - (S0 + s)) : (S0 + s);
            // Flow control against the projection GEMMs that still run beside this launch (another stream, the SMs the clusters
            // leave idle): *ready = number of leading scan steps whose G rows are written (frames [0, n) of the forward, [T-n, T)
            // of the backward direction), advanced by lcb_store_i32 behind every projected chunk.  ONE launch then covers the
            // whole scan -- clusters are not re-synchronised at chunk boundaries, which matters once groups of different
            // length run at different speeds -- and only this prefetch warp, 8 steps ahead of the chain, ever waits.
            if (p.ready && S0 + s >= ready_seen) {
                if (lane == 0) {
                    uint64_t t0 = 0; uint32_t spins = 0;
                    int v;
                    while ((v = ld_acquire_gpu_s32(p.ready)) <= S0 + s) {
                        __nanosleep(100);
                        if ((++spins & 0xff) == 0) {
                            const uint64_t now = globaltimer_ns();
                            if (t0 == 0) t0 = now;
                            if (now - t0 > LCB_WAIT_TIMEOUT_NS || dev_has_error()) { dev_set_error(DEV_ERR_MBAR_TIMEOUT); ok = false; break; }
                        }
                    }
                    ready_seen = v;
                    fence_proxy_async_all();              // generic-proxy acquire -> the async-proxy (TMA) read below
                }
                ready_seen = __shfl_sync(0xffffffffu, ready_seen, 0);
                ok = __shfl_sync(0xffffffffu, ok ? 1 : 0, 0) != 0;
                if (!ok) break;
            }
            if (lane == 0) mbar_arrive_expect_tx(&mbar_g[stage], (uint32_t)(BG * (GH ? 256 : 512)));     // the whole box counts (rows past B are zero-filled)
            if (lane == 0) tma_load_3d(Gsm + (size_t)stage * BG * REC_GCOLS, &tmG, &mbar_g[stage], dir * 4 * Hp + (int)cta * 128, b0, t);
        }
    } else if (role == 3) {
        // ============================ exchange warp: slice -> L2 scratch -> multicast into every CTA of the cluster ============================
        // The DSMEM port of an SM moves ~16 B/clk, in + out combined (measured: tools/micro/xch_bench.cu), so the 16 unicast
        // copies of an all-gather cost ~60 cycles per KB of operand; a bulk store to an L2-resident scratch followed by ONE
        // multicast bulk load costs ~750 cycles + 10 per KB and lands everywhere at once.
        if (lane == 0) {
            unsigned char* scr = p.xch + (size_t)(cid * NSG + sg) * 2 * NC * SLICE;
            const uint16_t mask = (uint16_t)((1u << NC) - 1u);
            for (int s = 0; s + 1 < S && ok; ++s) {
                ok = mbar_wait(&mbar_slice[s & 1], (uint32_t)((s >> 1) & 1));
                if (!ok) break;
                unsigned char* g = scr + ((size_t)(s & 1) * NC + cta) * SLICE;
                bulk_store_s2g(g, smem_u32(Msm + (s & 1) * SLICE), (uint32_t)SLICE);
                bulk_commit_group();
                bulk_wait_group_all();
                bulk_load_multicast(smem_u32(Bsm) + (uint32_t)((s + 1) & 1) * OPB + cta * (uint32_t)SLICE, g, (uint32_t)SLICE,
                                    smem_u32(&mbar_op[(s + 1) & 1]), mask);
                // Progress for the projection GEMMs that consume finished Mout rows beside this launch (lcb_wait_progress on their
                // stream): every compute warp stored its Mout rows of step s-1 (and the zero rows before S0) before it arrived on
                // mbar_slice for step s (release.cta, acquired by the wait above), so a gpu-scope release here covers them.
                // Every 16th step, behind the multicast: off the chain.
                if (prog && (s & 15) == 15) st_release_gpu_s32(prog, S0 + s);
            }
        }
        __syncwarp();
    } else if (role == 1) {
        // ============================ MMA issuers ============================
        // forward operands are fp16: a_format = b_format = F16 (0); A (weights) from TMEM, B = m_{t-1} K-major from smem.
        // The whole operand arrives at once (multicast), so each issuer waits ONCE per step and issues its half of the
        // Hp/16 K steps back to back; every MMA ACCUMULATES onto the accumulator the compute warps pre-loaded with the
        // hoisted x-part of the pre-activations (G tile), so issue order between the two threads does not matter.
        constexpr uint32_t idesc = make_idesc_bf16_f32(128, BG, 0, 0) & ~((7u << 7) | (7u << 10));
        const int iw = rw;
        // The issuer warp stays converged and its address arithmetic warp-uniform (canonical warp index, broadcast TMEM base);
        // one elected lane issues and commits.
        const uint32_t leader = elect_one() ? 1u : 0u;
        // LBO (next 8 units along K) = NUB*128 B, SBO (next 8 utterances) = 128 B; one K=16 MMA step = 2*NUB*128 B
        const uint64_t bb0 = make_smem_desc_noswz(smem_u32(Bsm), NUB * 128, 128);
        const uint32_t b_lo0 = (uint32_t)bb0, b_hi = (uint32_t)(bb0 >> 32);
        // each issuer accumulates into ITS OWN accumulator (columns [iw*BG, iw*BG + BG)): the tensor pipe executes one
        // thread's MMAs in issue order, so the partial sums -- and the sum the compute warps form from them -- are
        // bit-reproducible run to run and independent of how issuer threads interleave
        const uint32_t d_tmem = tmem_base + REC_TMEM_ACC + (uint32_t)((sg * NIW + iw) * BG);
        const int nk = Hp >> 4;
        for (int s = 0; s < S && ok; ++s) {
            REC_PROBE(0);
            const uint32_t par = (uint32_t)(s & 1);
            if (leader && iw == 0 && s + 1 < S) mbar_arrive_expect_tx(&mbar_op[(s + 1) & 1], (uint32_t)(NC * SLICE));   // the buffer that step s fills
            ok = mbar_wait(mbar_acc, (uint32_t)(s & 1));                 // accumulator = G tile of step s
            // Two sub-groups: the weight passes take strict turns.  Issued at the same time they share the tensor pipe, both
            // finish late and the sub-groups fall into phase; one after the other, each pass runs at full rate under the other
            // sub-group's exchange and gate math (without turns: forward equal, BPTT 3330 -> 3560 cycles per step,
            // profiles/r01_rec_experiments_lock_stasync.txt).  The turn is awaited BEFORE the operand: in steady state the other
            // sub-group's pass ran while this one's exchange was in flight, and a try_wait on an already completed phase still
            // costs ~100 cycles -- behind the operand wait they sat on the serial chain of every step.
            // Passes are counted per sub-group from its own first step: pass n of sub-group 0 follows pass n-1 of sub-group 1,
            // pass n of sub-group 1 follows pass n of sub-group 0 -- as long as the other sub-group HAS that pass (a group with
            // shorter utterances retires early and its partner then runs alone).  Waits and arrivals pair up one to one.
            if (NSG == 2 && ok) {
                if (sg == 0 ? (s >= 1 && s - 1 < Koth) : (s < Koth)) ok = mbar_wait(&mbar_turn[sg], (uint32_t)((sg ? s : s - 1) & 1));
            }
            REC_PROBE(4);
            if (ok && s > 0) ok = mbar_wait(&mbar_op[par], (uint32_t)(((s - 1) >> 1) & 1));
            if (!ok) break;
            REC_PROBE(1);
            tc_fence_after();
            const uint32_t b_lo_s = b_lo0 + ((par * OPB) >> 4);
            // (no special case for scan step 0, whose m_{-1} = 0: operand buffer 0 starts zero-filled, so the pass adds W' * 0 --
            // and a data-dependent branch around the issue would cost the warp-uniformity proof again)
            // issuer 0 accumulates onto the G tile; the others start their accumulator with a non-accumulating MMA
            if (iw > 0) umma_f16_ts_elect<false>(d_tmem, tmem_base + 8 * iw, b_lo_s + (2 * NUB * 128 / 16) * iw, b_hi, idesc, leader);
            else umma_f16_ts_elect<true>(d_tmem, tmem_base, b_lo_s, b_hi, idesc, leader);
#pragma unroll 4
            for (int kk = iw + NIW; kk < nk; kk += NIW)
                umma_f16_ts_elect(d_tmem, tmem_base + 8 * kk, b_lo_s + (2 * NUB * 128 / 16) * kk, b_hi, idesc, leader);
            REC_PROBE(7);
            if (leader) {
                umma_commit(mbar_mma);
                if (NSG == 2 && (sg == 0 ? (s < Koth) : (s + 1 < Koth))) mbar_arrive(&mbar_turn[sg ^ 1]);
            }
            REC_PROBE(2);
        }
        __syncwarp();
    } else if (role == 0) {
        // ============================ compute warps: (utterance block, TMEM lane quarter) ============================
        const int ub = rw >> 2, q = rw & 3;                    // q == warp % 4: the TMEM lane quarter this warp may access
        const int up = lane >> 2, g = lane & 3;
        const int unit = (int)cta * 32 + q * 8 + up;
        float wf = 0.f, wi = 0.f, wo = 0.f;
        if (p.peep) {
            wf = p.peep[(size_t)(dir * 3 + 0) * Hp + unit];
            wi = p.peep[(size_t)(dir * 3 + 1) * Hp + unit];
            wo = p.peep[(size_t)(dir * 3 + 2) * Hp + unit];
        }
        // this thread's two utterances: local rows ub*8 + 2g + j
        int len_j[2];
        float c_reg[2];
        bool pad_j[2];
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            const int b = b0 + ub * 8 + 2 * g + j;
            pad_j[j] = b >= B;
            len_j[j] = pad_j[j] ? 0 : p.lens[b];
            c_reg[j] = 0.f;
            if (S0 > 0 && !pad_j[j]) {                     // resume: the carried cell state is the saved c of the previous scan step
                const int tp = dir ? (T - S0) : (S0 - 1);  // (a backward-direction utterance that has not started yet carries 0)
                if (tp < len_j[j]) c_reg[j] = p.cst[((size_t)tp * B + b) * ((size_t)2 * Hp) + (size_t)dir * Hp + unit];
            }
        }
        const uint32_t t_addr = tmem_base + ((uint32_t)(q * 32) << 16) + REC_TMEM_ACC + sg * NIW * BG + ub * 8;
        const float fbias = p.forget_bias;
        // running output pointers of utterance j = 0 at this launch's first frame (utterance j = 1 is the next row, + ld2 elements;
        // a scan step moves them by one frame = B rows, backwards for the backward direction)
        const int bj0 = b0 + ub * 8 + 2 * g;
        const long long tstep = (dir ? -1LL : 1LL) * (long long)B * (long long)ld2;
        const size_t idx0 = ((size_t)(dir ? (T - 1 - LS0) : LS0) * B + bj0) * ld2 + (size_t)dir * Hp + unit;
        const bool save = p.gates != nullptr;
        __half* mo_p = p.Mout + idx0;
        const bool twin = p.Mbf != nullptr;
        __nv_bfloat16* mb_p = (twin ? p.Mbf : reinterpret_cast<__nv_bfloat16*>(p.Mout)) + idx0;        // (never dereferenced unless twin)
        uint2* ga_p = (save ? p.gates : reinterpret_cast<uint2*>(p.Mout)) + idx0;        // (never dereferenced unless save)
        float* cs_p = (save ? p.cst : reinterpret_cast<float*>(p.Mout)) + idx0;
        // frame at which an utterance's final state is emitted: its last live step in this direction's own order (-1: never)
        const int t_last[2] = {dir ? (len_j[0] > 0 ? 0 : -1) : len_j[0] - 1, dir ? (len_j[1] > 0 ? 0 : -1) : len_j[1] - 1};
        // accumulator := hoisted x-part of step s2's pre-activations (G tile, forget bias folded in; padding utterances 0),
        // then release the G stage and tell the issuers that they may accumulate onto it
        // shared-space addresses (32-bit, explicit ld/st.shared: the generic-address forms cost an address-space check per access)
        constexpr int GE = GH ? 2 : 4;                            // bytes per staged G element (a stage keeps its fp32-sized slot either way)
        const uint32_t g_addr = smem_u32(Gsm) + (uint32_t)(((ub * 8 + 2 * g) * REC_GCOLS + q * 32 + up) * GE);     // row of utterance j = 0
        const uint32_t ms_addr = smem_u32(Msm) + (uint32_t)((q * NUB + ub) * 128 + (2 * g) * 16 + up * 2);        // staged m_t, utterance j = 0
        auto load_acc = [&](int s2) {
            const int stage = s2 % SG;
            if (ok) ok = mbar_wait(&mbar_g[stage], (uint32_t)((s2 / SG) & 1));
            const uint32_t ga = g_addr + (uint32_t)(stage * BG * REC_GCOLS * 4);
            uint32_t a0[4], a1[4];
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                const uint32_t gr = ga + (uint32_t)(j * REC_GCOLS * GE);
                if constexpr (GH) {
                    a0[j] = pad_j[j] ? 0u : __float_as_uint(lds_f16(gr));
                    a0[2 + j] = pad_j[j] ? 0u : __float_as_uint(lds_f16(gr + 16));
                    a1[j] = pad_j[j] ? 0u : __float_as_uint(lds_f16(gr + 32) + fbias);
                    a1[2 + j] = pad_j[j] ? 0u : __float_as_uint(lds_f16(gr + 48));
                } else {
                    a0[j] = pad_j[j] ? 0u : lds_b32(gr);
                    a0[2 + j] = pad_j[j] ? 0u : lds_b32(gr + 32);
                    a1[j] = pad_j[j] ? 0u : __float_as_uint(__uint_as_float(lds_b32(gr + 64)) + fbias);
                    a1[2 + j] = pad_j[j] ? 0u : lds_b32(gr + 96);
                }
            }
            tmem_st_16x256b_x1(t_addr, a0);
            tmem_st_16x256b_x1(t_addr + (16u << 16), a1);
            // (the second issuer's accumulator needs no zero fill: its first MMA of a step overwrites it; before scan step 0,
            // which issues no MMAs, it is zeroed once below)
            tmem_st_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) { mbar_arrive(&mbar_gfree[stage]); mbar_arrive(mbar_acc); }
        };
        if (NIW > 1) {
            const uint32_t zz[4] = {0u, 0u, 0u, 0u};
            tmem_st_16x256b_x1(t_addr + BG, zz);
            tmem_st_16x256b_x1(t_addr + BG + (16u << 16), zz);
        }
        // steps of this launch before the group's first live frame (backward direction of a group of short utterances): m = 0
        auto zero_steps = [&](int n) {
            for (int i = 0; i < n; ++i) {
                if (!pad_j[0]) *mo_p = __float2half_rn(0.f);
                if (!pad_j[1]) mo_p[ld2] = __float2half_rn(0.f);
                if (twin) {
                    if (!pad_j[0]) *mb_p = __float2bfloat16_rn(0.f);
                    if (!pad_j[1]) mb_p[ld2] = __float2bfloat16_rn(0.f);
                }
                mo_p += tstep; mb_p += tstep;
            }
            ga_p += (long long)n * tstep; cs_p += (long long)n * tstep;
        };
        zero_steps(act0_me);
        if (S > 0) load_acc(0);

        for (int s = 0; s < S; ++s) {
            const int t = dir ? (T - 1 - (S0 + s)) : (S0 + s);
            REC_PROBE(8);
            // rows of the quarter are gate-major: lanes [0,16) hold gates i,j ; lanes [16,32) gates f,o.  The accumulator
            // already contains the x-part (put there by load_acc below), so this is the complete pre-activation.
            if (ok) ok = mbar_wait(mbar_mma, (uint32_t)(s & 1));
            REC_PROBE(10);
            tc_fence_after();
            float zi[2], zj[2], zf[2], zo[2];
            {
                uint32_t a0[4], a1[4], b0[4] = {0u, 0u, 0u, 0u}, b1[4] = {0u, 0u, 0u, 0u};
                tmem_ld_16x256b_x1(t_addr, a0);                       // (i | j) x 2 utts: x-part + issuer 0's K steps
                tmem_ld_16x256b_x1(t_addr + (16u << 16), a1);         // (f | o)
                if (NIW > 1) {
                    tmem_ld_16x256b_x1(t_addr + BG, b0);              // issuer 1's K steps
                    tmem_ld_16x256b_x1(t_addr + BG + (16u << 16), b1);
                }
                tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    zi[j] = __uint_as_float(a0[j]) + __uint_as_float(b0[j]); zj[j] = __uint_as_float(a0[2 + j]) + __uint_as_float(b0[2 + j]);
                    zf[j] = __uint_as_float(a1[j]) + __uint_as_float(b1[j]); zo[j] = __uint_as_float(a1[2 + j]) + __uint_as_float(b1[2 + j]);
                }
            }
            REC_PROBE(11);
            float ig[2], fg[2], jt[2], cn[2], og[2], tc[2], mo[2];
            bool live[2];
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                const float cp = c_reg[j];
                sig_sig_tanh(zi[j] + wi * cp, zf[j] + wf * cp, zj[j], ig[j], fg[j], jt[j]);
                cn[j] = fg[j] * cp + ig[j] * jt[j];
                sig_tanh(zo[j] + wo * cn[j], cn[j], og[j], tc[j]);
                live[j] = t < len_j[j];
                if (live[j]) c_reg[j] = cn[j];
                mo[j] = live[j] ? og[j] * tc[j] : 0.f;
            }
            const __half2 mh = __floats2half2_rn(mo[0], mo[1]);
            const uint32_t mhb = *reinterpret_cast<const uint32_t*>(&mh);
            REC_PROBE(15);
            if (s + 1 < S) {
                // stage the warp's [8 utts][8 units] fp16 block = core matrix (q, ub) of this CTA's operand slice (double-
                // buffered by step parity) and hand it to the exchange warp
                const uint32_t ma = ms_addr + (uint32_t)((s & 1) * SLICE);
                sts_b16(ma, (uint16_t)(mhb & 0xffffu));
                sts_b16(ma + 16, (uint16_t)(mhb >> 16));
                fence_proxy_async_smem();                              // generic-proxy stores -> visible to the bulk (async-proxy) store
                __syncwarp();
                REC_PROBE(12);
                if (lane == 0) mbar_arrive(&mbar_slice[s & 1]);
                load_acc(s + 1);                                       // off the chain: the operand is >= 1000 cycles away
            }
            REC_PROBE(13);
            // ---- off the critical path: outputs and saved activations (running pointers: one add per array and step) ----
            if (!pad_j[0]) {
                *mo_p = __low2half(mh);
                if (twin) *mb_p = __float2bfloat16_rn(mo[0]);
                if (save) {
                    const __half2 g01 = __floats2half2_rn(ig[0], jt[0]), g23 = __floats2half2_rn(fg[0], og[0]);
                    *ga_p = make_uint2(*reinterpret_cast<const uint32_t*>(&g01), *reinterpret_cast<const uint32_t*>(&g23));
                    *cs_p = cn[0];
                }
                if (t == t_last[0] && p.cfin) {          // final state = state at the last live step in this direction's own order
                    p.cfin[((size_t)bj0 * 2 + dir) * Hp + unit] = cn[0];
                    p.mfin[((size_t)bj0 * 2 + dir) * Hp + unit] = mo[0];
                }
            }
            if (!pad_j[1]) {
                mo_p[ld2] = __high2half(mh);
                if (twin) mb_p[ld2] = __float2bfloat16_rn(mo[1]);
                if (save) {
                    const __half2 g01 = __floats2half2_rn(ig[1], jt[1]), g23 = __floats2half2_rn(fg[1], og[1]);
                    ga_p[ld2] = make_uint2(*reinterpret_cast<const uint32_t*>(&g01), *reinterpret_cast<const uint32_t*>(&g23));
                    cs_p[ld2] = cn[1];
                }
                if (t == t_last[1] && p.cfin) {
                    p.cfin[((size_t)(bj0 + 1) * 2 + dir) * Hp + unit] = cn[1];
                    p.mfin[((size_t)(bj0 + 1) * 2 + dir) * Hp + unit] = mo[1];
                }
            }
            mo_p += tstep; mb_p += tstep; ga_p += tstep; cs_p += tstep;
            REC_PROBE(14);
        }
        zero_steps(LS - act1_me);                 // ... and behind its last one (forward direction)
        if (prog) {                               // every row of this sub-group and CTA is written: the launch's last scan step
            asm volatile("bar.sync %0, %1;" ::"r"(1 + sg), "n"(NCW * 32) : "memory");      // the sub-group's compute warps
            if (rw == 0 && lane == 0) { __threadfence(); st_release_gpu_s32(prog, p.s_end); }
        }
    }
    if (role == 4 && prog && wl == 0 && lane == 0) st_release_gpu_s32(prog, p.s_end);      // an empty sub-group has nothing to write
    tc_fence_before();
    cluster_sync_all();          // nobody leaves while multicast traffic addressed to it may still be in flight
    if (warp == NSG * NCW) { tc_fence_after(); tmem_dealloc<512>(tmem_base); }
}

// ---- BPTT v2 ----
template <int BG> struct RecBwd2Cfg {
    static constexpr int NUB = BG / 8;
    static constexpr int NCW = 4 * NUB;                    // compute warps: (utterance block | M tile, TMEM lane quarter)
    static constexpr int THREADS = 32 * (NCW + 4);         // + 4 MMA issuer warps (one per 128-unit M tile)
    static constexpr int TPW = 16 / NCW;                   // (M tile, quarter) partial tiles per compute warp
    static constexpr int PT = 32 * BG * 2;                 // bytes of one partial tile: [32 units][BG utts] bf16
    static constexpr int BARS = 8;                         // mma[4] dz red[2] (+pad)
    __host__ __device__ static size_t bp_bytes() { return (size_t)2 * BG * 128; }              // dz operand: 2 K blocks x BG rows x 128 B
    __host__ __device__ static size_t red_bytes(int NC) { return (size_t)NC * PT; }            // one reduce buffer
    static size_t smem_bytes(int NC) { return 1024 + 2 * bp_bytes() + 2 * red_bytes(NC) + (size_t)2 * 16 * PT + BARS * 8 + 64; }
};

template <int BG>
__global__ void __launch_bounds__(RecBwd2Cfg<BG>::THREADS, 1)
lstm_rec_bwd2_kernel(const RecBwdParams p)
{
    using Cfg = RecBwd2Cfg<BG>;
    constexpr int NUB = Cfg::NUB, NCW = Cfg::NCW, TPW = Cfg::TPW, PT = Cfg::PT;
    constexpr int NCH = BG / 8;                        // 16-byte chunks per partial-tile row (8 utterances each)
    constexpr int SWS = 64 / BG;                       // row-swizzle period: chunk c of unit row u sits at c ^ ((u / SWS) & (NCH-1))
    extern __shared__ unsigned char smem_raw[];
    unsigned char* smem = align_1024(smem_raw);
    const int Hp = p.Hp, NC = p.NC, T = p.T, B = p.B;
    const int MB = (Hp + 127) >> 7;                    // M tiles of 128 units
    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;   // provably warp-uniform
    const int role = warp < NCW ? 0 : 1;
    const int rw = role == 0 ? warp : warp - NCW;

    // operand dz_t (K' = 128 gate rows), SW128 K-major; double-buffered by step parity: a CTA's reduce buffer can complete
    // (and its next phase A start) while its OWN later M tiles of the previous step are still being multiplied
    unsigned char* Bp = smem;                                                    // [2][2 K blocks][BG rows x 128 B]
    unsigned char* red = Bp + 2 * Cfg::bp_bytes();                               // [2][NC][PT]
    unsigned char* pst = red + 2 * Cfg::red_bytes(NC);                           // [2 step parities][16 tiles][PT] staging
    uint64_t* bars = reinterpret_cast<uint64_t*>(pst + (size_t)2 * 16 * PT);
    uint64_t* mbar_mma = bars;                         // [4] partial dm tile j complete
    uint64_t* mbar_dz = bars + 4;                      //     dz_t staged by all compute warps
    uint64_t* mbar_red = bars + 5;                     // [2] partial dm slices from the whole cluster have landed
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + Cfg::BARS);

    const uint32_t cta = cluster_ctarank();
    const int cid = (int)cluster_id_x();
    const int dir = p.ndir == 2 ? (cid & 1) : 0, bg = p.ndir == 2 ? (cid >> 1) : cid;
    const int b0 = bg * BG;
    const size_t ld2 = (size_t)2 * Hp, ld8 = (size_t)8 * Hp;
    const uint32_t red_bytes = (uint32_t)Cfg::red_bytes(NC);

    if (threadIdx.x == 0) {
        for (int i = 0; i < 4; ++i) mbar_init(&mbar_mma[i], 1);
        mbar_init(mbar_dz, NCW);
        mbar_init(&mbar_red[0], 1);
        mbar_init(&mbar_red[1], 1);
        fence_mbar_init();
    }
    if (warp == NCW) tmem_alloc<512>(tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);      // (warp-uniform to the compiler as well)

    // ---- W' -> tensor memory as A' = [units (lanes), own 128 gate rows (K')]: tile j in columns [64 j, 64 j + 64) ----
    if (role == 0) {
        const int q = warp & 3;
        for (int jt = warp >> 2; jt < MB; jt += NUB) {
            const int u = jt * 128 + q * 32 + lane;
            const __nv_bfloat16* wrow = p.W + ((size_t)dir * Hp + (u < Hp ? u : 0)) * 4 * Hp + (size_t)cta * 128;
            for (int ch = 0; ch < 8; ++ch) {
                uint4 v0 = make_uint4(0u, 0u, 0u, 0u), v1 = v0;
                if (u < Hp) {
                    v0 = __ldg(reinterpret_cast<const uint4*>(wrow + ch * 16));
                    v1 = __ldg(reinterpret_cast<const uint4*>(wrow + ch * 16 + 8));
                }
                const uint32_t r[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
                tmem_st_32x32b_x8(tmem_base + ((uint32_t)(q * 32) << 16) + jt * 64 + ch * 8, r);
            }
        }
        tmem_st_wait();
    }
    if (threadIdx.x == 0 && T > 1) mbar_arrive_expect_tx(&mbar_red[0], red_bytes);   // armed before anybody can send
    tc_fence_before();
    cluster_sync_all();
    tc_fence_after();

    long long* prof = (blockIdx.x == 0 && lane == 0 && rw == 0) ? g_rec_prof : nullptr;
    const int prof_steps = g_rec_prof_steps;
    bool ok = true;
    if (role == 1) {
        // ============================ MMA issuer warps: one 128-unit M tile each ============================
        constexpr uint32_t idesc = make_idesc_bf16_f32(128, BG, 0, 0);           // bf16 x bf16, A (TMEM) K-major
        const int jt = rw;
        // converged issuer warp, warp-uniform address arithmetic, one elected lane issues and commits (see umma_f16_ts_elect)
        const uint32_t leader = elect_one() ? 1u : 0u;
        if (jt < MB) {
            const uint32_t a_tmem = tmem_base + jt * 64;
            const uint64_t bb0 = make_smem_desc_sw128(smem_u32(Bp), 16, 1024);
            const uint32_t b_lo0 = (uint32_t)bb0, b_hi = (uint32_t)(bb0 >> 32);
            const uint32_t d_tmem = tmem_base + REC_TMEM_ACC + jt * BG;
            for (int s = 0; s < T && ok; ++s) {
                REC_PROBE(0);
                ok = mbar_wait(mbar_dz, (uint32_t)(s & 1));                      // dz_t staged by all compute warps
                if (!ok) break;
                REC_PROBE(1);
                // the reduce buffer the NEXT step's partials go to: its previous contents were consumed in phase A
                // of this step (all compute warps arrived on mbar_dz after reading them)
                if (leader && jt == 0 && s + 2 < T) mbar_arrive_expect_tx(&mbar_red[(s + 1) & 1], red_bytes);
                tc_fence_after();
                if (s + 1 < T) {                      // the last step's dm_{-1} is never used
                    const uint32_t b_lo_s = b_lo0 + (uint32_t)(((s & 1) * (int)Cfg::bp_bytes()) / 16);
                    umma_f16_ts_elect<false>(d_tmem, a_tmem, b_lo_s, b_hi, idesc, leader);
#pragma unroll
                    for (int kk = 1; kk < 8; ++kk)
                        umma_f16_ts_elect<true>(d_tmem, a_tmem + 8 * kk, b_lo_s + (uint32_t)(((kk >> 2) * (BG * 128)) / 16 + (kk & 3) * 2), b_hi, idesc, leader);
                }
                REC_PROBE(7);
                if (leader) umma_commit(&mbar_mma[jt]);
                REC_PROBE(2);
            }
        }
        __syncwarp();
    } else {
        // ============================ compute warps ============================
        const int ub = rw >> 2, q = rw & 3;
        const int up = lane >> 2, g = lane & 3;
        const int ul = q * 8 + up;                          // local unit (0..31)
        const int unit = (int)cta * 32 + ul;
        float wf = 0.f, wi = 0.f, wo = 0.f;
        if (p.peep) {
            wf = p.peep[(size_t)(dir * 3 + 0) * Hp + unit];
            wi = p.peep[(size_t)(dir * 3 + 1) * Hp + unit];
            wo = p.peep[(size_t)(dir * 3 + 2) * Hp + unit];
        }
        int len_j[2];
        float dcc[2];
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            const int b = b0 + ub * 8 + 2 * g + j;
            len_j[j] = (b < B) ? p.lens[b] : 0;
            dcc[j] = 0.f;
        }
        float db[4] = {0.f, 0.f, 0.f, 0.f};
        float dpf = 0.f, dpi = 0.f, dpo = 0.f;

        // raw prefetch of the next step's saved activations: loads only, no arithmetic, so they stay in flight
        // behind the current step.  Row indices advance by a constant stride per step.
        struct Pre { uint2 gp[2]; float c[2], cp[2], dmo[2]; };
        const long long row_stride = (dir ? 1 : -1) * (long long)B * (long long)ld2;      // elements per time step, in scan order
        long long idx_j[2];                                                                // element index of (t(s), b_j, dir, unit)
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            const int b = b0 + ub * 8 + 2 * g + j;
            idx_j[j] = ((long long)(dir ? 0 : T - 1) * B + (b < B ? b : 0)) * (long long)ld2 + (long long)dir * Hp + unit;
        }
        auto load_pre = [&](int s, Pre& r) {          // loads step s (idx_j must already point at step s)
            const int t = dir ? s : (T - 1 - s);
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                const bool live = (s < T) && (t < len_j[j]);
                r.gp[j] = make_uint2(0u, 0u); r.c[j] = 0.f; r.cp[j] = 0.f; r.dmo[j] = 0.f;
                if (live) {
                    r.gp[j] = __ldg(p.gates + idx_j[j]);
                    r.c[j] = __ldg(p.cst + idx_j[j]);
                    r.dmo[j] = __ldg(p.dM + idx_j[j]);
                    // previous step in the direction's own order: fwd t-1, bwd t+1 (zero initial state) = the row one
                    // stride AHEAD in scan order
                    const int tp = dir ? (t + 1) : (t - 1);
                    const bool has_prev = dir ? (tp < len_j[j]) : (tp >= 0);
                    if (has_prev) r.cp[j] = __ldg(p.cst + idx_j[j] + row_stride);
                }
            }
        };
        Pre cur, nxt;
        load_pre(0, cur);
        // reduce-buffer read offset of this thread: unit row ul, 16-byte chunk ub (swizzled), utterance pair g
        const uint32_t rd_off = (uint32_t)(ul * (2 * BG) + ((ub ^ ((ul / SWS) & (NCH - 1))) << 4) + 4 * g);
        // per-thread constants of the dz staging: smem offsets of the local MMA B operand and global columns
        uint32_t bp_off[2][4];
        int gcol[4];
#pragma unroll
        for (int gate = 0; gate < 4; ++gate) {
            gcol[gate] = dir * 4 * Hp + packed_col(unit, gate);
            const int kp = packed_col(ul, gate);
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                const int bl = ub * 8 + 2 * g + j;
                bp_off[j][gate] = (uint32_t)((kp >> 6) * (BG * 128) + (bl >> 3) * 1024 + (bl & 7) * 128 +
                                             ((((kp & 63) >> 3) ^ (bl & 7)) << 4) + (kp & 7) * 2);
            }
        }
        long long grow_j[2];                          // element index of dG row (t(s), b_j)
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            const int b = b0 + ub * 8 + 2 * g + j;
            grow_j[j] = ((long long)(dir ? 0 : T - 1) * B + (b < B ? b : 0)) * (long long)ld8;
        }
        const long long grow_stride = (dir ? 1 : -1) * (long long)B * (long long)ld8;
        const uint32_t red_addr = smem_u32(red);

        for (int s = 0; s < T; ++s) {
            const int t = dir ? s : (T - 1 - s);
            REC_PROBE(8);
#pragma unroll
            for (int j = 0; j < 2; ++j) idx_j[j] += row_stride;
            load_pre(s + 1, nxt);                                  // latency hidden behind this step
            REC_PROBE(9);
            // ---- phase A: dm_rec from the reduce buffer of the previous step, then dz_t ----
            float dmr[2] = {0.f, 0.f};
            if (s > 0) {
                if (ok) ok = mbar_wait_cluster_acq(&mbar_red[(s - 1) & 1], (uint32_t)(((s - 1) >> 1) & 1));
                REC_PROBE(10);
                const unsigned char* rb = red + (size_t)((s - 1) & 1) * red_bytes + rd_off;
                float2 accv[4] = {{0.f, 0.f}, {0.f, 0.f}, {0.f, 0.f}, {0.f, 0.f}};
                int src = 0;
                for (; src + 4 <= NC; src += 4) {                  // 4 independent loads in flight
                    __nv_bfloat162 v[4];
#pragma unroll
                    for (int k = 0; k < 4; ++k) v[k] = *reinterpret_cast<const __nv_bfloat162*>(rb + (size_t)(src + k) * PT);
#pragma unroll
                    for (int k = 0; k < 4; ++k) { const float2 f = __bfloat1622float2(v[k]); accv[k].x += f.x; accv[k].y += f.y; }
                }
                for (; src < NC; ++src) {
                    const float2 f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(rb + (size_t)src * PT));
                    accv[0].x += f.x; accv[0].y += f.y;
                }
                dmr[0] = (accv[0].x + accv[1].x) + (accv[2].x + accv[3].x);
                dmr[1] = (accv[0].y + accv[1].y) + (accv[2].y + accv[3].y);
            }
            unsigned char* Bps = Bp + (size_t)(s & 1) * Cfg::bp_bytes();
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                const int b = b0 + ub * 8 + 2 * g + j;
                const bool live = t < len_j[j];
                float dzi = 0.f, dzj = 0.f, dzf = 0.f, dzo = 0.f;
                if (live) {
                    const float2 g01 = __half22float2(*reinterpret_cast<const __half2*>(&cur.gp[j].x));
                    const float2 g23 = __half22float2(*reinterpret_cast<const __half2*>(&cur.gp[j].y));
                    const float ig = g01.x, jt = g01.y, fg = g23.x, og = g23.y, cp = cur.cp[j];
                    const float tc = tanhf_fast(cur.c[j]);
                    const float dm = cur.dmo[j] + dmr[j];
                    dzo = dm * tc * og * (1.f - og);
                    const float dc = dcc[j] + dm * og * (1.f - tc * tc) + dzo * wo;
                    dzf = dc * cp * fg * (1.f - fg);
                    dzi = dc * jt * ig * (1.f - ig);
                    dzj = dc * ig * (1.f - jt * jt);
                    dcc[j] = dc * fg + dzi * wi + dzf * wf;
                    dpi += dzi * cp; dpf += dzf * cp; dpo += dzo * cur.c[j];
                    db[0] += dzi; db[1] += dzj; db[2] += dzf; db[3] += dzo;
                } else {
                    dcc[j] = 0.f;
                }
                const float dzg[4] = {dzi, dzj, dzf, dzo};
                // gate columns are gate-major inside each 8-unit group: col(unit, gate) = packed_col(unit, gate)
                __nv_bfloat16* dgrow = p.dG + grow_j[j];
#pragma unroll
                for (int gate = 0; gate < 4; ++gate) {
                    const __nv_bfloat16 v = __float2bfloat16(dzg[gate]);
                    *reinterpret_cast<__nv_bfloat16*>(Bps + bp_off[j][gate]) = v;     // local MMA B operand dz_t (K-major, 128B swizzle)
                    if (b < B) dgrow[gcol[gate]] = v;
                }
                grow_j[j] += grow_stride;
            }
            REC_PROBE(11);
            fence_proxy_async_smem();                  // locally staged dz_t -> visible to the tensor core
            __syncwarp();
            if (lane == 0) mbar_arrive(mbar_dz);       // one arrival per warp (the fences of all its lanes precede it)
            REC_PROBE(12);
            // ---- phase B: partial dm_{t-1} tiles -> owners' reduce buffers (bulk DSMEM copies) ----
            if (s + 1 < T) {
#pragma unroll 1
                for (int k = 0; k < TPW; ++k) {
                    const int jt = (rw >> 2) + k * NUB;            // this warp's M tile; its TMEM lane quarter is q
                    if (jt >= MB) break;
                    if (ok) ok = mbar_wait(&mbar_mma[jt], (uint32_t)(s & 1));
                    if (k == 0) { REC_PROBE(13); }
                    tc_fence_after();
                    // the 32 rows of this (tile, quarter) are 32 consecutive units of ONE owner CTA: stage them as a
                    // [32 units][BG utts] bf16 tile (16-byte chunks XOR-swizzled by the unit row so that neither these
                    // stores nor the owner's reads conflict) and ship it with a single bulk DSMEM copy
                    unsigned char* pw = pst + (size_t)(((s & 1) * 16 + jt * 4 + q) * PT);
#pragma unroll
                    for (int h = 0; h < BG / 16; ++h) {
                        uint32_t a[16];
                        tmem_ld_32x32b_x16(tmem_base + ((uint32_t)(q * 32) << 16) + REC_TMEM_ACC + jt * BG + h * 16, a);
                        tmem_ld_wait();
                        const int sw = (lane / SWS) & (NCH - 1);
                        uint4* row = reinterpret_cast<uint4*>(pw + lane * (2 * BG));
                        row[(2 * h) ^ sw] = make_uint4(pack_bf16x2(__uint_as_float(a[0]), __uint_as_float(a[1])), pack_bf16x2(__uint_as_float(a[2]), __uint_as_float(a[3])),
                                                       pack_bf16x2(__uint_as_float(a[4]), __uint_as_float(a[5])), pack_bf16x2(__uint_as_float(a[6]), __uint_as_float(a[7])));
                        row[(2 * h + 1) ^ sw] = make_uint4(pack_bf16x2(__uint_as_float(a[8]), __uint_as_float(a[9])), pack_bf16x2(__uint_as_float(a[10]), __uint_as_float(a[11])),
                                                           pack_bf16x2(__uint_as_float(a[12]), __uint_as_float(a[13])), pack_bf16x2(__uint_as_float(a[14]), __uint_as_float(a[15])));
                    }
                    __syncwarp();
                    const int u0 = jt * 128 + q * 32;
                    if (lane == 0 && u0 < Hp) {
                        const uint32_t owner = (uint32_t)(u0 >> 5);
                        fence_proxy_async_smem();
                        bulk_copy_s2c(mapa_shared(red_addr + (uint32_t)(s & 1) * red_bytes + cta * (uint32_t)PT, owner),
                                      smem_u32(pw), (uint32_t)PT, mapa_shared(smem_u32(&mbar_red[s & 1]), owner));
                    }
                }
            }
            tc_fence_before();
            REC_PROBE(14);
            cur = nxt;
        }
        // ---- parameter gradients held in registers: reduce the 4 lanes of a unit, then atomics ----
#pragma unroll
        for (int o = 1; o <= 2; o <<= 1) {
#pragma unroll
            for (int k = 0; k < 4; ++k) db[k] += __shfl_xor_sync(0xffffffffu, db[k], o);
            dpf += __shfl_xor_sync(0xffffffffu, dpf, o);
            dpi += __shfl_xor_sync(0xffffffffu, dpi, o);
            dpo += __shfl_xor_sync(0xffffffffu, dpo, o);
        }
        if (g == 0) {
#pragma unroll
            for (int k = 0; k < 4; ++k) atomicAdd(p.dbias + (size_t)dir * 4 * Hp + packed_col(unit, k), db[k]);
            if (p.dpeep) {
                atomicAdd(p.dpeep + (size_t)(dir * 3 + 0) * Hp + unit, dpf);
                atomicAdd(p.dpeep + (size_t)(dir * 3 + 1) * Hp + unit, dpi);
                atomicAdd(p.dpeep + (size_t)(dir * 3 + 2) * Hp + unit, dpo);
            }
        }
    }
    tc_fence_before();
    cluster_sync_all();
    if (warp == NCW) { tc_fence_after(); tmem_dealloc<512>(tmem_base); }
}

// ---- BPTT v3 (Hp = 512): 4 x 4 decomposition of the per-step product dm_{t-1} = W' dz_t ----
// CTA x of the 16-CTA cluster owns 32 units (cell state, gate math) as before, but multiplies the [128 units of M tile
// mr = x & 3] x [512 gate rows of K block kc = x >> 2] block of W' (resident in tensor memory).  Per step:
//   all-gather of dz_t inside the group of 4 CTAs that own K block kc (= x's own group): each CTA stages its
//     [BG utts x 128 gate rows] bf16 slice, bulk-stores it to an L2-resident scratch and multicasts it into the four CTAs;
//   32 MMAs 128 x BG x 16 (two issuer threads, accumulating onto a pre-zeroed accumulator);
//   reduce-scatter of the partial dm tile: quarter q of the accumulator belongs to owner CTA 4 mr + q -> ONE 64*BG-byte
//     DSMEM copy per quarter; every owner receives 4 partial tiles (from the CTAs x' with x' & 3 == owner >> 2).
// DSMEM traffic per CTA and step drops from 2 x BG KB (16-way reduce-scatter, the bottleneck of v1/v2: the SM's DSMEM
// port moves ~16 B/clk in + out) to 2 x BG/4 KB; the all-gather rides the L2 multicast path (~750 cycles + 10 per KB).
// NSG = 2 (BG = 16): two independent 16-utterance groups per cluster share the resident weights, as in lstm_rec_fwd2_kernel;
// their weight passes take strict turns on the tensor pipe.  Warp order: all compute warps (so that warp % 4 stays the
// TMEM lane quarter), then the issuers, then the exchange warps.
template <int BG, int NSG = 1> struct RecBwd3Cfg {
    static_assert(NSG == 1 || BG == 16, "two sub-groups only for 16-utterance groups");
    static constexpr int NUB = BG / 8;
    static constexpr int NCW = 4 * NUB;                    // compute warps per sub-group: (utterance block, TMEM lane quarter)
    static constexpr int NIW = LCB_REC_NIW;                // MMA issuer warps per sub-group (each with its own accumulator)
    static constexpr int THREADS = 32 * NSG * (NCW + NIW + 1);   // + one exchange warp per sub-group
    static constexpr int SLICE = BG * 256;                 // dz slice of one CTA: 2 K sub-blocks x [BG rows x 128 B]
    static constexpr int PT = 32 * BG * 2;                 // partial dm tile [32 units][BG utts] bf16
    static constexpr int BARS = 10;                        // op[2] slice[2] red[2] acc mma (+pad)
    __host__ __device__ static constexpr size_t sg_bytes() { return (2 * 4 * (size_t)SLICE + 2 * (size_t)SLICE + 2 * 4 * (size_t)PT + 2 * 4 * (size_t)PT + BARS * 8 + 1023) & ~(size_t)1023; }
    static size_t smem_bytes() { return 1024 + NSG * sg_bytes() + 64; }
};

template <int BG, int NSG>
__global__ void __launch_bounds__(RecBwd3Cfg<BG, NSG>::THREADS, 1)
lstm_rec_bwd3_kernel(const RecBwdParams p, const __grid_constant__ CUtensorMap tmDG)
{
    using Cfg = RecBwd3Cfg<BG, NSG>;
    constexpr int NUB = Cfg::NUB, NCW = Cfg::NCW, NIW = Cfg::NIW, SLICE = Cfg::SLICE, PT = Cfg::PT;
    constexpr int NCH = BG / 8;                        // 16-byte chunks per partial-tile row (8 utterances each)
    constexpr int SWS = 64 / BG;                       // row-swizzle period: chunk c of unit row u sits at c ^ ((u / SWS) & (NCH-1))
    constexpr int Hp = 512, NC = 16;
    extern __shared__ unsigned char smem_raw[];
    unsigned char* smem_all = align_1024(smem_raw);
    const int T = p.T, B = p.B;
    const int S0 = p.s_begin, S = p.s_end - p.s_begin;   // this launch runs scan steps S0 .. S0+S-1 (local index s = 0..S-1)
    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;   // provably warp-uniform
    int role = warp < NSG * NCW ? 0 : (warp < NSG * (NCW + NIW) ? 1 : 2);       // compute | MMA issuers | exchange
    const int sg = role == 0 ? warp / NCW : (role == 1 ? (warp - NSG * NCW) / NIW : warp - NSG * (NCW + NIW));
    const int rw = role == 0 ? warp - sg * NCW : (role == 1 ? warp - NSG * NCW - sg * NIW : 0);
    unsigned char* smem = smem_all + (size_t)sg * Cfg::sg_bytes();

    unsigned char* Op = smem;                                        // [2 parities][4 src][2 K sub-blocks][BG rows x 128 B]  SW128 K-major
    unsigned char* Stg = Op + 2 * 4 * SLICE;                         // [2 parities][SLICE]   own dz slice (same layout)
    unsigned char* red = Stg + 2 * SLICE;                            // [2 parities][4 src][PT]
    unsigned char* pst = red + 2 * 4 * PT;                           // [2 parities][4 quarters][PT]  partial-tile staging
    uint64_t* bars = reinterpret_cast<uint64_t*>(pst + 2 * 4 * PT);
    uint64_t* mbar_op = bars;                          // [2] the four dz slices of this K block have landed
    uint64_t* mbar_slice = bars + 2;                   // [2] own dz slice staged (count NCW)
    uint64_t* mbar_red = bars + 4;                     // [2] four partial dm tiles have landed
    uint64_t* mbar_acc = bars + 6;                     //     accumulator zeroed (count NCW)
    uint64_t* mbar_mma = bars + 7;                     //     partial dm tile complete (count NIW)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem_all + (size_t)NSG * Cfg::sg_bytes());
    // NSG = 2: the sub-groups' weight passes take strict turns on the tensor pipe.  mbar_turn[g] completes a phase each time the
    // OTHER sub-group has issued a pass (NIW issuer arrivals): waiting on an mbarrier suspends the issuer warp instead of
    // spinning on a shared-memory flag (the step is power-capped; idle warps should not burn issue slots)
    uint64_t* mbar_turn = reinterpret_cast<uint64_t*>(tmem_slot + 2);           // [2]

    const uint32_t cta = cluster_ctarank();
    const int cid = (int)cluster_id_x();
    const int dir = p.ndir == 2 ? (cid & 1) : 0, cg = p.ndir == 2 ? (cid >> 1) : cid;   // direction, utterance group of the cluster
    const int bg = cg * NSG + sg;
    const int b0 = bg * BG;
    const bool paired = NSG == 2 && cg * NSG * BG + BG < B;   // both sub-groups of this cluster hold utterances
    const int kc = (int)(cta >> 2), mr = (int)(cta & 3);
    const size_t ld2 = (size_t)2 * Hp;

    if (lane == 0 && role == 0 && rw == 0) {
        mbar_init(&mbar_op[0], 1); mbar_init(&mbar_op[1], 1);
        mbar_init(&mbar_slice[0], NCW); mbar_init(&mbar_slice[1], NCW);
        mbar_init(&mbar_red[0], 1); mbar_init(&mbar_red[1], 1);
        mbar_init(mbar_acc, NCW);
        mbar_init(mbar_mma, NIW);
        if (sg == 0) { mbar_init(&mbar_turn[0], NIW); mbar_init(&mbar_turn[1], NIW); }
        fence_mbar_init();
    }
    if (warp == NSG * NCW) tmem_alloc<512>(tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);      // (warp-uniform to the compiler as well)

    // ---- W' block -> tensor memory: lane = unit of M tile mr, columns = the 512 packed gate rows of K block kc (bf16 pairs) ----
    if (role == 0) {
        const int q = warp & 3;
        const __nv_bfloat16* wrow = p.W + ((size_t)dir * Hp + mr * 128 + q * 32 + lane) * 4 * Hp + (size_t)kc * 512;
        for (int ch = sg * NUB + (rw >> 2); ch < 32; ch += NUB * NSG) {        // (the compute warps of all sub-groups share the work)
            const uint4 v0 = __ldg(reinterpret_cast<const uint4*>(wrow + ch * 16));
            const uint4 v1 = __ldg(reinterpret_cast<const uint4*>(wrow + ch * 16 + 8));
            const uint32_t r[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
            tmem_st_32x32b_x8(tmem_base + ((uint32_t)(q * 32) << 16) + ch * 8, r);
        }
        // zero this warp's part of the accumulator (every MMA accumulates)
        const uint32_t z[8] = {0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u};
        tmem_st_32x32b_x8(tmem_base + ((uint32_t)(q * 32) << 16) + REC_TMEM_ACC + sg * NIW * BG + (rw >> 2) * 8, z);
        if (NIW > 1) tmem_st_32x32b_x8(tmem_base + ((uint32_t)(q * 32) << 16) + REC_TMEM_ACC + sg * NIW * BG + BG + (rw >> 2) * 8, z);
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(mbar_acc);
    }
    if (lane == 0 && role == 0 && rw == 0 && S0 + 1 < T) mbar_arrive_expect_tx(&mbar_red[0], 4u * PT);      // armed before anybody can send
    tc_fence_before();
    cluster_sync_all();
    tc_fence_after();

    long long* prof = (blockIdx.x == 0 && lane == 0 && rw == 0 && role < 2 && sg == (g_rec_prof_steps >> 16)) ? g_rec_prof : nullptr;
    const int prof_steps = g_rec_prof_steps & 0xffff;
    bool ok = true;
    if (b0 >= B) {                                     // an empty second sub-group (in every CTA of the cluster alike) idles
        if (role == 2 && lane == 0 && p.progress) st_release_gpu_s32(p.progress + (size_t)(cid * NSG + sg) * NC + cta, T);
        role = 3;
    }
    if (role == 2) {
        // ============================ exchange warp: own dz slice -> L2 scratch -> multicast into the 4 CTAs of the group ============================
        // The staged slice is also the layer output dG[t, b0.., own 128 packed gate columns]: two 128B-swizzled TMA tensor
        // stores per step (rows past the batch end are clipped), off the serial chain.
        if (lane == 0) {
            unsigned char* scr = p.xch + (size_t)(cid * NSG + sg) * 2 * NC * SLICE;
            const uint16_t mask = (uint16_t)(0xFu << (4 * kc));
            const int col0 = dir * 4 * Hp + (int)cta * 128;
            int* prog = p.progress ? p.progress + (size_t)(cid * NSG + sg) * NC + cta : nullptr;
            for (int s = 0; s < S && ok; ++s) {
                ok = mbar_wait(&mbar_slice[s & 1], (uint32_t)((s >> 1) & 1));
                if (!ok) break;
                const uint32_t src = smem_u32(Stg + (s & 1) * SLICE);
                const int t = dir ? (S0 + s) : (T - 1 - (S0 + s));
                if (S0 + s + 1 < T) {
                    unsigned char* g = scr + ((size_t)(s & 1) * NC + cta) * SLICE;
                    bulk_store_s2g(g, src, (uint32_t)SLICE);
                    bulk_commit_group();
                    bulk_wait_group_all();                 // scratch copy performed (and the dG stores of step s-1, long done)
                    bulk_load_multicast(smem_u32(Op) + (uint32_t)(((s & 1) * 4 + mr) * SLICE), g, (uint32_t)SLICE, smem_u32(&mbar_op[s & 1]), mask);
                    tma_store_3d(&tmDG, src, col0, b0, t);                     // behind the multicast: off the chain; read out of
                    tma_store_3d(&tmDG, src + BG * 128, col0 + 64, b0, t);     // Stg[s&1] before step s+1's wait_group returns
                    bulk_commit_group();
                    // progress for the GEMMs that consume released dG rows beside this launch: the wait_group above also
                    // covered the dG stores of every earlier step (published here, behind the multicast: off the chain)
                    if (prog) { fence_proxy_async_all(); st_release_gpu_s32(prog, S0 + s); }
                } else {
                    tma_store_3d(&tmDG, src, col0, b0, t);
                    tma_store_3d(&tmDG, src + BG * 128, col0 + 64, b0, t);
                    bulk_commit_group();
                }
            }
            bulk_wait_group_all();                          // dG of the last steps is written before the kernel ends
            if (prog && ok) { fence_proxy_async_all(); st_release_gpu_s32(prog, S0 + S); }
        }
        __syncwarp();
    } else if (role == 1) {
        // ============================ MMA issuers ============================
        constexpr uint32_t idesc = make_idesc_bf16_f32(128, BG, 0, 0);           // bf16 x bf16, A (TMEM) K-major
        const int iw = rw;
        // converged issuer warp, warp-uniform address arithmetic, one elected lane issues and commits (see umma_f16_ts_elect)
        const uint32_t leader = elect_one() ? 1u : 0u;
        const uint64_t bb0 = make_smem_desc_sw128(smem_u32(Op), 16, 1024);
        const uint32_t b_lo0 = (uint32_t)bb0, b_hi = (uint32_t)(bb0 >> 32);
        // one accumulator per issuer (columns [iw*BG, iw*BG + BG)): each thread's MMAs execute in its issue order, so the
        // partial sums and their sum in phase B are bit-reproducible whatever the interleaving of the issuer threads
        const uint32_t d_tmem = tmem_base + REC_TMEM_ACC + (uint32_t)((sg * NIW + iw) * BG);
        for (int s = 0; s < S && S0 + s + 1 < T && ok; ++s) {                 // the last step's dm_{-1} is never used
            REC_PROBE(0);
            if (leader && iw == 0) mbar_arrive_expect_tx(&mbar_op[s & 1], 4u * SLICE);
            ok = mbar_wait(mbar_acc, (uint32_t)(s & 1));                      // accumulator zeroed
            if (ok && paired) {                                               // strict turns, awaited before the operand (see lstm_rec_fwd2_kernel)
                if (sg == 1 || s > 0) ok = mbar_wait(&mbar_turn[sg], (uint32_t)((sg ? s : s - 1) & 1));   // pass n of sub-group 1 follows pass n of sub-group 0
            }
            REC_PROBE(4);
            if (ok) ok = mbar_wait(&mbar_op[s & 1], (uint32_t)((s >> 1) & 1)); // dz_t of the whole K block landed
            if (!ok) break;
            REC_PROBE(1);
            tc_fence_after();
            const uint32_t b_lo_s = b_lo0 + (uint32_t)(((s & 1) * 4 * SLICE) >> 4);
            // idx = src * 8 + kk; the first MMA of a step overwrites this issuer's accumulator, the others accumulate
            umma_f16_ts_elect<false>(d_tmem, tmem_base + 8 * iw, b_lo_s + (uint32_t)(iw * 2), b_hi, idesc, leader);      // (src 0, kk = iw < 4)
#pragma unroll 4
            for (int idx = iw + NIW; idx < 32; idx += NIW) {
                const int src = idx >> 3, kk = idx & 7;
                umma_f16_ts_elect(d_tmem, tmem_base + 8 * idx,
                                  b_lo_s + (uint32_t)((src * SLICE + (kk >> 2) * (BG * 128)) / 16 + (kk & 3) * 2), b_hi, idesc, leader);
            }
            REC_PROBE(7);
            if (leader) {
                umma_commit(mbar_mma);
                if (paired) mbar_arrive(&mbar_turn[sg ^ 1]);
            }
            // the reduce buffer the NEXT step's partials go to: its previous contents were consumed in phase A of
            // this step (our own dz slice, part of the operand awaited above, was staged after reading them), and nobody can
            // send step s+1's partials before our dz of step s+1 exists
            if (leader && iw == 0 && s + 1 < S && S0 + s + 2 < T) mbar_arrive_expect_tx(&mbar_red[(s + 1) & 1], 4u * PT);
            REC_PROBE(2);
        }
        __syncwarp();
    } else if (role == 0) {
        // ============================ compute warps ============================
        const int ub = rw >> 2, q = rw & 3;
        const int up = lane >> 2, g = lane & 3;
        const int ul = q * 8 + up;                          // local unit (0..31)
        const int unit = (int)cta * 32 + ul;
        float wf = 0.f, wi = 0.f, wo = 0.f;
        if (p.peep) {
            wf = p.peep[(size_t)(dir * 3 + 0) * Hp + unit];
            wi = p.peep[(size_t)(dir * 3 + 1) * Hp + unit];
            wo = p.peep[(size_t)(dir * 3 + 2) * Hp + unit];
        }
        int len_j[2];
        float dcc[2];
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            const int b = b0 + ub * 8 + 2 * g + j;
            len_j[j] = (b < B) ? p.lens[b] : 0;
            dcc[j] = 0.f;
        }
        // a launch that resumes (S0 > 0) takes the recurrent dm of its first step and the carried dc from the previous launch
        float dm_in[2] = {0.f, 0.f};
        if (S0 > 0) {
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                const int b = b0 + ub * 8 + 2 * g + j;
                if (b < B) {
                    const float2 cv = *reinterpret_cast<const float2*>(p.carry + (((size_t)b * 2 + dir) * Hp + unit) * 2);
                    dm_in[j] = cv.x; dcc[j] = cv.y;
                }
            }
        }
        float db[4] = {0.f, 0.f, 0.f, 0.f};
        float dpf = 0.f, dpi = 0.f, dpo = 0.f;

        // raw prefetch of the next step's saved activations (loads only, stay in flight behind the current step)
        struct Pre { uint2 gp[2]; float c[2], cp[2], dmo[2]; };
        const long long row_stride = (dir ? 1 : -1) * (long long)B * (long long)ld2;      // elements per time step, in scan order
        // running pointers of utterance j = 0 at the step to be prefetched next (utterance j = 1 is the next row, + ld2 elements; a
        // scan step moves them by one frame); predicated loads, no branches: the prefetch used to cost ~90 instructions and two
        // divergent branches per step, issued in front of the wait for the partial tiles
        const size_t pidx0 = ((size_t)(dir ? S0 : T - 1 - S0) * B + (size_t)(b0 + ub * 8 + 2 * g)) * ld2 + (size_t)dir * Hp + unit;
        const uint2* gq = p.gates + pidx0;
        const float* cq = p.cst + pidx0;
        const float* dq = p.dM + pidx0;
        auto load_pre = [&](int s, Pre& r) {              // s: GLOBAL scan step (the pointers already stand on it); advances them
            const int t = dir ? s : (T - 1 - s);
            const int tp = dir ? (t + 1) : (t - 1);
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                const bool live = (s < T) && (t < len_j[j]);           // (a padding utterance has len 0: never live, never read)
                const bool has_prev = live && (dir ? (tp < len_j[j]) : (tp >= 0));
                const size_t o = j ? ld2 : 0;
                r.gp[j] = ldg_u2_if(gq + o, live);
                r.c[j] = ldg_f32_if(cq + o, live);
                r.dmo[j] = ldg_f32_if(dq + o, live);
                r.cp[j] = ldg_f32_if(cq + o + row_stride, has_prev);
            }
            gq += row_stride; cq += row_stride; dq += row_stride;
        };
        Pre cur, nxt;
        load_pre(S0, cur);
        // reduce-buffer read offset of this thread: unit row ul, 16-byte chunk ub (swizzled), utterance pair g
        const uint32_t rd_off = (uint32_t)(ul * (2 * BG) + ((ub ^ ((ul / SWS) & (NCH - 1))) << 4) + 4 * g);
        // dz staging offsets (own slice, SW128 K-major: row = utterance, 64 gate rows per 128-byte row) and global columns
        uint32_t bp_off[2][4];
#pragma unroll
        for (int gate = 0; gate < 4; ++gate) {
            const int kp = packed_col(ul, gate);
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                const int bl = ub * 8 + 2 * g + j;
                bp_off[j][gate] = (uint32_t)((kp >> 6) * (BG * 128) + (bl >> 3) * 1024 + (bl & 7) * 128 +
                                             ((((kp & 63) >> 3) ^ (bl & 7)) << 4) + (kp & 7) * 2);
            }
        }
        const uint32_t red_addr = smem_u32(red), stg_addr = smem_u32(Stg), pst_addr = smem_u32(pst);
        const uint32_t acc_addr = tmem_base + ((uint32_t)(q * 32) << 16) + REC_TMEM_ACC + sg * NIW * BG + ub * 8;

        auto reduce_partials = [&](int sp, float (&dmr)[2]) {          // sum of the four partial tiles sent in (local) step sp
            const uint32_t rb = red_addr + (uint32_t)((sp & 1) * 4 * PT) + rd_off;
            uint32_t v[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) v[k] = lds_b32(rb + (uint32_t)(k * PT));
            const float2 f0 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&v[0])), f1 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&v[1])),
                         f2 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&v[2])), f3 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&v[3]));
            dmr[0] = (f0.x + f1.x) + (f2.x + f3.x);
            dmr[1] = (f0.y + f1.y) + (f2.y + f3.y);
        };
        for (int s = 0; s < S; ++s) {
            REC_PROBE(8);
            load_pre(S0 + s + 1, nxt);
            REC_PROBE(9);
            // ---- phase A: dm_rec = sum of the four partial tiles of the previous step, then dz_t ----
            float dmr[2] = {dm_in[0], dm_in[1]};
            if (s > 0) {
                if (ok) ok = mbar_wait(&mbar_red[(s - 1) & 1], (uint32_t)(((s - 1) >> 1) & 1));   // async-proxy deliveries + complete_tx: a CTA-scope wait suffices
                REC_PROBE(10);
                reduce_partials(s - 1, dmr);
            }
            const uint32_t stg = stg_addr + (uint32_t)((s & 1) * SLICE);
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                // branch-free: for a frame past the utterance end (or a padding utterance) the prefetch delivered zeros for
                // the saved gates / cell state / dM, which makes every product below an exact 0 and resets the carried
                // dc through fg = 0 -- so both cells' dependency chains interleave instead of running back to back
                const float2 g01 = __half22float2(*reinterpret_cast<const __half2*>(&cur.gp[j].x));
                const float2 g23 = __half22float2(*reinterpret_cast<const __half2*>(&cur.gp[j].y));
                const float ig = g01.x, jt = g01.y, fg = g23.x, og = g23.y, cp = cur.cp[j];
                const float tc = tanh_ex2(cur.c[j]);                   // same formula as the forward pass; tanh(0) == 0 exactly
                const float dm = cur.dmo[j] + dmr[j];
                const float dzo = dm * tc * og * (1.f - og);
                const float dc = dcc[j] + dm * og * (1.f - tc * tc) + dzo * wo;
                const float dzf = dc * cp * fg * (1.f - fg);
                const float dzi = dc * jt * ig * (1.f - ig);
                const float dzj = dc * ig * (1.f - jt * jt);
                dcc[j] = dc * fg + dzi * wi + dzf * wf;
                dpi += dzi * cp; dpf += dzf * cp; dpo += dzo * cur.c[j];
                db[0] += dzi; db[1] += dzj; db[2] += dzf; db[3] += dzo;
                const float dzg[4] = {dzi, dzj, dzf, dzo};
#pragma unroll
                for (int gate = 0; gate < 4; ++gate) {
                    const __nv_bfloat16 v = __float2bfloat16(dzg[gate]);
                    sts_b16(stg + bp_off[j][gate], *reinterpret_cast<const uint16_t*>(&v));   // own slice of dz_t = MMA B operand = dG tile
                }
            }
            REC_PROBE(11);
            fence_proxy_async_smem();                      // generic-proxy stores -> visible to the bulk (async-proxy) stores
            __syncwarp();
            if (lane == 0) mbar_arrive(&mbar_slice[s & 1]);
            REC_PROBE(12);
            if (S0 + s + 1 < T) {
                // ---- phase B: quarter q of the partial dm tile -> owner CTA 4 mr + q (one bulk DSMEM copy per quarter) ----
                if (ok) ok = mbar_wait(mbar_mma, (uint32_t)(s & 1));
                REC_PROBE(13);
                tc_fence_after();
                uint32_t a[8], a2[8];
                tmem_ld_32x32b_x8(acc_addr, a);            // unit row = lane of the quarter, utterances ub*8 .. ub*8+7
                if (NIW > 1) tmem_ld_32x32b_x8(acc_addr + BG, a2);      // the second issuer's partial sum
                tmem_ld_wait();
                if (NIW > 1) {
#pragma unroll
                    for (int i = 0; i < 8; ++i) a[i] = __float_as_uint(__uint_as_float(a[i]) + __uint_as_float(a2[i]));
                }
                // staged as [32 units][BG utts] bf16, 16-byte chunks XOR-swizzled by the unit row (conflict-free here and
                // for the owner's reads)
                const uint32_t pw = pst_addr + (uint32_t)(((s & 1) * 4 + q) * PT);
                sts_v4(pw + lane * (2 * BG) + ((ub ^ ((lane / SWS) & (NCH - 1))) << 4),
                       pack_bf16x2(__uint_as_float(a[0]), __uint_as_float(a[1])), pack_bf16x2(__uint_as_float(a[2]), __uint_as_float(a[3])),
                       pack_bf16x2(__uint_as_float(a[4]), __uint_as_float(a[5])), pack_bf16x2(__uint_as_float(a[6]), __uint_as_float(a[7])));
                asm volatile("bar.sync %0, %1;" ::"r"(1 + sg * 4 + q), "n"(NUB * 32) : "memory");   // the NUB warps of quarter q of this sub-group
                // (one bulk copy per quarter: 512 16-byte st.async per step were measured ~450 cycles slower than this)
                if (ub == 0 && lane == 0) {
                    const uint32_t owner = (uint32_t)(4 * mr + q);
                    fence_proxy_async_smem();
                    bulk_copy_s2c(mapa_shared(red_addr + (uint32_t)(((s & 1) * 4 + kc) * PT), owner),
                                  pw, (uint32_t)PT, mapa_shared(smem_u32(&mbar_red[s & 1]), owner));
                }
                REC_PROBE(14);
                // the accumulators are read: the next pass may overwrite them (each issuer's first MMA of a step does not
                // accumulate, so there is no zero fill)
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(mbar_acc);
            }
            cur = nxt;
        }
        // ---- hand-over to the launch that continues at scan step S0 + S: recurrent dm of its first step, carried dc ----
        if (S0 + S < T) {
            float dmr[2] = {0.f, 0.f};
            if (ok) ok = mbar_wait(&mbar_red[(S - 1) & 1], (uint32_t)(((S - 1) >> 1) & 1));
            reduce_partials(S - 1, dmr);
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                const int b = b0 + ub * 8 + 2 * g + j;
                if (b < B) *reinterpret_cast<float2*>(p.carry + (((size_t)b * 2 + dir) * Hp + unit) * 2) = make_float2(dmr[j], dcc[j]);
            }
        }
        // ---- parameter gradients held in registers: reduce the 4 lanes of a unit, then atomics ----
#pragma unroll
        for (int o = 1; o <= 2; o <<= 1) {
#pragma unroll
            for (int k = 0; k < 4; ++k) db[k] += __shfl_xor_sync(0xffffffffu, db[k], o);
            dpf += __shfl_xor_sync(0xffffffffu, dpf, o);
            dpi += __shfl_xor_sync(0xffffffffu, dpi, o);
            dpo += __shfl_xor_sync(0xffffffffu, dpo, o);
        }
        if (g == 0) {
#pragma unroll
            for (int k = 0; k < 4; ++k) atomicAdd(p.dbias + (size_t)dir * 4 * Hp + packed_col(unit, k), db[k]);
            if (p.dpeep) {
                atomicAdd(p.dpeep + (size_t)(dir * 3 + 0) * Hp + unit, dpf);
                atomicAdd(p.dpeep + (size_t)(dir * 3 + 1) * Hp + unit, dpi);
                atomicAdd(p.dpeep + (size_t)(dir * 3 + 2) * Hp + unit, dpo);
            }
        }
    }
    tc_fence_before();
    cluster_sync_all();
    if (warp == NSG * NCW) { tc_fence_after(); tmem_dealloc<512>(tmem_base); }
}

// =================================================================================================
// host side
// =================================================================================================
static bool rec_plan(int Hp, int& nc) {
    if (Hp < 64 || (Hp & 63) || Hp > 512) return false;
    nc = Hp / 32;                                   // 32 units (128 gate rows = one MMA M tile) per CTA
    return RecFwd2Cfg<16, 2>::smem_bytes(Hp / 64) <= 232448 && RecBwd2Cfg<32>::smem_bytes(nc) <= 232448;
}

template <typename K, typename... Args>
static int launch_cluster(K kern, int grid, int threads, size_t smem, int nc, cudaStream_t st, const Args&... params) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return LCB_ERR_CUDA;
    if (nc > 8 && cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) != cudaSuccess) return LCB_ERR_CUDA;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)grid, 1, 1);
    cfg.blockDim = dim3((unsigned)threads, 1, 1);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = (unsigned)nc;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    g_launches += 1;
    cudaError_t e = cudaLaunchKernelEx(&cfg, kern, params...);
    if (e != cudaSuccess) { cudaGetLastError(); return LCB_ERR_CUDA; }
    return LCB_OK;
}

template <typename K>
static int max_clusters(K kern, int threads, size_t smem, int nc) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(nc * 8), 1, 1);
    cfg.blockDim = dim3((unsigned)threads, 1, 1);
    cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = (unsigned)nc; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (nc > 8) cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    int n = 0;
    if (cudaOccupancyMaxActiveClusters(&n, kern, &cfg) != cudaSuccess) { cudaGetLastError(); return -1; }
    return n;
}

// clusters of one 16-utterance group each that the device keeps resident at once (queried once per cluster size)
static int resident_clusters(int nc, int which) {
    static int cap[3][17];
    if (cap[which][nc] == 0) {
        int n = which == 2 ? max_clusters(lstm_rec_fwd2_kernel<16, 2>, RecFwd2Cfg<16, 2>::THREADS, RecFwd2Cfg<16, 2>::smem_bytes(nc * 32 / 64), nc)
              : which ? max_clusters(lstm_rec_bwd2_kernel<16>, RecBwd2Cfg<16>::THREADS, RecBwd2Cfg<16>::smem_bytes(nc), nc)
                      : max_clusters(lstm_rec_fwd2_kernel<16, 1>, RecFwd2Cfg<16>::THREADS, RecFwd2Cfg<16>::smem_bytes(nc * 32 / 64), nc);
        cap[which][nc] = n > 0 ? n : 1;
    }
    return cap[which][nc];
}

// Forward layout: ncd clusters per direction, the first n_paired of them hold two 16-utterance groups (stepping independently,
// sharing the resident weights: ~1.24 us per scan step), the others one group alone (~1.08 us).  As many clusters as the device
// keeps resident are used (7 of 16 CTAs on a B200: 3 per direction); with length-sorted batches the paired clusters get the
// lowest group indices = the shortest utterances, whose sub-groups retire early (see the kernel), so the cluster that carries
// two groups and the one that carries the longest group alone finish about together (profiles/r02_rec_fwd_layout.txt).
static int g_fwd_layout_mode = 0;         // debug (lcb_debug_fwd_layout): 1 = pair everything as the round-1 kernels did
static void fwd_layout(int B, int nc, int ndir, int& ncd, int& n_paired) {
    const int groups = (B + 15) / 16;
    if (g_fwd_layout_mode == 1 && ndir * groups > resident_clusters(nc, 0)) { ncd = (groups + 1) / 2; n_paired = groups - ncd; return; }
    const int cap = resident_clusters(nc, 0) / ndir > 0 ? resident_clusters(nc, 0) / ndir : 1;   // resident clusters per direction
    if (groups <= cap) { ncd = groups; n_paired = 0; }
    else if (groups <= 2 * cap) { ncd = cap; n_paired = groups - cap; }
    else { ncd = (groups + 1) / 2; n_paired = groups - ncd; }                                   // several waves: pair everything
}

// Utterances per cluster: 16 while every 16-utterance group gets its own resident cluster (shortest per-step chain), 32
// otherwise -- as two independent 16-utterance sub-groups sharing the cluster's resident weights where that kernel exists
// (forward: always; BPTT: the 4 x 4 kernel of Hp = 512), else as one lockstep group of 32.
static int choose_bg(int B, int nc, int ndir, int which) {
    const int groups = (B + 15) / 16;
    return (ndir * groups <= resident_clusters(nc, which)) ? 16 : 32;
}

}  // namespace lcb

using namespace lcb;

// debug: device buffer of steps*16 int64 clock samples filled by the next lcb_lstm_rec_fwd launch (nullptr disables)
extern "C" int lcb_debug_rec_profile(long long* buf, int steps)
{
    if (cudaMemcpyToSymbol(lcb::g_rec_prof, &buf, sizeof(buf)) != cudaSuccess) return LCB_ERR_CUDA;
    if (cudaMemcpyToSymbol(lcb::g_rec_prof_steps, &steps, sizeof(steps)) != cudaSuccess) return LCB_ERR_CUDA;
    return LCB_OK;
}

// one-word store on a stream (the "frames projected so far" counter of lcb_lstm_rec_fwd_range_hl)
__global__ void store_i32_kernel(int* dst, int v) { *dst = v; __threadfence(); }
extern "C" int lcb_store_i32(int32_t* dst, int32_t value, void* stream)
{
    if (!dst) return LCB_ERR_NULL_POINTER;
    store_i32_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(dst, value);
    lcb::g_launches += 1;
    return cudaGetLastError() == cudaSuccess ? LCB_OK : LCB_ERR_CUDA;
}

// debug: forward cluster layout override (0 = automatic, 1 = two groups in every cluster whenever one group per cluster does not fit)
extern "C" int lcb_debug_fwd_layout(int mode) { lcb::g_fwd_layout_mode = mode; return LCB_OK; }

// how many clusters of the forward (which = 0) / backward (which = 1) kernel (one 16-utterance group per cluster) the
// device can keep resident at once
extern "C" int lcb_lstm_rec_max_clusters(int Hp, int which)
{
    int nc;
    if (!rec_plan(Hp, nc)) return LCB_ERR_UNSUPPORTED;
    return resident_clusters(nc, which == 2 ? 2 : (which ? 1 : 0));        // 2: the forward kernel with two sub-groups per cluster
}

extern "C" int lcb_lstm_rec_config(int Hp, int* units_per_cta_div32, int* cluster_size)
{
    int nc;
    if (!rec_plan(Hp, nc)) return LCB_ERR_UNSUPPORTED;
    if (units_per_cta_div32) *units_per_cta_div32 = 1;
    if (cluster_size) *cluster_size = nc;
    return LCB_OK;
}

// SMs (= CTAs, one per SM) a forward (which = 0) / BPTT (which = 1) launch occupies: what a GEMM overlapped with it on another
// stream must leave free (its max_ctas = lcb_device_sm_count() - this)
extern "C" int lcb_lstm_rec_grid(int B, int Hp, int num_dirs, int which)
{
    int nc;
    if (!rec_plan(Hp, nc) || B <= 0 || num_dirs < 1 || num_dirs > 2) return LCB_ERR_UNSUPPORTED;
    if (!which) {
        int ncd, npair;
        fwd_layout(B, nc, num_dirs, ncd, npair);
        return num_dirs * ncd * nc;
    }
    const int bgs = choose_bg(B, nc, num_dirs, 1);
    return num_dirs * ((B + bgs - 1) / bgs) * nc;
}

// bytes of the L2 exchange scratch the kernels use (per launch; must not be shared by concurrently running launches)
extern "C" size_t lcb_lstm_rec_workspace_bytes(int B, int Hp)
{
    int nc;
    if (!rec_plan(Hp, nc) || B <= 0) return 0;
    const size_t clusters = 2 * (size_t)((B + 15) / 16);
    return clusters * 2 * 16 * 8192;               // [cluster][2 step parities][16 CTAs][<= 8 KB slice (BPTT dz slice, 32 utterances)]
}

extern "C" int lcb_lstm_rec_fwd(const float* G, const void* WfoldT, const float* peep, const int32_t* lens,
                                void* Mout, void* gates, float* cst, float* cfin, float* mfin,
                                int T, int B, int Hp, int num_dirs, float forget_bias,
                                void* workspace, size_t workspace_bytes, void* stream)
{
    return lcb_lstm_rec_fwd_range(G, WfoldT, peep, lens, Mout, gates, cst, cfin, mfin, T, B, Hp, num_dirs, forget_bias, 0, T,
                                  workspace, workspace_bytes, stream);
}

extern "C" int lcb_lstm_rec_fwd_range(const float* G, const void* WfoldT, const float* peep, const int32_t* lens,
                                      void* Mout, void* gates, float* cst, float* cfin, float* mfin,
                                      int T, int B, int Hp, int num_dirs, float forget_bias, int s_begin, int s_end,
                                      void* workspace, size_t workspace_bytes, void* stream)
{
    return lcb_lstm_rec_fwd_range_hl(G, WfoldT, peep, lens, nullptr, nullptr, Mout, gates, cst, cfin, mfin, T, B, Hp, num_dirs, forget_bias,
                                     s_begin, s_end, workspace, workspace_bytes, stream);
}

extern "C" int lcb_lstm_rec_fwd_range_hl(const float* G, const void* WfoldT, const float* peep, const int32_t* lens,
                                         const int32_t* lens_host, const int32_t* ready_steps,
                                         void* Mout, void* gates, float* cst, float* cfin, float* mfin,
                                         int T, int B, int Hp, int num_dirs, float forget_bias, int s_begin, int s_end,
                                         void* workspace, size_t workspace_bytes, void* stream)
{
    return lcb_lstm_rec_fwd_range_pg(G, 0, WfoldT, peep, lens, lens_host, ready_steps, Mout, nullptr, gates, cst, cfin, mfin, T, B, Hp, num_dirs,
                                     forget_bias, s_begin, s_end, nullptr, workspace, workspace_bytes, stream);
}

// words of the progress array of lcb_lstm_rec_fwd_range_pg for a batch of B utterances
extern "C" int lcb_lstm_rec_fwd_progress_words(int B, int Hp, int num_dirs)
{
    int nc;
    if (!rec_plan(Hp, nc) || B <= 0 || num_dirs < 1 || num_dirs > 2) return 0;
    int ncd, npair;
    fwd_layout(B, nc, num_dirs, ncd, npair);
    return num_dirs * ncd * (npair > 0 ? 2 : 1) * nc;
}

// The same launch publishing its progress: word (cluster, sub-group, CTA) of `progress` (zeroed by the caller,
// lcb_lstm_rec_fwd_progress_words words) counts the leading scan steps whose Mout rows that CTA has written -- advanced every 16 steps
// and set to s_end when the sub-group is done.  lcb_wait_progress on another stream releases the output-projection GEMMs of the
// finished frames while the recurrence is still running; the launch itself never waits for it.
extern "C" int lcb_lstm_rec_fwd_range_pg(const void* G, int g_dtype, const void* WfoldT, const float* peep, const int32_t* lens,
                                         const int32_t* lens_host, const int32_t* ready_steps,
                                         void* Mout, void* Mout_bf16, void* gates, float* cst, float* cfin, float* mfin,
                                         int T, int B, int Hp, int num_dirs, float forget_bias, int s_begin, int s_end,
                                         int32_t* progress, void* workspace, size_t workspace_bytes, void* stream)
{
    if (g_dtype != 0 && g_dtype != 2) return LCB_ERR_UNSUPPORTED;          // 0: fp32, 2: fp16 (dtype codes of lcb_gemm16)
    const bool gh = g_dtype == 2;
    if (!G || !WfoldT || !lens || !Mout || !workspace) return LCB_ERR_NULL_POINTER;
    if (T <= 0 || B <= 0 || num_dirs < 1 || num_dirs > 2) return LCB_ERR_BAD_SHAPE;
    if (s_begin < 0 || s_end > T || s_begin >= s_end) return LCB_ERR_BAD_SHAPE;
    if (s_begin > 0 && !cst) return LCB_ERR_NULL_POINTER;       // resuming reads the saved cell state
    if ((cfin == nullptr) != (mfin == nullptr) || (gates == nullptr) != (cst == nullptr)) return LCB_ERR_NULL_POINTER;
    int nc;
    if (!rec_plan(Hp, nc)) return LCB_ERR_UNSUPPORTED;
    if (((uintptr_t)G & 15) || ((uintptr_t)WfoldT & 15) || ((uintptr_t)workspace & 15)) return LCB_ERR_MISALIGNED;
    if (workspace_bytes < lcb_lstm_rec_workspace_bytes(B, Hp)) return LCB_ERR_WORKSPACE_TOO_SMALL;
    RecFwdParams p;
    p.G = (const float*)G; p.Wt = (const __half*)WfoldT; p.peep = peep; p.lens = lens; p.Mout = (__half*)Mout;
    p.gates = (uint2*)gates; p.cst = cst; p.cfin = cfin; p.mfin = mfin;
    p.Mbf = (__nv_bfloat16*)Mout_bf16;
    p.T = T; p.B = B; p.Hp = Hp; p.NC = nc; p.ndir = num_dirs; p.forget_bias = forget_bias; p.s_begin = s_begin; p.s_end = s_end;
    p.xch = (unsigned char*)workspace;
    int ncd, npair;
    fwd_layout(B, nc, num_dirs, ncd, npair);
    p.n_paired = npair;
    p.ready = ready_steps;
    p.progress = progress;
    for (int g = 0; g < REC_MAXGRP; ++g) {
        int m = T;
        if (lens_host && g * 16 < B) {
            m = 0;
            for (int u = g * 16; u < B && u < g * 16 + 16; ++u) m = lens_host[u] > m ? lens_host[u] : m;
        }
        p.gmax[g] = m;
    }
    const int ncl = num_dirs * ncd;
    cudaStream_t st = (cudaStream_t)stream;
    // G as a 3-D tensor [T][B][8Hp] fp32, box = 128 packed gate columns x 16 utterances of one frame (dense, no swizzle)
    CUtensorMap tm;
    const uint64_t ge = gh ? 2 : 4;          // (a 16-bit tensor map moves fp16 as it moves bf16)
    if (!make_tmap_3d(&tm, !gh, G, (uint64_t)8 * Hp, (uint64_t)B, (uint64_t)T, (uint64_t)8 * Hp * ge, (uint64_t)B * 8 * Hp * ge,
                      128, 16u, 1, false)) return LCB_ERR_CUDA;
    if (npair == 0)
        return gh ? launch_cluster(lstm_rec_fwd2_kernel<16, 1, true>, ncl * nc, RecFwd2Cfg<16>::THREADS, RecFwd2Cfg<16>::smem_bytes(Hp / 64), nc, st, p, tm)
                  : launch_cluster(lstm_rec_fwd2_kernel<16, 1, false>, ncl * nc, RecFwd2Cfg<16>::THREADS, RecFwd2Cfg<16>::smem_bytes(Hp / 64), nc, st, p, tm);
    // too many 16-utterance groups for one wave of clusters: two of them in some (or all) clusters, stepping independently
    return gh ? launch_cluster(lstm_rec_fwd2_kernel<16, 2, true>, ncl * nc, RecFwd2Cfg<16, 2>::THREADS, RecFwd2Cfg<16, 2>::smem_bytes(Hp / 64), nc, st, p, tm)
              : launch_cluster(lstm_rec_fwd2_kernel<16, 2, false>, ncl * nc, RecFwd2Cfg<16, 2>::THREADS, RecFwd2Cfg<16, 2>::smem_bytes(Hp / 64), nc, st, p, tm);
}

extern "C" int lcb_lstm_rec_bwd(const float* dM, const void* gates, const float* cst, const void* Wfold, const float* peep,
                                const int32_t* lens, void* dG, float* dbias, float* dpeep,
                                int T, int B, int Hp, int num_dirs, void* workspace, size_t workspace_bytes, void* stream)
{
    return lcb_lstm_rec_bwd_range(dM, gates, cst, Wfold, peep, lens, dG, dbias, dpeep, T, B, Hp, num_dirs, 0, T, nullptr,
                                  workspace, workspace_bytes, stream);
}

// words of the progress array of lcb_lstm_rec_bwd_range_pg for a batch of B utterances (0: this cell size has no progress output)
extern "C" int lcb_lstm_rec_bwd_progress_words(int B, int Hp, int num_dirs)
{
    int nc;
    if (!rec_plan(Hp, nc) || B <= 0 || Hp != 512) return 0;
    const int bgs0 = choose_bg(B, nc, num_dirs, 1);
    return num_dirs * ((B + bgs0 - 1) / bgs0) * (bgs0 == 32 ? 2 : 1) * nc;
}

// Spin (one thread per word, nanosleep back-off, bounded) until every word of `progress` has reached `target`: enqueued on the
// stream of the GEMMs that consume the dG rows a running lcb_lstm_rec_bwd_range_pg launch has released.  The BPTT launch never
// waits for anything here, so this cannot deadlock -- serialised, it finds the final values.
__global__ void wait_progress_kernel(const int* progress, int n, int target)
{
    const int i = threadIdx.x + blockIdx.x * blockDim.x;
    if (i >= n) return;
    uint64_t t0 = 0; uint32_t spins = 0;
    while (ld_acquire_gpu_s32(progress + i) < target) {
        __nanosleep(200);
        if ((++spins & 0xff) == 0) {
            const uint64_t now = globaltimer_ns();
            if (t0 == 0) t0 = now;
            if (now - t0 > LCB_WAIT_TIMEOUT_NS || dev_has_error()) { dev_set_error(DEV_ERR_MBAR_TIMEOUT); return; }
        }
    }
}
extern "C" int lcb_wait_progress(const int32_t* progress, int n, int target, void* stream)
{
    if (!progress || n <= 0) return LCB_ERR_NULL_POINTER;
    wait_progress_kernel<<<(n + 127) / 128, 128, 0, (cudaStream_t)stream>>>(progress, n, target);
    lcb::g_launches += 1;
    return cudaGetLastError() == cudaSuccess ? LCB_OK : LCB_ERR_CUDA;
}

// 1 when lcb_lstm_rec_bwd_range accepts partial ranges for this cell size (the 4 x 4 kernel, Hp = 512)
extern "C" int lcb_lstm_rec_bwd_can_split(int Hp)
{
    return Hp == 512 ? 1 : 0;
}

extern "C" int lcb_lstm_rec_bwd_range(const float* dM, const void* gates, const float* cst, const void* Wfold, const float* peep,
                                      const int32_t* lens, void* dG, float* dbias, float* dpeep,
                                      int T, int B, int Hp, int num_dirs, int s_begin, int s_end, float* carry,
                                      void* workspace, size_t workspace_bytes, void* stream)
{
    return lcb_lstm_rec_bwd_range_pg(dM, gates, cst, Wfold, peep, lens, dG, dbias, dpeep, T, B, Hp, num_dirs, s_begin, s_end, carry,
                                     nullptr, workspace, workspace_bytes, stream);
}

extern "C" int lcb_lstm_rec_bwd_range_pg(const float* dM, const void* gates, const float* cst, const void* Wfold, const float* peep,
                                         const int32_t* lens, void* dG, float* dbias, float* dpeep,
                                         int T, int B, int Hp, int num_dirs, int s_begin, int s_end, float* carry, int32_t* progress,
                                         void* workspace, size_t workspace_bytes, void* stream)
{
    if (progress && Hp != 512) return LCB_ERR_UNSUPPORTED;
    if (!dM || !gates || !cst || !Wfold || !lens || !dG || !dbias || !workspace) return LCB_ERR_NULL_POINTER;
    if (T <= 0 || B <= 0 || num_dirs < 1 || num_dirs > 2) return LCB_ERR_BAD_SHAPE;
    if (s_begin < 0 || s_end > T || s_begin >= s_end) return LCB_ERR_BAD_SHAPE;
    const bool whole = s_begin == 0 && s_end == T;
    if (!whole && !carry) return LCB_ERR_NULL_POINTER;
    if (carry && ((uintptr_t)carry & 7)) return LCB_ERR_MISALIGNED;
    if ((peep == nullptr) != (dpeep == nullptr)) return LCB_ERR_NULL_POINTER;
    int nc;
    if (!rec_plan(Hp, nc)) return LCB_ERR_UNSUPPORTED;
    if (((uintptr_t)Wfold & 15) || ((uintptr_t)workspace & 15)) return LCB_ERR_MISALIGNED;
    if (workspace_bytes < lcb_lstm_rec_workspace_bytes(B, Hp)) return LCB_ERR_WORKSPACE_TOO_SMALL;
    if (!whole && !lcb_lstm_rec_bwd_can_split(Hp)) return LCB_ERR_UNSUPPORTED;
    RecBwdParams p;
    p.dM = dM; p.gates = (const uint2*)gates; p.cst = cst; p.W = (const __nv_bfloat16*)Wfold; p.peep = peep; p.lens = lens;
    p.dG = (__nv_bfloat16*)dG; p.dbias = dbias; p.dpeep = dpeep;
    p.T = T; p.B = B; p.Hp = Hp; p.NC = nc; p.ndir = num_dirs; p.s_begin = s_begin; p.s_end = s_end; p.carry = carry;
    p.progress = progress;
    p.xch = (unsigned char*)workspace;
    cudaStream_t st = (cudaStream_t)stream;
    const int bgs0 = choose_bg(B, nc, num_dirs, 1);
    const int ncl = num_dirs * ((B + bgs0 - 1) / bgs0);
    if (Hp == 512) {
        // two 16-utterance sub-groups per cluster instead of one group of 32
        // dG as a 3-D tensor [T][B][8Hp] bf16, box = 64 columns x 16 utterances, 128B swizzle (= the MMA operand layout of a slice)
        CUtensorMap tm;
        if (((uintptr_t)dG & 15) || !make_tmap_3d(&tm, false, dG, (uint64_t)8 * Hp, (uint64_t)B, (uint64_t)T, (uint64_t)8 * Hp * 2,
                                                  (uint64_t)B * 8 * Hp * 2, 64, 16u, 1, true)) return LCB_ERR_CUDA;
        if (bgs0 == 32)
            return launch_cluster(lstm_rec_bwd3_kernel<16, 2>, ncl * nc, RecBwd3Cfg<16, 2>::THREADS, RecBwd3Cfg<16, 2>::smem_bytes(), nc, st, p, tm);
        return launch_cluster(lstm_rec_bwd3_kernel<16, 1>, ncl * nc, RecBwd3Cfg<16>::THREADS, RecBwd3Cfg<16>::smem_bytes(), nc, st, p, tm);
    }
    if (bgs0 == 16)
        return launch_cluster(lstm_rec_bwd2_kernel<16>, ncl * nc, RecBwd2Cfg<16>::THREADS, RecBwd2Cfg<16>::smem_bytes(nc), nc, st, p);
    return launch_cluster(lstm_rec_bwd2_kernel<32>, ncl * nc, RecBwd2Cfg<32>::THREADS, RecBwd2Cfg<32>::smem_bytes(nc), nc, st, p);
}
