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
// Per step: two issuer warps (K halves, own accumulators) issue D[128 gate rows, 16 utts] = W'^T_slice *
// m_{t-1}^T; 8 compute warps read their accumulator fragment with tcgen05.ld.16x256b (the packing above puts
// the four gates of a unit in ONE thread), add the prefetched G tile, apply gates / peepholes / cell update
// / length mask in registers (the fp32 cell state never leaves registers), push m_t (fp16) into the operand
// buffer of every CTA of the cluster with st.async (DSMEM, completion counted on the receiver's mbarrier --
// no barrier.cluster and no release fence in the loop), then write m_t and the saved activations to HBM off
// the critical path.  A loader warp keeps an 8-deep ring of G tiles in flight (one bulk copy per lane).
// The backward direction is the same scan in descending absolute time under the mask t < len[b]
// (state stays at its zero initial value until t = len[b]-1), so no reversed copies are ever made.
//
// BPTT runs the mirrored scan with W' (bf16) resident in TMEM as [units, own gate rows]: dz_t is formed in
// registers, staged locally as the MMA B operand, partial dm_{t-1} tiles are reduce-scattered to their owner
// CTAs with st.async (double-buffered reduce buffer, mbarrier tx-counted).
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

constexpr int REC_BG = 16;        // utterances per cluster
constexpr int REC_GROW = 132;     // padded fp32 row of a staged G tile: 2*132 = 8 (mod 32) -> conflict-free gate reads
constexpr int REC_SG = 8;         // G prefetch ring depth
constexpr int REC_TMEM_ACC = 256; // first accumulator column (weights occupy [0, Hp/2) forward, [0, 64*MB) backward)

struct RecFwdParams {
    const float* G;               // [T*B][8Hp] fp32  x_t * W_x + bias, packed gate columns
    const __half* Wt;             // [8Hp][Hp] fp16   (W_proj * W_h)^T, rows = packed gate columns of both directions
    const float* peep;            // [2][3][Hp]  (w_f, w_i, w_o) or nullptr
    const int* lens;              // [B]
    __half* Mout;                 // [T*B][2Hp]   m_t, fp16 (0 where t >= len)
    uint2* gates;                 // [T*B][2Hp]   saved (i, tanh j, f, o) as 4 x fp16   (nullptr: inference)
    float* cst;                   // [T*B][2Hp]   saved cell state c_t                   (nullptr: inference)
    float* cfin;                  // [B][2][Hp] final cell state   (nullable)
    float* mfin;                  // [B][2][Hp] final m (pre-projection output) (nullable)
    int T, B, Hp, NC;
    float forget_bias;
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
};

__device__ __forceinline__ unsigned char* align_1024(unsigned char* p) {
    return reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(p) + 1023) & ~(uintptr_t)1023);
}

// packed gate column of (unit, gate): gate-major inside each group of 8 units, so that one warp's
// tcgen05.ld.16x256b delivers all four gates of a unit to the same thread
__host__ __device__ __forceinline__ int packed_col(int unit, int gate) { return ((unit >> 3) << 5) + (gate << 3) + (unit & 7); }

// ---- forward kernel resources, per sub-group (NSG sub-groups of 16 utterances share one cluster and its weights) ----
constexpr int FWD_NCW = 8;                    // compute warps per sub-group: (column half, TMEM lane quarter)
constexpr int FWD_NIW = 2;                    // MMA issuer warps per sub-group (measured: 2 beat 4 -- the tensor pipe is the limit)
constexpr int FWD_BARS = 4 + 2 * REC_SG + 16; // mma[4] g[SG] gfree[SG] op[2][8]
template <int NSG> struct RecFwdCfg {
    static constexpr int THREADS = 32 * NSG * (FWD_NCW + FWD_NIW + 1);
    __host__ __device__ static size_t sg_bytes(int KB) {          // operand double buffer + G ring + slice staging, 1024-aligned
        size_t b = (size_t)2 * KB * REC_BG * 128 + (size_t)REC_SG * REC_BG * REC_GROW * 4 + 2 * 1024;
        return (b + 1023) & ~(size_t)1023;
    }
    static size_t smem_bytes(int KB) { return 1024 + NSG * sg_bytes(KB) + NSG * FWD_BARS * 8 + 64; }
};

// =================================================================================================
// forward
// =================================================================================================
template <int NSG>
__global__ void __launch_bounds__(32 * NSG * 11, 1)
lstm_rec_fwd_kernel(const RecFwdParams p)
{
    constexpr int BG = REC_BG, SG = REC_SG, NCW = FWD_NCW, NIW = FWD_NIW;
    using Cfg = RecFwdCfg<NSG>;
    extern __shared__ unsigned char smem_raw[];
    unsigned char* smem = align_1024(smem_raw);
    const int Hp = p.Hp, KB = Hp >> 6, NC = p.NC, T = p.T, B = p.B;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // warp roles: [0, 8 NSG) compute, [8 NSG, 10 NSG) MMA issuers, [10 NSG, 11 NSG) loaders
    int role, sg, rw;                                   // rw = warp index inside (role, sub-group)
    if (warp < NCW * NSG) { role = 0; sg = warp / NCW; rw = warp % NCW; }
    else if (warp < (NCW + NIW) * NSG) { role = 1; sg = (warp - NCW * NSG) / NIW; rw = (warp - NCW * NSG) % NIW; }
    else { role = 2; sg = warp - (NCW + NIW) * NSG; rw = 0; }

    unsigned char* sgbase = smem + (size_t)sg * Cfg::sg_bytes(KB);
    unsigned char* Bsm = sgbase;                                                 // [2][KB][16 x 128 B] operand m_{t-1}
    float* Gsm = reinterpret_cast<float*>(Bsm + (size_t)2 * KB * BG * 128);     // [SG][16][132]
    unsigned char* Msm = reinterpret_cast<unsigned char*>(Gsm + (size_t)SG * BG * REC_GROW);   // [2 step parities][1 KB slice]
    uint64_t* bars_all = reinterpret_cast<uint64_t*>(smem + (size_t)NSG * Cfg::sg_bytes(KB));
    uint64_t* bars = bars_all + (size_t)sg * FWD_BARS;
    uint64_t* mbar_mma = bars;                         // [4]  accumulator of issuer i complete
    uint64_t* mbar_g = bars + 4;                       // [SG] G tile landed (bulk-copy tx)
    uint64_t* mbar_gfree = bars + 4 + SG;              // [SG] compute warps are done with the G stage
    uint64_t* mbar_op = bars + 4 + 2 * SG;             // [2][8] per 64-unit K block of the operand: both source CTAs' slices landed
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars_all + (size_t)NSG * FWD_BARS);

    const uint32_t cta = cluster_ctarank();
    const int cid = (int)cluster_id_x();
    const int dir = cid & 1, bg = cid >> 1;
    const int b0 = (bg * NSG + sg) * BG;               // first utterance of this sub-group
    const bool sg_active = b0 < B;
    const size_t ld2 = (size_t)2 * Hp;

    if (threadIdx.x == 0) {
        for (int g2 = 0; g2 < NSG; ++g2) {
            uint64_t* bb = bars_all + (size_t)g2 * FWD_BARS;
            for (int i = 0; i < 4; ++i) mbar_init(&bb[i], 1);
            for (int s = 0; s < SG; ++s) { mbar_init(&bb[4 + s], 1); mbar_init(&bb[4 + SG + s], NCW); }
            for (int i = 0; i < 16; ++i) mbar_init(&bb[4 + 2 * SG + i], 1);
        }
        fence_mbar_init();
    }
    if (warp == NCW * NSG) tmem_alloc<512>(tmem_slot);
    {   // zero the operand buffers of every sub-group (m_{-1} = 0)
        for (int g2 = 0; g2 < NSG; ++g2) {
            uint4* bz = reinterpret_cast<uint4*>(smem + (size_t)g2 * Cfg::sg_bytes(KB));
            const int n16 = 2 * KB * BG * 128 / 16;
            for (int i = threadIdx.x; i < n16; i += blockDim.x) bz[i] = make_uint4(0u, 0u, 0u, 0u);
        }
    }
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    // ---- W'^T slice -> tensor memory: row r of the slice lives in TMEM lane r, 16 fp16 per 8 columns ----
    if (role == 0) {
        const int q = warp & 3;
        const __half* wrow = p.Wt + ((size_t)dir * 4 * Hp + (size_t)cta * 128 + q * 32 + lane) * Hp;
        for (int ch = warp >> 2; ch < Hp / 16; ch += 2 * NSG) {
            const uint4 v0 = __ldg(reinterpret_cast<const uint4*>(wrow + ch * 16));
            const uint4 v1 = __ldg(reinterpret_cast<const uint4*>(wrow + ch * 16 + 8));
            const uint32_t r[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
            tmem_st_32x32b_x8(tmem_base + ((uint32_t)(q * 32) << 16) + ch * 8, r);
        }
        tmem_st_wait();
    }
    tc_fence_before();
    cluster_sync_all();          // weights resident; every CTA of the cluster is initialised before any DSMEM traffic
    tc_fence_after();

    long long* prof = (blockIdx.x == 0 && lane == 0 && sg == 0 && ((role == 1 && rw == 0) || (role == 0 && rw == 0))) ? g_rec_prof : nullptr;
    const int prof_steps = g_rec_prof_steps;
    const int nvalid = (B - b0) < BG ? (B - b0) : BG;           // utterances of this sub-group that exist

    bool ok = true;
    if (!sg_active) {
        // nothing to do for this sub-group (batch smaller than the cluster's capacity)
    } else if (role == 2) {
        // ============================ loader warp: G tile prefetch, one 512-byte bulk copy per lane ============================
        for (int s = 0; s < T; ++s) {
            const int stage = s % SG;
            if (s >= SG) {              // wait until the compute warps drained this stage (step s - SG)
                if (lane == 0 && ok) ok = mbar_wait(&mbar_gfree[stage], (uint32_t)(((s - SG) / SG) & 1));
                ok = __shfl_sync(0xffffffffu, ok ? 1 : 0, 0) != 0;
                if (!ok) break;
            }
            const int t = dir ? (T - 1 - s) : s;
            if (lane == 0) mbar_arrive_expect_tx(&mbar_g[stage], (uint32_t)(nvalid * 512));
            __syncwarp();
            if (lane < nvalid)
                bulk_load_1d(Gsm + ((size_t)stage * BG + lane) * REC_GROW,
                             p.G + ((size_t)t * B + b0 + lane) * 8 * Hp + (size_t)dir * 4 * Hp + (size_t)cta * 128,
                             512, &mbar_g[stage]);
        }
    } else if (role == 1) {
        // ============================ MMA issuer warps ============================
        // forward operands are fp16 (|m| < 1, small weights): a_format = b_format = F16 (0); A from TMEM, K-major
        constexpr uint32_t idesc = make_idesc_bf16_f32(128, BG, 0, 0) & ~((7u << 7) | (7u << 10));
        // Issuer i owns the K blocks kb = i, i+NIW, ... (64 units each = the slices of source CTAs 2kb, 2kb+1) and
        // its own accumulator; it starts as soon as ITS next block has landed, so MMA issue overlaps the DSMEM
        // ingress of the other blocks.
        const int iw = rw;
        if (lane == 0 && iw < KB) {
            // operand m_{t-1}: no-swizzle K-major core matrices, [unit/8][utt/8] blocks of 128 B (8 utts x 8 units):
            // LBO (next 8 units) = 256 B, SBO (next 8 utterances) = 128 B; one K=16 MMA step = 512 B
            const uint64_t bb0 = make_smem_desc_noswz(smem_u32(Bsm), 256, 128);
            const uint32_t b_lo0 = (uint32_t)bb0, b_hi = (uint32_t)(bb0 >> 32);
            const uint32_t d_tmem = tmem_base + REC_TMEM_ACC + (sg * NIW + iw) * BG;
            for (int s = 0; s < T && ok; ++s) {
                REC_PROBE(0);
                const uint32_t par = (uint32_t)(s & 1);
                if (s + 1 < T)                        // arm the K-block barriers of the buffer that step s fills
                    for (int kb = iw; kb < KB; kb += NIW) mbar_arrive_expect_tx(&mbar_op[((s + 1) & 1) * 8 + kb], 2048u);
                const uint32_t b_lo_s = b_lo0 + (uint32_t)((par * KB * BG * 128) >> 4);
                uint32_t first = 0;
                for (int kb = iw; kb < KB; kb += NIW) {
                    if (s > 0) {
                        ok = mbar_wait_cluster_acq(&mbar_op[par * 8 + kb], (uint32_t)(((s - 1) >> 1) & 1));
                        if (!ok) break;
                        fence_proxy_async_smem();    // DSMEM-delivered operand -> visible to the tensor core (async proxy)
                    }
                    if (kb == iw) { REC_PROBE(1); }
                    tc_fence_after();
                    const uint32_t at = tmem_base + (uint32_t)(kb * 32), bl = b_lo_s + (uint32_t)(kb * (4 * 512 / 16));
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        umma_f16_ts_lohi(d_tmem, at + 8 * k, bl + (512 / 16) * k, b_hi, idesc, first);
                        first = 1u;
                    }
                }
                if (!ok) break;
                REC_PROBE(7);
                umma_commit(&mbar_mma[iw]);
                REC_PROBE(2);
            }
        }
        __syncwarp();
    } else {
        // ============================ compute warps: (column half, TMEM lane quarter) ============================
        const int cg = rw >> 2, q = rw & 3;                    // q == warp % 4: the TMEM lane quarter this warp may access
        const int up = lane >> 2, g = lane & 3;
        const int unit0 = (int)cta * 32 + q * 8;               // first of the 8 units of this warp
        const int unit = unit0 + up;
        float wf = 0.f, wi = 0.f, wo = 0.f;
        if (p.peep) {
            wf = p.peep[(size_t)(dir * 3 + 0) * Hp + unit];
            wi = p.peep[(size_t)(dir * 3 + 1) * Hp + unit];
            wo = p.peep[(size_t)(dir * 3 + 2) * Hp + unit];
        }
        // this thread's two utterances: local rows cg*8 + 2g + j
        int len_j[2];
        float c_reg[2];
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            const int b = b0 + cg * 8 + 2 * g + j;
            len_j[j] = (b < B) ? p.lens[b] : 0;
            c_reg[j] = 0.f;
        }
        const uint32_t b_addr = smem_u32(Bsm);
        const uint32_t t_addr = tmem_base + ((uint32_t)(q * 32) << 16) + REC_TMEM_ACC + sg * NIW * BG + cg * 8;
        const int nacc = KB < NIW ? KB : NIW;                 // accumulators in use (one per active issuer)

        for (int s = 0; s < T; ++s) {
            const int t = dir ? (T - 1 - s) : s;
            const int stage = s % SG;
            REC_PROBE(8);
            if (ok) ok = mbar_wait(&mbar_g[stage], (uint32_t)((s / SG) & 1));
            REC_PROBE(9);
            // rows of the quarter are gate-major: lanes [0,16) hold gates i,j ; lanes [16,32) gates f,o.
            // Sum the per-issuer accumulators as they complete.
            float zs[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) zs[k] = 0.f;
#pragma unroll
            for (int i = 0; i < NIW; ++i) {
                if (i < nacc) {
                    if (ok) ok = mbar_wait(&mbar_mma[i], (uint32_t)(s & 1));
                    if (i == 0) { REC_PROBE(10); }
                    tc_fence_after();
                    uint32_t a0[4], a1[4];
                    tmem_ld_16x256b_x1(t_addr + i * BG, a0);                       // (i | j) x 2 utts
                    tmem_ld_16x256b_x1(t_addr + i * BG + (16u << 16), a1);         // (f | o)
                    tmem_ld_wait();
#pragma unroll
                    for (int k = 0; k < 4; ++k) { zs[k] += __uint_as_float(a0[k]); zs[4 + k] += __uint_as_float(a1[k]); }
                }
            }
            REC_PROBE(11);
            const float* gt = Gsm + (size_t)stage * BG * REC_GROW + q * 32 + up;
            float zi[2], zj[2], zf[2], zo[2];
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                const float* gr = gt + (cg * 8 + 2 * g + j) * REC_GROW;
                const bool pad = (b0 + cg * 8 + 2 * g + j) >= B;         // padding utterance: keep it at exact zero
                zi[j] = pad ? 0.f : zs[j] + gr[0];
                zj[j] = pad ? 0.f : zs[2 + j] + gr[8];
                zf[j] = pad ? 0.f : zs[4 + j] + gr[16];
                zo[j] = pad ? 0.f : zs[6 + j] + gr[24];
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&mbar_gfree[stage]);          // G stage may be refilled
            tc_fence_before();                                       // our TMEM reads are complete (wait::ld above)
            REC_PROBE(12);
            float ig[2], fg[2], jt[2], cn[2], og[2], tc[2], mo[2];
            bool live[2];
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                const float cp = c_reg[j];
                ig[j] = sigmoidf_fast(zi[j] + wi * cp);
                fg[j] = sigmoidf_fast(zf[j] + p.forget_bias + wf * cp);
                jt[j] = tanhf_fast(zj[j]);
                cn[j] = fg[j] * cp + ig[j] * jt[j];
                og[j] = sigmoidf_fast(zo[j] + wo * cn[j]);
                tc[j] = tanhf_fast(cn[j]);
                live[j] = t < len_j[j];
                if (live[j]) c_reg[j] = cn[j];
                mo[j] = live[j] ? og[j] * tc[j] : 0.f;
            }
            const __half2 mh = __floats2half2_rn(mo[0], mo[1]);
            REC_PROBE(15);
            if (s + 1 < T) {
                // stage the warp's [8 utts][8 units] fp16 block = core matrix (q, cg) of this CTA's 1 KB operand slice
                // (double-buffered by step parity: the copies of step s are known to have been read once step s+2's
                // accumulator exists)
                unsigned char* slice = Msm + (s & 1) * 1024;
                __half* ms16 = reinterpret_cast<__half*>(slice + (q * 2 + cg) * 128);
                ms16[(2 * g) * 8 + up] = __low2half(mh);
                ms16[(2 * g + 1) * 8 + up] = __high2half(mh);
                asm volatile("bar.sync %0, 256;" ::"r"(1 + sg) : "memory");   // the 8 compute warps of the sub-group: slice complete
                // 1 KB bulk DSMEM copy per destination CTA (two per warp), destinations rotated by the sender's rank so
                // that no receiver is hit by all senders at once; completion is counted on the receiver's K-block barrier
                if (lane < 2 && 2 * rw + lane < NC) {
                    uint32_t dst = cta + (uint32_t)(2 * rw + lane); if (dst >= (uint32_t)NC) dst -= (uint32_t)NC;
                    fence_proxy_async_smem();
                    bulk_copy_s2c(mapa_shared(b_addr + (uint32_t)(((s + 1) & 1) * KB * BG * 128) + cta * 1024u, dst),
                                  smem_u32(slice), 1024u, mapa_shared(smem_u32(&mbar_op[((s + 1) & 1) * 8 + (cta >> 1)]), dst));
                }
            }
            REC_PROBE(13);
            // ---- off the critical path: outputs and saved activations ----
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                const int b = b0 + cg * 8 + 2 * g + j;
                if (b < B) {
                    const size_t idx = ((size_t)t * B + b) * ld2 + (size_t)dir * Hp + unit;
                    p.Mout[idx] = j ? __high2half(mh) : __low2half(mh);
                    if (p.gates) {
                        const __half2 g01 = __floats2half2_rn(ig[j], jt[j]), g23 = __floats2half2_rn(fg[j], og[j]);
                        p.gates[idx] = make_uint2(*reinterpret_cast<const uint32_t*>(&g01), *reinterpret_cast<const uint32_t*>(&g23));
                        p.cst[idx] = cn[j];
                    }
                    // final state = state at the last live step in this direction's own order
                    const bool last = dir ? (t == 0 && live[j]) : (t == len_j[j] - 1);
                    if (last && p.cfin) {
                        p.cfin[((size_t)b * 2 + dir) * Hp + unit] = cn[j];
                        p.mfin[((size_t)b * 2 + dir) * Hp + unit] = og[j] * tc[j];
                    }
                }
            }
            REC_PROBE(14);
        }
    }
    tc_fence_before();
    cluster_sync_all();          // nobody leaves while DSMEM traffic addressed to it may still be in flight
    if (warp == NCW * NSG) { tc_fence_after(); tmem_dealloc<512>(tmem_base); }
}

// =================================================================================================
// backward (BPTT)
// =================================================================================================
constexpr int BWD_NCW = 8;                    // compute warps per sub-group
constexpr int BWD_BARS = 8;                   // mma[4] dz red[2] (+pad)
template <int NSG> struct RecBwdCfg {
    static constexpr int THREADS = 32 * (NSG * BWD_NCW + 4);     // + 4 MMA issuer warps (one per 128-unit M tile), shared by the sub-groups
    __host__ __device__ static size_t sg_bytes(int NC) {          // dz operand + reduce double buffer (bf16) + partial staging (bf16)
        size_t b = (size_t)2 * REC_BG * 128 + (size_t)2 * NC * 32 * REC_BG * 2 + (size_t)4 * 8 * 1024;
        return (b + 1023) & ~(size_t)1023;
    }
    static size_t smem_bytes(int NC) { return 1024 + NSG * sg_bytes(NC) + NSG * BWD_BARS * 8 + 64; }
};

template <int NSG>
__global__ void __launch_bounds__(32 * (NSG * 8 + 4), 1)
lstm_rec_bwd_kernel(const RecBwdParams p)
{
    constexpr int BG = REC_BG, NCW = BWD_NCW;
    using Cfg = RecBwdCfg<NSG>;
    extern __shared__ unsigned char smem_raw[];
    unsigned char* smem = align_1024(smem_raw);
    const int Hp = p.Hp, NC = p.NC, T = p.T, B = p.B;
    const int MB = (Hp + 127) >> 7;                    // M tiles of 128 units
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    int role, sg, rw;
    if (warp < NCW * NSG) { role = 0; sg = warp / NCW; rw = warp % NCW; }
    else { role = 1; sg = 0; rw = warp - NCW * NSG; }          // issuer warps serve every sub-group in turn

    unsigned char* sgbase = smem + (size_t)sg * Cfg::sg_bytes(NC);
    unsigned char* Bp = sgbase;                                                  // [2][16 x 128 B] operand dz_t (K' = 128 gate rows)
    __nv_bfloat16* red = reinterpret_cast<__nv_bfloat16*>(Bp + (size_t)2 * BG * 128);          // [2][NC][32][16] bf16
    __nv_bfloat16* pst = red + (size_t)2 * NC * 32 * BG;                        // [2 step parities][2 tile slots][8 warps][32][16] bf16
    uint64_t* bars_all = reinterpret_cast<uint64_t*>(smem + (size_t)NSG * Cfg::sg_bytes(NC));
    uint64_t* bars = bars_all + (size_t)sg * BWD_BARS;
    uint64_t* mbar_mma = bars;                         // [4] partial dm tile j complete
    uint64_t* mbar_dz = bars + 4;                      //     dz_t staged by all compute threads
    uint64_t* mbar_red = bars + 5;                     // [2] partial dm slices from the whole cluster have landed
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars_all + (size_t)NSG * BWD_BARS);

    const uint32_t cta = cluster_ctarank();
    const int cid = (int)cluster_id_x();
    const int dir = cid & 1, bg = cid >> 1;
    const int b0 = (bg * NSG + sg) * BG;
    const bool sg_active = b0 < B;
    const size_t ld2 = (size_t)2 * Hp, ld8 = (size_t)8 * Hp;
    const uint32_t red_bytes = (uint32_t)(NC * 32 * BG * 2);    // one reduce buffer: a [32 units][16 utts] bf16 slice from every CTA

    if (threadIdx.x == 0) {
        for (int g2 = 0; g2 < NSG; ++g2) {
            uint64_t* bb = bars_all + (size_t)g2 * BWD_BARS;
            for (int i = 0; i < 4; ++i) mbar_init(&bb[i], 1);
            mbar_init(&bb[4], NCW * 32);
            mbar_init(&bb[5], 1);
            mbar_init(&bb[6], 1);
        }
        fence_mbar_init();
    }
    if (warp == NCW * NSG) tmem_alloc<512>(tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    // ---- W' -> tensor memory as A' = [units (lanes), own 128 gate rows (K')]: tile j in columns [64 j, 64 j + 64) ----
    if (role == 0) {
        const int q = warp & 3;
        for (int jt = warp >> 2; jt < MB; jt += 2 * NSG) {
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
    // arm the first reduce buffer of every active sub-group before anybody can send
    if (threadIdx.x < NSG && T > 1 && ((bg * NSG + (int)threadIdx.x) * BG < B))
        mbar_arrive_expect_tx(bars_all + (size_t)threadIdx.x * BWD_BARS + 5, red_bytes);
    tc_fence_before();
    cluster_sync_all();
    tc_fence_after();

    long long* prof = (blockIdx.x == 0 && lane == 0 && sg == 0 && rw == 0) ? g_rec_prof : nullptr;
    const int prof_steps = g_rec_prof_steps;
    bool ok = true;
    if (role == 1) {
        // ============================ MMA issuer warps: one 128-unit M tile each, sub-groups in turn ============================
        constexpr uint32_t idesc = make_idesc_bf16_f32(128, BG, 0, 0);           // bf16 x bf16, A (TMEM) K-major
        const int jt = rw;
        if (lane == 0 && jt < MB) {
            const uint32_t a_tmem = tmem_base + jt * 64;
            for (int s = 0; s < T && ok; ++s) {
                for (int g2 = 0; g2 < NSG && ok; ++g2) {
                    if ((bg * NSG + g2) * BG >= B) continue;                      // idle sub-group
                    uint64_t* bb = bars_all + (size_t)g2 * BWD_BARS;              // mma[4] dz red[2] of sub-group g2
                    const uint64_t bb0 = make_smem_desc_sw128(smem_u32(smem + (size_t)g2 * Cfg::sg_bytes(NC)), 16, 1024);
                    const uint32_t b_lo0 = (uint32_t)bb0, b_hi = (uint32_t)(bb0 >> 32);
                    const uint32_t d_tmem = tmem_base + REC_TMEM_ACC + (g2 * 4 + jt) * BG;
                    if (g2 == 0) { REC_PROBE(0); }
                    ok = mbar_wait(&bb[4], (uint32_t)(s & 1));                    // dz_t staged by all compute threads
                    if (!ok) break;
                    if (g2 == 0) { REC_PROBE(1); }
                    // the reduce buffer the NEXT step's partials go to: its previous contents were consumed in phase A
                    // of this step (all compute threads arrived on mbar_dz after reading them)
                    if (jt == 0 && s + 2 < T) mbar_arrive_expect_tx(&bb[5 + ((s + 1) & 1)], red_bytes);
                    tc_fence_after();
                    if (s + 1 < T) {                      // the last step's dm_{-1} is never used
#pragma unroll
                        for (int kk = 0; kk < 8; ++kk)
                            umma_f16_ts_lohi(d_tmem, a_tmem + 8 * kk, b_lo0 + (uint32_t)((kk >> 2) * (BG * 128 / 16) + (kk & 3) * 2), b_hi,
                                             idesc, kk ? 1u : 0u);
                    }
                    if (g2 == 0) { REC_PROBE(7); }
                    umma_commit(&bb[jt]);
                    if (g2 == 0) { REC_PROBE(2); }
                }
            }
        }
        __syncwarp();
    } else if (!sg_active) {
        // idle sub-group
    } else {
        // ============================ compute warps ============================
        const int cg = rw >> 2, q = rw & 3;
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
            const int b = b0 + cg * 8 + 2 * g + j;
            len_j[j] = (b < B) ? p.lens[b] : 0;
            dcc[j] = 0.f;
        }
        float db[4] = {0.f, 0.f, 0.f, 0.f};
        float dpf = 0.f, dpi = 0.f, dpo = 0.f;

        // raw prefetch of the next step's saved activations: loads only, no arithmetic, so they stay in flight
        // behind the current step (the first version applied tanh inside and stalled ~2600 cycles per step).
        // Row indices advance by a constant stride per step, so no 64-bit multiplies sit in the loop.
        struct Pre { uint2 gp[2]; float c[2], cp[2], dmo[2]; };
        const long long row_stride = (dir ? 1 : -1) * (long long)B * (long long)ld2;      // elements per time step, in scan order
        long long idx_j[2];                                                                // element index of (t(s), b_j, dir, unit)
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            const int b = b0 + cg * 8 + 2 * g + j;
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
        const uint32_t red_addr = smem_u32(red);
        // per-thread constants of the dz staging: smem offsets of the local MMA B operand and global columns
        uint32_t bp_off[2][4];
        int gcol[4];
#pragma unroll
        for (int gate = 0; gate < 4; ++gate) {
            gcol[gate] = dir * 4 * Hp + packed_col(unit, gate);
            const int kp = packed_col(ul, gate);
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                const int bl = cg * 8 + 2 * g + j;
                bp_off[j][gate] = (uint32_t)((kp >> 6) * (BG * 128) + (bl >> 3) * 1024 + (bl & 7) * 128 +
                                             ((((kp & 63) >> 3) ^ (bl & 7)) << 4) + (kp & 7) * 2);
            }
        }
        long long grow_j[2];                          // element index of dG row (t(s), b_j)
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            const int b = b0 + cg * 8 + 2 * g + j;
            grow_j[j] = ((long long)(dir ? 0 : T - 1) * B + (b < B ? b : 0)) * (long long)ld8;
        }
        const long long grow_stride = (dir ? 1 : -1) * (long long)B * (long long)ld8;

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
                const __nv_bfloat16* rb = red + (size_t)((s - 1) & 1) * NC * 32 * BG + (size_t)ul * BG + cg * 8 + 2 * g;
                float2 accv[4] = {{0.f, 0.f}, {0.f, 0.f}, {0.f, 0.f}, {0.f, 0.f}};
                int src = 0;
                for (; src + 4 <= NC; src += 4) {                  // 4 independent loads in flight
                    __nv_bfloat162 v[4];
#pragma unroll
                    for (int k = 0; k < 4; ++k) v[k] = *reinterpret_cast<const __nv_bfloat162*>(rb + (size_t)(src + k) * 32 * BG);
#pragma unroll
                    for (int k = 0; k < 4; ++k) { const float2 f = __bfloat1622float2(v[k]); accv[k].x += f.x; accv[k].y += f.y; }
                }
                for (; src < NC; ++src) {
                    const float2 f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(rb + (size_t)src * 32 * BG));
                    accv[0].x += f.x; accv[0].y += f.y;
                }
                dmr[0] = (accv[0].x + accv[1].x) + (accv[2].x + accv[3].x);
                dmr[1] = (accv[0].y + accv[1].y) + (accv[2].y + accv[3].y);
            }
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                const int b = b0 + cg * 8 + 2 * g + j;
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
                    *reinterpret_cast<__nv_bfloat16*>(Bp + bp_off[j][gate]) = v;      // local MMA B operand dz_t (K-major, 128B swizzle)
                    if (b < B) dgrow[gcol[gate]] = v;
                }
                grow_j[j] += grow_stride;
            }
            REC_PROBE(11);
            fence_proxy_async_smem();                  // locally staged dz_t -> visible to the tensor core
            mbar_arrive(mbar_dz);
            REC_PROBE(12);
            // ---- phase B: partial dm_{t-1} tiles -> owners' reduce buffers (bulk DSMEM copies) ----
            if (s + 1 < T) {
                // this warp's tiles: jt = cg, cg + 2 (one at a time: the kernel is register-bound at 640 threads)
#pragma unroll 1
                for (int k = 0; k < 2; ++k) {
                    const int jt = cg + 2 * k;
                    if (jt >= MB) break;
                    if (ok) ok = mbar_wait(&mbar_mma[jt], (uint32_t)(s & 1));
                    if (k == 0) { REC_PROBE(13); }
                    tc_fence_after();
                    uint32_t a[16];
                    tmem_ld_32x32b_x16(tmem_base + ((uint32_t)(q * 32) << 16) + REC_TMEM_ACC + (sg * 4 + jt) * BG, a);
                    tmem_ld_wait();
                    // the 32 rows of this (tile, quarter) are 32 consecutive units of ONE owner CTA: stage them as a
                    // [32][16] bf16 slice and ship it with a single 1 KB bulk DSMEM copy
                    __nv_bfloat16* pw = pst + (size_t)((((s & 1) * 2 + k) * 8 + rw) * 512);
                    uint4* ps = reinterpret_cast<uint4*>(pw + lane * 16);
                    ps[0] = make_uint4(pack_bf16x2(__uint_as_float(a[0]), __uint_as_float(a[1])), pack_bf16x2(__uint_as_float(a[2]), __uint_as_float(a[3])),
                                       pack_bf16x2(__uint_as_float(a[4]), __uint_as_float(a[5])), pack_bf16x2(__uint_as_float(a[6]), __uint_as_float(a[7])));
                    ps[1] = make_uint4(pack_bf16x2(__uint_as_float(a[8]), __uint_as_float(a[9])), pack_bf16x2(__uint_as_float(a[10]), __uint_as_float(a[11])),
                                       pack_bf16x2(__uint_as_float(a[12]), __uint_as_float(a[13])), pack_bf16x2(__uint_as_float(a[14]), __uint_as_float(a[15])));
                    __syncwarp();
                    const int u0 = jt * 128 + q * 32;
                    if (lane == 0 && u0 < Hp) {
                        const uint32_t owner = (uint32_t)(u0 >> 5);
                        fence_proxy_async_smem();
                        bulk_copy_s2c(mapa_shared(red_addr + (uint32_t)((((s & 1) * NC + (int)cta) * 32) * BG * 2), owner),
                                      smem_u32(pw), 1024u, mapa_shared(smem_u32(&mbar_red[s & 1]), owner));
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
        // each unit's sums now sit in its 4 lanes; column halves and sub-groups add up through the atomics
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
    if (warp == NCW * NSG) { tc_fence_after(); tmem_dealloc<512>(tmem_base); }
}

// =================================================================================================
// host side
// =================================================================================================
static bool rec_plan(int Hp, int& nc) {
    if (Hp < 64 || (Hp & 63) || Hp > 512) return false;
    nc = Hp / 32;                                   // 32 units (128 gate rows = one MMA M tile) per CTA
    return RecFwdCfg<2>::smem_bytes(Hp / 64) <= 232448 && RecBwdCfg<2>::smem_bytes(nc) <= 232448;
}

template <typename K, typename P>
static int launch_cluster(K kern, int grid, int threads, size_t smem, int nc, cudaStream_t st, const P& params) {
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
    cudaError_t e = cudaLaunchKernelEx(&cfg, kern, params);
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

// sub-groups per cluster: one 16-utterance group per cluster while all clusters fit in one wave (more SMs per
// utterance), two interleaved groups per cluster otherwise (half the clusters, and each group's exchange / MMA
// latency hides behind the other's gate math).  LCB_REC_NSG=1|2 overrides (experiments).
static int choose_nsg(int B, int nc, int which) {
    static int forced = -1;
    if (forced < 0) { const char* e = getenv("LCB_REC_NSG"); forced = e ? atoi(e) : 0; }
    if (forced == 1 || forced == 2) return forced;
    static int cap[2][17];
    if (cap[which][nc] == 0) {
        int n = which ? max_clusters(lstm_rec_bwd_kernel<1>, RecBwdCfg<1>::THREADS, RecBwdCfg<1>::smem_bytes(nc), nc)
                      : max_clusters(lstm_rec_fwd_kernel<1>, RecFwdCfg<1>::THREADS, RecFwdCfg<1>::smem_bytes(nc * 32 / 64), nc);
        cap[which][nc] = n > 0 ? n : 1;
    }
    const int groups = (B + REC_BG - 1) / REC_BG;
    return (2 * groups <= cap[which][nc]) ? 1 : 2;
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

// how many clusters of the forward (which = 0) / backward (which = 1) kernel (one sub-group per cluster) the device
// can keep resident at once
extern "C" int lcb_lstm_rec_max_clusters(int Hp, int which)
{
    int nc;
    if (!rec_plan(Hp, nc)) return LCB_ERR_UNSUPPORTED;
    int n = which ? max_clusters(lstm_rec_bwd_kernel<1>, RecBwdCfg<1>::THREADS, RecBwdCfg<1>::smem_bytes(nc), nc)
                  : max_clusters(lstm_rec_fwd_kernel<1>, RecFwdCfg<1>::THREADS, RecFwdCfg<1>::smem_bytes(Hp / 64), nc);
    return n < 0 ? LCB_ERR_CUDA : n;
}

extern "C" int lcb_lstm_rec_config(int Hp, int* units_per_cta_div32, int* cluster_size)
{
    int nc;
    if (!rec_plan(Hp, nc)) return LCB_ERR_UNSUPPORTED;
    if (units_per_cta_div32) *units_per_cta_div32 = 1;
    if (cluster_size) *cluster_size = nc;
    return LCB_OK;
}

extern "C" int lcb_lstm_rec_fwd(const float* G, const void* WfoldT, const float* peep, const int32_t* lens,
                                void* Mout, void* gates, float* cst, float* cfin, float* mfin,
                                int T, int B, int Hp, float forget_bias, void* stream)
{
    if (!G || !WfoldT || !lens || !Mout) return LCB_ERR_NULL_POINTER;
    if (T <= 0 || B <= 0) return LCB_ERR_BAD_SHAPE;
    if ((cfin == nullptr) != (mfin == nullptr) || (gates == nullptr) != (cst == nullptr)) return LCB_ERR_NULL_POINTER;
    int nc;
    if (!rec_plan(Hp, nc)) return LCB_ERR_UNSUPPORTED;
    if (((uintptr_t)G & 15) || ((uintptr_t)WfoldT & 15)) return LCB_ERR_MISALIGNED;
    RecFwdParams p;
    p.G = G; p.Wt = (const __half*)WfoldT; p.peep = peep; p.lens = lens; p.Mout = (__half*)Mout;
    p.gates = (uint2*)gates; p.cst = cst; p.cfin = cfin; p.mfin = mfin;
    p.T = T; p.B = B; p.Hp = Hp; p.NC = nc; p.forget_bias = forget_bias;
    const int nsg = choose_nsg(B, nc, 0);
    const int ncl = 2 * ((B + REC_BG * nsg - 1) / (REC_BG * nsg));
    if (nsg == 1)
        return launch_cluster(lstm_rec_fwd_kernel<1>, ncl * nc, RecFwdCfg<1>::THREADS, RecFwdCfg<1>::smem_bytes(Hp / 64), nc, (cudaStream_t)stream, p);
    return launch_cluster(lstm_rec_fwd_kernel<2>, ncl * nc, RecFwdCfg<2>::THREADS, RecFwdCfg<2>::smem_bytes(Hp / 64), nc, (cudaStream_t)stream, p);
}

extern "C" int lcb_lstm_rec_bwd(const float* dM, const void* gates, const float* cst, const void* Wfold, const float* peep,
                                const int32_t* lens, void* dG, float* dbias, float* dpeep,
                                int T, int B, int Hp, void* stream)
{
    if (!dM || !gates || !cst || !Wfold || !lens || !dG || !dbias) return LCB_ERR_NULL_POINTER;
    if (T <= 0 || B <= 0) return LCB_ERR_BAD_SHAPE;
    if ((peep == nullptr) != (dpeep == nullptr)) return LCB_ERR_NULL_POINTER;
    int nc;
    if (!rec_plan(Hp, nc)) return LCB_ERR_UNSUPPORTED;
    if ((uintptr_t)Wfold & 15) return LCB_ERR_MISALIGNED;
    RecBwdParams p;
    p.dM = dM; p.gates = (const uint2*)gates; p.cst = cst; p.W = (const __nv_bfloat16*)Wfold; p.peep = peep; p.lens = lens;
    p.dG = (__nv_bfloat16*)dG; p.dbias = dbias; p.dpeep = dpeep;
    p.T = T; p.B = B; p.Hp = Hp; p.NC = nc;
    const int nsg = choose_nsg(B, nc, 1);
    const int ncl = 2 * ((B + REC_BG * nsg - 1) / (REC_BG * nsg));
    if (nsg == 1)
        return launch_cluster(lstm_rec_bwd_kernel<1>, ncl * nc, RecBwdCfg<1>::THREADS, RecBwdCfg<1>::smem_bytes(nc), nc, (cudaStream_t)stream, p);
    return launch_cluster(lstm_rec_bwd_kernel<2>, ncl * nc, RecBwdCfg<2>::THREADS, RecBwdCfg<2>::smem_bytes(nc), nc, (cudaStream_t)stream, p);
}
