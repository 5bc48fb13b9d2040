// mos.cu -- output layer of the acoustic model (K3 / K3b in DESIGN.md).
//
// Forward replaces nnet/moe.py:29-72 (create_moe) or, for num_experts == 0, the affine layer of
// nnet/bilstm.py:237-250, plus the reshape to [B,T,V] (bilstm.py:250):
//     pi[n,k]  = softmax_k(x_n . Wp[:,k] + bp[k])
//     y[n,v]   = sum_k pi[n,k] * tau * tanh(x_n . W[:, k*V+v] + b[k*V+v])
// One fused tcgen05 GEMM per 128-row tile with the mixture applied in the epilogue: the [N,K,V]
// expert tensor (moe.py:60) is never written anywhere.  Device weight layout is v-major
// (row v*K+k of Wall = column k*V+v of the reference's W) so the K experts of one target are
// adjacent accumulator columns and the sum over k is a running sum in one thread; the K prior rows
// follow at row K*V.  Work per tile: pass 0 = prior logits (N = K rounded to 16) -> softmax -> smem,
// then ceil(K*V/256) passes of 256 columns; passes alternate between two TMEM accumulators so the
// tanh/mixture epilogue of one pass overlaps the MMAs of the next.  Rows are time-major
// (n = t*B + b); the epilogue writes logits batch-major [B,T,V] as the CTC kernel and the reference API
// expect them.
//
// Backward (recompute, row-chunked): z is recomputed by lcb_gemm16 for a chunk of rows that stays
// L2-resident, mos_bwd_kernel turns (z, dy) into dz (bf16, same column order) with one warp per row,
// and dX / dW / db come from the generic GEMM + column sums.
#include <cuda_fp16.h>
#include "ptx.cuh"
#include "tma_host.h"
#include "lstm_ctc_b200.h"

namespace lcb {

// OUT_EPQ epilogue warps per TMEM lane quarter: the epilogue (tanh, mask, mixture sum: ~70 dependent instructions per column) is
// latency-bound with one warp per scheduler, and the tensor pipe finishes a 256-column pass ~6x faster than four warps consume it.
// Warps of a quarter share its 32 rows and take the 32-column chunks of a pass in turn -- possible when the K experts of a target
// never straddle a chunk (32 % K == 0, or the affine layer); otherwise only the first warp of each quarter works, as before.
constexpr int OUT_EPQ = 4;
constexpr int OUT_BM = 128, OUT_BN = 256, OUT_BK = 64, OUT_STAGES = 3, OUT_THREADS = 64 + 128 * OUT_EPQ;
constexpr int OUT_A_BYTES = OUT_BM * OUT_BK * 2, OUT_B_BYTES = OUT_BN * OUT_BK * 2;
constexpr int OUT_STAGE_BYTES = OUT_A_BYTES + OUT_B_BYTES;

struct OutFwdParams {
    const float* bias;     // [KV + K]  (affine: [V])
    float* logits;         // [B,T,V] batch-major
    int N, D2, T, B, V, K; // K == 0: affine
    int KV;                // K*V (affine: V)
    int Kp16;              // K rounded up to 16 (MMA N of the prior pass)
    int pis;               // row stride of pi in smem (odd)
    float tau;
    float inv_keep;        // 1 / keep_prob (1 when dropout is off)
    uint32_t thr;          // 16-bit keep threshold (65536: dropout off)
    unsigned long long seed_pi, seed_d;   // mask streams: mixture weights [N,K] (moe.py:46), expert logits [N,K*V] (moe.py:61)
};

__global__ void __launch_bounds__(OUT_THREADS, 1)
out_fwd_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW,
               const __grid_constant__ CUtensorMap tmWp, const OutFwdParams p)
{
    extern __shared__ unsigned char smem_raw[];
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    unsigned char* tiles = smem;
    float* pi_s = reinterpret_cast<float*>(smem + OUT_STAGES * OUT_STAGE_BYTES);            // [128][pis]
    float* bias_s = pi_s + (size_t)OUT_BM * p.pis;                                          // [KV + K] copy of the bias
    uint64_t* bars = reinterpret_cast<uint64_t*>(bias_s + p.KV + p.K + 2);
    bars = reinterpret_cast<uint64_t*>((reinterpret_cast<uintptr_t>(bars) + 7) & ~(uintptr_t)7);
    uint64_t* full_bar = bars;
    uint64_t* empty_bar = bars + OUT_STAGES;
    uint64_t* tfull_bar = bars + 2 * OUT_STAGES;
    uint64_t* tempty_bar = bars + 2 * OUT_STAGES + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * OUT_STAGES + 4);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tiles_m = (p.N + OUT_BM - 1) / OUT_BM;
    const int nkb = (p.D2 + OUT_BK - 1) / OUT_BK;
    const int nchunks = (p.KV + OUT_BN - 1) / OUT_BN;
    const int npass = nchunks + (p.K > 0 ? 1 : 0);

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmX); tma_prefetch_desc(&tmW); tma_prefetch_desc(&tmWp);
        for (int s = 0; s < OUT_STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        for (int s = 0; s < 2; ++s) { mbar_init(&tfull_bar[s], 1); mbar_init(&tempty_bar[s], 4 * OUT_EPQ); }
        fence_mbar_init();
    }
    if (warp == 1) tmem_alloc<512>(tmem_slot);
    for (int i = threadIdx.x; i < p.KV + p.K; i += blockDim.x) bias_s[i] = p.bias[i];
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            int stage = 0; uint32_t phase = 0; bool ok = true;
            for (int tile = blockIdx.x; tile < tiles_m && ok; tile += gridDim.x) {
                const int m0 = tile * OUT_BM;
                for (int ps = 0; ps < npass && ok; ++ps) {
                    const bool prior = (p.K > 0 && ps == 0);
                    const int n0 = prior ? p.KV : (ps - (p.K > 0 ? 1 : 0)) * OUT_BN;
                    const uint32_t bbytes = prior ? (uint32_t)p.Kp16 * OUT_BK * 2 : (uint32_t)OUT_B_BYTES;
                    for (int kb = 0; kb < nkb; ++kb) {
                        if (!mbar_wait(&empty_bar[stage], phase ^ 1)) { ok = false; break; }
                        unsigned char* sa = tiles + stage * OUT_STAGE_BYTES;
                        unsigned char* sb = sa + OUT_A_BYTES;
                        mbar_arrive_expect_tx(&full_bar[stage], OUT_A_BYTES + bbytes);
                        tma_load_2d(sa, &tmX, &full_bar[stage], kb * OUT_BK, m0);
                        tma_load_2d(sb, prior ? &tmWp : &tmW, &full_bar[stage], kb * OUT_BK, n0);
                        if (++stage == OUT_STAGES) { stage = 0; phase ^= 1; }
                    }
                }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        if (lane == 0) {
            // fp16 x fp16 -> f32, both K-major; N set per pass
            const uint32_t idesc_base = make_idesc_bf16_f32(OUT_BM, 8, 0, 0) & ~((7u << 7) | (7u << 10) | (63u << 17));
            int stage = 0; uint32_t phase = 0; int acc = 0; uint32_t acc_phase = 0; bool ok = true;
            for (int tile = blockIdx.x; tile < tiles_m && ok; tile += gridDim.x) {
                for (int ps = 0; ps < npass && ok; ++ps) {
                    const bool prior = (p.K > 0 && ps == 0);
                    const uint32_t nmma = prior ? (uint32_t)p.Kp16 : (uint32_t)OUT_BN;
                    const uint32_t idesc = idesc_base | ((nmma >> 3) << 17);
                    if (!mbar_wait(&tempty_bar[acc], acc_phase ^ 1)) { ok = false; break; }
                    tc_fence_after();
                    const uint32_t d_tmem = tmem_base + acc * OUT_BN;
                    for (int kb = 0; kb < nkb; ++kb) {
                        if (!mbar_wait(&full_bar[stage], phase)) { ok = false; break; }
                        tc_fence_after();
                        const uint32_t sa = smem_u32(tiles + stage * OUT_STAGE_BYTES);
                        const uint32_t sb = sa + OUT_A_BYTES;
#pragma unroll
                        for (int k = 0; k < OUT_BK / 16; ++k)
                            umma_f16_ss(d_tmem, make_smem_desc_sw128(sa + k * 32, 16, 1024), make_smem_desc_sw128(sb + k * 32, 16, 1024),
                                        idesc, (kb > 0 || k > 0) ? 1u : 0u);
                        umma_commit(&empty_bar[stage]);
                        if (++stage == OUT_STAGES) { stage = 0; phase ^= 1; }
                    }
                    umma_commit(&tfull_bar[acc]);
                    if (++acc == 2) { acc = 0; acc_phase ^= 1; }
                }
            }
        }
        __syncwarp();
    } else {
        // ================= epilogue: softmax over experts, tanh mixture, batch-major store =================
        const int q = warp & 3;                                   // TMEM lane quarter this warp may read (= warp % 4)
        const int eh = (warp - 2) >> 2;                           // which of the quarter's OUT_EPQ warps
        const int rloc = q * 32 + lane;
        float* pir = pi_s + (size_t)rloc * p.pis;
        int acc = 0; uint32_t acc_phase = 0; bool ok = true;
        const int K = p.K, V = p.V, KV = p.KV;
        const bool share = OUT_EPQ > 1 && (K == 0 || (32 % K) == 0);   // chunks of 32 columns hold whole targets: warps take turns
        const bool idle = !share && eh > 0;
        const int c_first = share ? eh * 32 : 0, c_step = share ? OUT_EPQ * 32 : 32;
        for (int tile = blockIdx.x; tile < tiles_m && ok; tile += gridDim.x) {
            const int n = tile * OUT_BM + rloc;
            const bool rowok = n < p.N;
            const int t = rowok ? n / p.B : 0, b = rowok ? n % p.B : 0;
            float* orow = p.logits + ((size_t)b * p.T + t) * V;
            int kk = 0, vv = 0; float accv = 0.f;                 // running mixture state across passes
            for (int ps = 0; ps < npass && ok; ++ps) {
                const bool prior = (K > 0 && ps == 0);
                if (!mbar_wait(&tfull_bar[acc], acc_phase)) { ok = false; break; }
                tc_fence_after();
                const uint32_t t_addr = tmem_base + ((uint32_t)(q * 32) << 16) + acc * OUT_BN;
                if (prior && OUT_EPQ > 1 && share)                 // the quarter's warps are done reading the previous tile's pi
                    asm volatile("bar.sync %0, %1;" ::"r"(1 + q), "n"(32 * OUT_EPQ) : "memory");
                if (idle) {
                    // (K does not divide 32: the first warp of the quarter carries the running mixture state through all columns)
                } else if (prior && eh > 0) {
                    // the quarter's first warp forms pi
                } else if (prior) {
                    // pass A: max; pass B: exp, sum -> smem; then normalise
                    float mx = -INFINITY;
                    for (int c0 = 0; c0 < K; c0 += 32) {
                        uint32_t r[32];
                        tmem_ld_32x32b_x32(t_addr + c0, r);
                        tmem_ld_wait();
#pragma unroll
                        for (int j = 0; j < 32; ++j) if (c0 + j < K) mx = fmaxf(mx, __uint_as_float(r[j]) + p.bias[KV + c0 + j]);
                    }
                    float sum = 0.f;
                    for (int c0 = 0; c0 < K; c0 += 32) {
                        uint32_t r[32];
                        tmem_ld_32x32b_x32(t_addr + c0, r);
                        tmem_ld_wait();
#pragma unroll
                        for (int j = 0; j < 32; ++j) if (c0 + j < K) {
                            const float e = __expf(__uint_as_float(r[j]) + p.bias[KV + c0 + j] - mx);
                            pir[c0 + j] = e; sum += e;
                        }
                    }
                    const float inv = 1.f / sum;
                    // y_prior = dropout(softmax(...))  (moe.py:45-46): fold mask and 1/keep into the stored weights
                    for (int k = 0; k < K; ++k)
                        pir[k] = rng_keepq(p.seed_pi, (uint64_t)n * K + k, p.thr) ? pir[k] * inv * p.inv_keep : 0.f;
                } else {
                    const int c_base = (ps - (K > 0 ? 1 : 0)) * OUT_BN;
                    for (int c0 = c_first; c0 < OUT_BN && c_base + c0 < KV; c0 += c_step) {
                        uint32_t r[32];
                        tmem_ld_32x32b_x32(t_addr + c0, r);
                        tmem_ld_wait();
                        if (share && K > 0) { kk = 0; accv = 0.f; vv = (c_base + c0) / K; }      // a chunk starts at a target boundary
                        if (K > 0) {
                            // phase 1 (independent per column -> instruction-level parallelism): th[j] = dropout mask * tanh(z + b).
                            // One 64-bit hash decides four consecutive elements of the [N, K*V] mask stream (moe.py:61); the
                            // element index n*KV + col is not 4-aligned in general, so the hash is refreshed at block boundaries.
                            float th[32];
                            const int colb = c_base + c0;
                            if (p.thr < 65536u) {
                                const unsigned long long e0 = (unsigned long long)n * (unsigned long long)KV + (unsigned long long)colb;
                                const int ph = (int)(e0 & 3ull);
                                unsigned long long W[9];                   // the <= 9 four-element blocks these 32 columns touch
#pragma unroll
                                for (int q4 = 0; q4 < 9; ++q4) W[q4] = rng_u64(p.seed_d, (e0 >> 2) + q4);
#pragma unroll
                                for (int j = 0; j < 32; ++j) {
                                    const int e = (ph + j) & 3;
                                    const unsigned long long wd = (ph + (j & 3) >= 4) ? W[(j >> 2) + 1] : W[j >> 2];
                                    const bool keep = rng_keep16(wd, e, p.thr);
                                    const float bz = (colb + j < KV) ? bias_s[colb + j] : 0.f;
                                    th[j] = keep ? tanhf_fast(__uint_as_float(r[j]) + bz) : 0.f;
                                }
                            } else {
#pragma unroll
                                for (int j = 0; j < 32; ++j) {
                                    const float bz = (colb + j < KV) ? bias_s[colb + j] : 0.f;
                                    th[j] = tanhf_fast(__uint_as_float(r[j]) + bz);
                                }
                            }
                            // phase 2 (serial in k): y_v = tau/keep * sum_k pi_k * th_k
#pragma unroll
                            for (int j = 0; j < 32; ++j) {
                                if (colb + j < KV) {
                                    accv = fmaf(pir[kk], th[j], accv);
                                    if (++kk == K) {
                                        if (rowok) orow[vv] = p.tau * p.inv_keep * accv;
                                        ++vv; kk = 0; accv = 0.f;
                                    }
                                }
                            }
                        } else {
#pragma unroll
                            for (int j = 0; j < 32; ++j) {
                                const int col = c_base + c0 + j;
                                if (col < KV && rowok) orow[col] = __uint_as_float(r[j]) + __ldg(p.bias + col);
                            }
                        }
                    }
                }
                if (prior && OUT_EPQ > 1 && share)                 // pi of this tile is in shared memory for the whole quarter
                    asm volatile("bar.sync %0, %1;" ::"r"(1 + q), "n"(32 * OUT_EPQ) : "memory");
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&tempty_bar[acc]);
                if (++acc == 2) { acc = 0; acc_phase ^= 1; }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) { tc_fence_after(); tmem_dealloc<512>(tmem_base); }
}

// ------------------------------------------------------------------------------------------------
// backward elementwise: one warp per row.
//   Z   [R, ldz] f32   recomputed x.Wall^T + bias for rows [n0, n0+R) (v-major expert cols, then K prior cols)
//   dY  [B,T,V]  f32   d loss / d logits (batch-major, from the CTC kernel)
//   dZ  [R, ldz] bf16  d loss / d z  (same column order; pad columns zeroed)
__global__ void __launch_bounds__(256)
mos_bwd_kernel(const float* __restrict__ Z, const float* __restrict__ dY, __nv_bfloat16* __restrict__ dZ,
               int n0, int R, int ldz, int T, int B, int V, int K, float tau,
               float inv_keep, uint32_t thr, unsigned long long seed_pi, unsigned long long seed_d)
{
    extern __shared__ float sm[];                     // per warp: dpi[K], pi[K]
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float* dpi = sm + (size_t)w * 2 * K;
    float* pi = dpi + K;
    const int KV = K * V;
    for (int r = blockIdx.x * 8 + w; r < R; r += gridDim.x * 8) {
        const int n = n0 + r;
        const int t = n / B, b = n % B;
        const float* z = Z + (size_t)r * ldz;
        const float* dy = dY + ((size_t)b * T + t) * V;
        __nv_bfloat16* dz = dZ + (size_t)r * ldz;
        for (int k = lane; k < K; k += 32) dpi[k] = 0.f;
        // prior softmax
        float mx = -INFINITY;
        for (int k = lane; k < K; k += 32) mx = fmaxf(mx, z[KV + k]);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        float s = 0.f;
        for (int k = lane; k < K; k += 32) { const float e = __expf(z[KV + k] - mx); pi[k] = e; s += e; }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        __syncwarp();
        const float inv = 1.f / s;
        for (int k = lane; k < K; k += 32) pi[k] *= inv;
        __syncwarp();
        // expert columns: dz = dy * pi * tau * (1 - tanh^2);  dpi[k] += dy * tau * tanh
        // forward: y_v = sum_k pi'_k * tau * th_k * m2/keep,  pi'_k = pi_k * m1/keep
        for (int c = lane; c < KV; c += 32) {
            const int v = c / K, k = c - v * K;
            const float th = tanhf_fast(z[c]);
            const float m2 = rng_keepq(seed_d, (uint64_t)n * KV + c, thr) ? inv_keep : 0.f;
            const float m1 = rng_keepq(seed_pi, (uint64_t)n * K + k, thr) ? inv_keep : 0.f;
            const float g = dy[v] * tau * m2;
            atomicAdd(&dpi[k], g * th * m1);                       // d loss / d pi_k
            dz[c] = __float2bfloat16(g * pi[k] * m1 * (1.f - th * th));
        }
        __syncwarp();
        // softmax backward: dlogit_k = pi_k * (dpi_k - sum_j pi_j dpi_j)
        float dot = 0.f;
        for (int k = lane; k < K; k += 32) dot += pi[k] * dpi[k];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, o);
        for (int k = lane; k < K; k += 32) dz[KV + k] = __float2bfloat16(pi[k] * (dpi[k] - dot));
        for (int c = KV + K + lane; c < ldz; c += 32) dz[c] = __float2bfloat16(0.f);
        __syncwarp();
    }
}

// The same, four consecutive expert columns per lane (K % 4 == 0, ldz % 4 == 0, 16-byte aligned rows): one float4 load, ONE
// hash word per mask stream per four elements (the streams give 16 bits to each of four consecutive element indices, and both
// n*K*V + c and n*K + k are multiples of four here), one 8-byte store.  FIXEDK (128 % K == 0): a lane meets the same four experts
// in every iteration, so d loss / d pi accumulates in registers and is reduced with shuffles; otherwise shared-memory atomics.
// The scalar kernel above spent ~200 instructions per element on the per-element hashes, the division and the atomics.
template <bool FIXEDK>
__global__ void __launch_bounds__(256)
mos_bwd_v4_kernel(const float* __restrict__ Z, const float* __restrict__ dY, __nv_bfloat16* __restrict__ dZ,
                  int n0, int R, int ldz, int T, int B, int V, int K, float tau,
                  float inv_keep, uint32_t thr, unsigned long long seed_pi, unsigned long long seed_d)
{
    extern __shared__ float sm[];                     // per warp: dpi[K], pi[K] (masked: pi * m1), m1[K]
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float* dpi = sm + (size_t)w * 3 * K;
    float* pi = dpi + K;
    float* m1s = pi + K;
    const int KV = K * V;
    const bool drop = thr < 65536u;
    const int q = K >> 2;                              // lanes per period of the expert index (FIXEDK)
    const int vstep = FIXEDK ? 128 / K : 0;
    for (int r = blockIdx.x * 8 + w; r < R; r += gridDim.x * 8) {
        const int n = n0 + r;
        const int t = n / B, b = n - t * B;
        const float* z = Z + (size_t)r * ldz;
        const float* dy = dY + ((size_t)b * T + t) * V;
        __nv_bfloat16* dz = dZ + (size_t)r * ldz;
        // prior softmax, mixture-weight mask
        float mx = -INFINITY;
        for (int k = lane; k < K; k += 32) mx = fmaxf(mx, z[KV + k]);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        float s = 0.f;
        for (int k = lane; k < K; k += 32) { const float e = __expf(z[KV + k] - mx); pi[k] = e; s += e; if (!FIXEDK) dpi[k] = 0.f; }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        const float inv = 1.f / s;
        for (int k4 = lane * 4; k4 < K; k4 += 128) {
            const uint64_t word = drop ? rng_u64(seed_pi, ((uint64_t)n * K + k4) >> 2) : 0ull;
#pragma unroll
            for (int e = 0; e < 4; ++e) m1s[k4 + e] = (!drop || rng_keep16(word, e, thr)) ? inv_keep : 0.f;
        }
        __syncwarp();
        for (int k = lane; k < K; k += 32) pi[k] *= inv;          // plain pi (softmax backward needs it unmasked)
        __syncwarp();
        // expert columns: dz = dy * pi * tau * (1 - tanh^2);  dpi[k] += dy * tau * tanh
        // forward: y_v = sum_k pi'_k * tau * th_k * m2/keep,  pi'_k = pi_k * m1/keep
        float acc[4] = {0.f, 0.f, 0.f, 0.f};
        int v = FIXEDK ? (lane * 4) / K : 0;
        const int kfix = FIXEDK ? lane * 4 - v * K : 0;
        for (int c = lane * 4; c < KV; c += 128) {
            int k;
            if (FIXEDK) k = kfix; else { v = c / K; k = c - v * K; }
            const float4 z4 = *reinterpret_cast<const float4*>(z + c);
            const uint64_t word = drop ? rng_u64(seed_d, ((uint64_t)n * KV + c) >> 2) : 0ull;
            const float dyv = dy[v] * tau;
            const float4 p4 = *reinterpret_cast<const float4*>(pi + k);
            const float4 a4 = *reinterpret_cast<const float4*>(m1s + k);
            const float zz[4] = {z4.x, z4.y, z4.z, z4.w}, pp[4] = {p4.x, p4.y, p4.z, p4.w}, aa[4] = {a4.x, a4.y, a4.z, a4.w};
            float o4[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const float th = tanhf_fast(zz[e]);
                const float g = (!drop || rng_keep16(word, e, thr)) ? dyv * inv_keep : 0.f;
                const float gm = g * aa[e];
                if (FIXEDK) acc[e] += gm * th; else if (gm != 0.f) atomicAdd(&dpi[k + e], gm * th);
                o4[e] = gm * pp[e] * (1.f - th * th);
            }
            *reinterpret_cast<uint2*>(dz + c) = make_uint2(pack_bf16x2(o4[0], o4[1]), pack_bf16x2(o4[2], o4[3]));
            if (FIXEDK) v += vstep;
        }
        if (FIXEDK) {
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                float a = acc[e];
                for (int o = 16; o >= q; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
                if (lane < q) dpi[lane * 4 + e] = a;
            }
        }
        __syncwarp();
        // softmax backward: dlogit_k = pi_k * (dpi_k - sum_j pi_j dpi_j)
        float dot = 0.f;
        for (int k = lane; k < K; k += 32) dot += pi[k] * dpi[k];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, o);
        for (int k = lane; k < K; k += 32) dz[KV + k] = __float2bfloat16(pi[k] * (dpi[k] - dot));
        for (int c = KV + K + lane; c < ldz; c += 32) dz[c] = __float2bfloat16(0.f);
        __syncwarp();
    }
}

// dlogits [B,T,V] f32 batch-major -> [N, ldo] bf16 time-major (pad columns zero): the affine layer's dZ
__global__ void pack_dlogits_kernel(const float* __restrict__ dY, __nv_bfloat16* __restrict__ out, int T, int B, int V, int ldo) {
    const size_t total = (size_t)T * B * ldo;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int c = (int)(i % ldo);
        const size_t n = i / ldo;
        const int b = (int)(n % B), t = (int)(n / B);
        out[i] = __float2bfloat16(c < V ? dY[((size_t)b * T + t) * V + c] : 0.f);
    }
}

}  // namespace lcb

using namespace lcb;

extern "C" int lcb_output_fwd(const void* X, int ldx, const void* Wall, const float* bias, float* logits,
                              int T, int B, int D2, int V, int K, float tau, float keep_prob, unsigned long long seed,
                              void* stream)
{
    if (!X || !Wall || !bias || !logits) return LCB_ERR_NULL_POINTER;
    if (T <= 0 || B <= 0 || D2 <= 0 || V <= 0 || K < 0) return LCB_ERR_BAD_SHAPE;
    if (!(keep_prob > 0.f) || keep_prob > 1.f) return LCB_ERR_BAD_SHAPE;
    if (K > 128) return LCB_ERR_UNSUPPORTED;
    if ((ldx & 7) || (D2 & 7) || ((uintptr_t)X & 15) || ((uintptr_t)Wall & 15)) return LCB_ERR_MISALIGNED;
    OutFwdParams p;
    p.bias = bias; p.logits = logits; p.N = T * B; p.D2 = D2; p.T = T; p.B = B; p.V = V; p.K = K;
    p.KV = (K > 0 ? K : 1) * V;
    p.Kp16 = K > 0 ? ((K + 15) & ~15) : 16;
    p.pis = (K | 1) + 2 * (K > 0 ? 0 : 0);
    p.tau = tau;
    p.inv_keep = 1.f / keep_prob;
    p.thr = keep_threshold16(keep_prob);
    p.seed_pi = seed; p.seed_d = seed ^ 0xD1B54A32D192ED03ull;
    const int rows = p.KV + K;
    CUtensorMap tx, tw, twp;
    if (!make_tmap_2d_bf16(&tx, X, (uint64_t)p.N, (uint64_t)D2, (uint64_t)ldx, OUT_BM, OUT_BK)) return LCB_ERR_CUDA;
    if (!make_tmap_2d_bf16(&tw, Wall, (uint64_t)rows, (uint64_t)D2, (uint64_t)D2, OUT_BN, OUT_BK)) return LCB_ERR_CUDA;
    if (!make_tmap_2d_bf16(&twp, Wall, (uint64_t)rows, (uint64_t)D2, (uint64_t)D2, (uint32_t)p.Kp16, OUT_BK)) return LCB_ERR_CUDA;
    const size_t smem = 1024 + (size_t)OUT_STAGES * OUT_STAGE_BYTES + (size_t)OUT_BM * p.pis * 4 + (size_t)(p.KV + K + 2) * 4 + 512;
    if (smem > 227 * 1024) return LCB_ERR_UNSUPPORTED;
    static size_t smem_set = 0;
    if (smem > smem_set) {
        if (cudaFuncSetAttribute(out_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return LCB_ERR_CUDA;
        smem_set = smem;
    }
    const int tiles_m = (p.N + OUT_BM - 1) / OUT_BM;
    const int grid = tiles_m < num_sms() ? tiles_m : num_sms();
    g_launches += 1; out_fwd_kernel<<<grid, OUT_THREADS, smem, (cudaStream_t)stream>>>(tx, tw, twp, p);
    return cudaGetLastError() == cudaSuccess ? LCB_OK : LCB_ERR_CUDA;
}

extern "C" int lcb_mos_bwd_dz(const float* Z, const float* dlogits, void* dZ, int n0, int R, int ldz,
                              int T, int B, int V, int K, float tau, float keep_prob, unsigned long long seed, void* stream)
{
    if (!(keep_prob > 0.f) || keep_prob > 1.f) return LCB_ERR_BAD_SHAPE;
    if (!Z || !dlogits || !dZ) return LCB_ERR_NULL_POINTER;
    if (R <= 0 || K <= 0 || V <= 0 || ldz < K * V + K) return LCB_ERR_BAD_SHAPE;
    int grid = (R + 7) / 8; if (grid > num_sms() * 8) grid = num_sms() * 8;
    const float inv_keep = 1.f / keep_prob;
    const uint32_t thr = keep_threshold16(keep_prob);
    const unsigned long long seed_d = seed ^ 0xD1B54A32D192ED03ull;
    const bool vec = (K % 4 == 0) && (ldz % 4 == 0) && (((uintptr_t)Z & 15) == 0) && (((uintptr_t)dZ & 7) == 0);
    g_launches += 1;
    if (vec) {
        const size_t smem = (size_t)8 * 3 * K * sizeof(float);
        if (128 % K == 0)
            mos_bwd_v4_kernel<true><<<grid, 256, smem, (cudaStream_t)stream>>>(Z, dlogits, (__nv_bfloat16*)dZ, n0, R, ldz, T, B, V, K, tau,
                                                                                inv_keep, thr, seed, seed_d);
        else
            mos_bwd_v4_kernel<false><<<grid, 256, smem, (cudaStream_t)stream>>>(Z, dlogits, (__nv_bfloat16*)dZ, n0, R, ldz, T, B, V, K, tau,
                                                                                 inv_keep, thr, seed, seed_d);
    } else {
        const size_t smem = (size_t)8 * 2 * K * sizeof(float);
        mos_bwd_kernel<<<grid, 256, smem, (cudaStream_t)stream>>>(Z, dlogits, (__nv_bfloat16*)dZ, n0, R, ldz, T, B, V, K, tau,
                                                                  inv_keep, thr, seed, seed_d);
    }
    return cudaGetLastError() == cudaSuccess ? LCB_OK : LCB_ERR_CUDA;
}

extern "C" int lcb_pack_dlogits(const float* dlogits, void* out, int T, int B, int V, int ldo, void* stream)
{
    if (!dlogits || !out) return LCB_ERR_NULL_POINTER;
    if (T <= 0 || B <= 0 || V <= 0 || ldo < V) return LCB_ERR_BAD_SHAPE;
    const size_t total = (size_t)T * B * ldo;
    size_t blocks = (total + 255) / 256; if (blocks > num_sms() * 16) blocks = num_sms() * 16;
    g_launches += 1; pack_dlogits_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(dlogits, (__nv_bfloat16*)out, T, B, V, ldo);
    return cudaGetLastError() == cudaSuccess ? LCB_OK : LCB_ERR_CUDA;
}
