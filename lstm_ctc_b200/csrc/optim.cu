// optim.cu -- fused parameter update on the flat fp32 weight / gradient buffers (K5 in DESIGN.md).
//
// Replaces, per training step (nnet/graph.py:183-200):
//   l2_loss over variables whose name lacks 'bias' (:183-189)      -> g += l2 * w on the decayed ranges
//   tf.clip_by_global_norm(grads, clip_norm) (:190-192)            -> scale = clip / max(||g||, clip)
//   optimizer.apply_gradients: Adam / SGD / Momentum (:37-48,197)  -> one pass over w, g, (m, v)
// Two launches and no host synchronisation: the squared norm stays on the device.
#include "ptx.cuh"
#include "lstm_ctc_b200.h"

namespace lcb {

struct NoDecayRanges { int n; long long lo[24]; long long hi[24]; };

__device__ __forceinline__ bool in_nodecay(const NoDecayRanges& r, long long i) {
    for (int k = 0; k < r.n; ++k) if (i >= r.lo[k] && i < r.hi[k]) return true;
    return false;
}

// g += l2 * w (decayed elements); sumsq += sum g^2 (double)
__global__ void __launch_bounds__(256)
l2_sumsq_kernel(const float* __restrict__ w, float* __restrict__ g, long long n, float l2, NoDecayRanges nd, double* __restrict__ sumsq)
{
    __shared__ double red[8];
    double acc = 0.0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        float gi = g[i];
        if (l2 != 0.f && !in_nodecay(nd, i)) { gi += l2 * w[i]; g[i] = gi; }
        acc += (double)gi * (double)gi;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0.0;
        for (int k = 0; k < 8; ++k) s += red[k];
        atomicAdd(sumsq, s);
    }
}

// opt: 0 sgd, 1 momentum, 2 adam.  s1 = momentum accumulator / Adam m, s2 = Adam v.
__global__ void __launch_bounds__(256)
apply_update_kernel(float* __restrict__ w, const float* __restrict__ g, float* __restrict__ s1, float* __restrict__ s2,
                    long long n, int opt, float lr, float lr_t, float beta1, float beta2, float eps, float momentum,
                    float clip_norm, const double* __restrict__ sumsq, float* __restrict__ gnorm_out)
{
    const float gn = (float)sqrt(*sumsq);
    const float scale = (clip_norm > 0.f) ? clip_norm / fmaxf(gn, clip_norm) : 1.f;
    if (blockIdx.x == 0 && threadIdx.x == 0 && gnorm_out) *gnorm_out = gn;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const float gi = g[i] * scale;
        float wi = w[i];
        if (opt == 2) {
            const float m = beta1 * s1[i] + (1.f - beta1) * gi;
            const float v = beta2 * s2[i] + (1.f - beta2) * gi * gi;
            s1[i] = m; s2[i] = v;
            wi -= lr_t * m / (sqrtf(v) + eps);
        } else if (opt == 1) {
            const float a = momentum * s1[i] + gi;
            s1[i] = a;
            wi -= lr * a;
        } else {
            wi -= lr * gi;
        }
        w[i] = wi;
    }
}

}  // namespace lcb

using namespace lcb;

extern "C" int lcb_optimizer_step(float* w, float* g, float* s1, float* s2, long long n, int opt,
                                  float lr, long long step, float beta1, float beta2, float eps, float momentum,
                                  float l2, float clip_norm, const long long* nodecay_ranges_host, int n_ranges,
                                  double* sumsq_scratch, float* gnorm_out, void* stream)
{
    if (!w || !g || !sumsq_scratch) return LCB_ERR_NULL_POINTER;
    if (n <= 0 || opt < 0 || opt > 2 || n_ranges < 0 || n_ranges > 24 || step < 1) return LCB_ERR_BAD_SHAPE;
    if ((opt >= 1 && !s1) || (opt == 2 && !s2)) return LCB_ERR_NULL_POINTER;
    NoDecayRanges nd;
    nd.n = n_ranges;
    for (int k = 0; k < n_ranges; ++k) { nd.lo[k] = nodecay_ranges_host[2 * k]; nd.hi[k] = nodecay_ranges_host[2 * k + 1]; }
    cudaStream_t st = (cudaStream_t)stream;
    long long blocks = (n + 256 * 8 - 1) / (256 * 8);
    if (blocks > num_sms() * 8) blocks = num_sms() * 8;
    if (blocks < 1) blocks = 1;
    cudaMemsetAsync(sumsq_scratch, 0, sizeof(double), st);
    g_launches += 2; l2_sumsq_kernel<<<(int)blocks, 256, 0, st>>>(w, g, n, l2, nd, sumsq_scratch);
    // tf.train.AdamOptimizer: lr_t = lr * sqrt(1 - beta2^t) / (1 - beta1^t)
    const double lr_t = (double)lr * sqrt(1.0 - pow((double)beta2, (double)step)) / (1.0 - pow((double)beta1, (double)step));
    apply_update_kernel<<<(int)blocks, 256, 0, st>>>(w, g, s1, s2, n, opt, lr, (float)lr_t, beta1, beta2, eps, momentum,
                                                     clip_norm, sumsq_scratch, gnorm_out);
    return cudaGetLastError() == cudaSuccess ? LCB_OK : LCB_ERR_CUDA;
}
