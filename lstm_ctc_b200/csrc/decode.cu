// decode.cu -- validation / inference tails of the path.
//   greedy_decode : tf.nn.ctc_greedy_decoder(merge_repeated=True) (nnet/graph.py:138-142): per-frame argmax,
//                   collapse repeats, drop blank (= V-1).  One warp per utterance.
//   posterior     : softmax(smooth_factor * logits) (nnet/graph.py:236) with the optional log and
//                   log-prior subtraction of bin/nnet-forward.py:87-91 fused in.  One warp per row.
#include "ptx.cuh"
#include "lstm_ctc_b200.h"

namespace lcb {

__global__ void greedy_decode_kernel(const float* __restrict__ logits, const int* __restrict__ seq_len,
                                     int* __restrict__ out, int* __restrict__ out_len, int B, int T, int V)
{
    const int b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (b >= B) return;
    int Tb = seq_len[b]; Tb = Tb > T ? T : Tb;
    int prev = -1, n = 0;
    for (int t = 0; t < Tb; ++t) {
        const float* x = logits + ((size_t)b * T + t) * V;
        float best = -INFINITY; int bi = 0x7fffffff;
        for (int v = lane; v < V; v += 32) { const float f = x[v]; if (f > best) { best = f; bi = v; } }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float ob = __shfl_xor_sync(0xffffffffu, best, o);
            const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
            if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }       // first maximum wins (argmax)
        }
        if (bi != prev && bi != V - 1) { if (lane == 0) out[(size_t)b * T + n] = bi; ++n; }
        prev = bi;
    }
    if (lane == 0) out_len[b] = n;
}

__global__ void posterior_kernel(const float* __restrict__ logits, float* __restrict__ out, long long rows, int V,
                                 float smooth, int apply_log, const float* __restrict__ log_prior, int blank_to_front)
{
    const int lane = threadIdx.x & 31;
    for (long long r = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); r < rows;
         r += (long long)gridDim.x * (blockDim.x >> 5)) {
        const float* x = logits + (size_t)r * V;
        float* o = out + (size_t)r * V;
        float mx = -INFINITY;
        for (int v = lane; v < V; v += 32) mx = fmaxf(mx, smooth * x[v]);
#pragma unroll
        for (int k = 16; k > 0; k >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, k));
        float s = 0.f;
        for (int v = lane; v < V; v += 32) s += __expf(smooth * x[v] - mx);
#pragma unroll
        for (int k = 16; k > 0; k >>= 1) s += __shfl_xor_sync(0xffffffffu, s, k);
        const float lse = mx + logf(s);
        for (int v = lane; v < V; v += 32) {
            float y = smooth * x[v] - lse;
            if (!apply_log) y = __expf(y);
            if (log_prior) y -= log_prior[v];
            // blank_to_front: column V-1 (<blk>) moves to column 0, the others shift up by one -- the order EESEN's latgen-faster
            // expects (`select-feats $[ntargets-1],0-$[ntargets-2]`, scripts/decode_ctc_lat.sh:163)
            o[blank_to_front ? (v == V - 1 ? 0 : v + 1) : v] = y;
        }
    }
}

}  // namespace lcb

using namespace lcb;

extern "C" int lcb_greedy_decode(const float* logits, const int32_t* seq_len, int32_t* out, int32_t* out_len,
                                 int B, int T, int V, void* stream)
{
    if (!logits || !seq_len || !out || !out_len) return LCB_ERR_NULL_POINTER;
    if (B <= 0 || T <= 0 || V < 2) return LCB_ERR_BAD_SHAPE;
    g_launches += 1;
    greedy_decode_kernel<<<(B + 3) / 4, 128, 0, (cudaStream_t)stream>>>(logits, seq_len, out, out_len, B, T, V);
    return cudaGetLastError() == cudaSuccess ? LCB_OK : LCB_ERR_CUDA;
}

extern "C" int lcb_posterior(const float* logits, float* out, long long rows, int V, float smooth_factor,
                             int apply_log, const float* log_prior, int blank_to_front, void* stream)
{
    if (!logits || !out) return LCB_ERR_NULL_POINTER;
    if (rows <= 0 || V <= 0) return LCB_ERR_BAD_SHAPE;
    long long blocks = (rows + 7) / 8; if (blocks > num_sms() * 8) blocks = num_sms() * 8;
    g_launches += 1;
    posterior_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(logits, out, rows, V, smooth_factor, apply_log, log_prior, blank_to_front);
    return cudaGetLastError() == cudaSuccess ? LCB_OK : LCB_ERR_CUDA;
}
