// elementwise.cu -- small HBM-bound helper kernels of the hot path (casts, packing, reductions).
// They replace TF's implicit layout ops on the path: the [B,T,D] batch-major pipeline tensor
// (nnet/pipeline.py:35-61) is re-laid time-major once, weights are re-cast to bf16 once per step.
#include <cuda_fp16.h>
#include "ptx.cuh"
#include "lstm_ctc_b200.h"

namespace lcb {

__device__ __forceinline__ float sat_f16(float x) { return fminf(fmaxf(x, -65504.f), 65504.f); }
__device__ __forceinline__ uint32_t pack_f16x2(float lo, float hi) {
    __half2 v = __floats2half2_rn(sat_f16(lo), sat_f16(hi));
    return *reinterpret_cast<uint32_t*>(&v);
}
// F16 = false: bf16 destination; true: fp16 destination (saturating)
template <bool F16>
__global__ void cast_f32_16_kernel(const float* __restrict__ src, uint16_t* __restrict__ dst, size_t n) {
    size_t i = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
    const size_t stride = (size_t)gridDim.x * blockDim.x * 4;
    for (; i + 3 < n; i += stride) {
        const float4 v = *reinterpret_cast<const float4*>(src + i);
        if constexpr (F16) *reinterpret_cast<uint2*>(dst + i) = make_uint2(pack_f16x2(v.x, v.y), pack_f16x2(v.z, v.w));
        else *reinterpret_cast<uint2*>(dst + i) = make_uint2(pack_bf16x2(v.x, v.y), pack_bf16x2(v.z, v.w));
    }
    if (i < n) for (size_t k = i; k < n && k < i + 4; ++k) {
        if constexpr (F16) { __half h = __float2half_rn(sat_f16(src[k])); dst[k] = *reinterpret_cast<uint16_t*>(&h); }
        else { __nv_bfloat16 h = __float2bfloat16(src[k]); dst[k] = *reinterpret_cast<uint16_t*>(&h); }
    }
}

// fp16 -> bf16 (wgrad pairs bf16 gradients with a bf16 copy of the fp16 forward activations:
// tcgen05 kind::f16 needs A and B in the same 16-bit format)
__global__ void f16_to_bf16_kernel(const __half2* __restrict__ src, __nv_bfloat162* __restrict__ dst, size_t n2) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n2; i += (size_t)gridDim.x * blockDim.x) {
        const float2 v = __half22float2(src[i]);
        dst[i] = __floats2bfloat162_rn(v.x, v.y);
    }
}

// in-place inverted dropout on a 16-bit tensor: x[i] = keep(seed, i) ? x[i] / keep_prob : 0
// (DropoutWrapper(output_keep_prob) on the LSTM layer outputs, nnet/bilstm.py:128,137; the same call on the
// bf16 gradient with the same seed is its backward)
template <bool F16>
__device__ __forceinline__ uint32_t drop2(uint32_t v2, bool k0, bool k1, float inv_keep) {
    float lo, hi;
    if constexpr (F16) { const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&v2)); lo = f.x; hi = f.y; }
    else { const float2 f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&v2)); lo = f.x; hi = f.y; }
    lo = k0 ? lo * inv_keep : 0.f;
    hi = k1 ? hi * inv_keep : 0.f;
    if constexpr (F16) return pack_f16x2(lo, hi);
    else return pack_bf16x2(lo, hi);
}
// 8 elements (16 bytes) per thread and iteration, two hashes
template <bool F16>
__global__ void dropout16_kernel(uint16_t* __restrict__ x, size_t n, float inv_keep, uint32_t thr16, uint64_t seed) {
    const size_t n8 = n >> 3;
    uint4* x8 = reinterpret_cast<uint4*>(x);
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += (size_t)gridDim.x * blockDim.x) {
        uint4 v = x8[i];
        const uint64_t w0 = rng_u64(seed, 2 * i), w1 = rng_u64(seed, 2 * i + 1);
        v.x = drop2<F16>(v.x, rng_keep16(w0, 0, thr16), rng_keep16(w0, 1, thr16), inv_keep);
        v.y = drop2<F16>(v.y, rng_keep16(w0, 2, thr16), rng_keep16(w0, 3, thr16), inv_keep);
        v.z = drop2<F16>(v.z, rng_keep16(w1, 0, thr16), rng_keep16(w1, 1, thr16), inv_keep);
        v.w = drop2<F16>(v.w, rng_keep16(w1, 2, thr16), rng_keep16(w1, 3, thr16), inv_keep);
        x8[i] = v;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {                 // tail (n not a multiple of 8)
        for (size_t k = n8 << 3; k < n; ++k) {
            const bool keep = rng_keep16(rng_u64(seed, k >> 2), (int)(k & 3), thr16);
            float v;
            if constexpr (F16) v = __half2float(reinterpret_cast<__half*>(x)[k]);
            else v = __bfloat162float(reinterpret_cast<__nv_bfloat16*>(x)[k]);
            v = keep ? v * inv_keep : 0.f;
            if constexpr (F16) reinterpret_cast<__half*>(x)[k] = __float2half_rn(sat_f16(v));
            else reinterpret_cast<__nv_bfloat16*>(x)[k] = __float2bfloat16(v);
        }
    }
}
__global__ void dropout_mask_kernel(uint8_t* __restrict__ m, size_t n, uint32_t thr16, uint64_t seed) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        m[i] = rng_keep16(rng_u64(seed, i >> 2), (int)(i & 3), thr16) ? 1 : 0;
}

// y += x on fp16 tensors (layer-0 residual: finput = finput + concat(fwd, bwd), nnet/bilstm.py:199-200)
__global__ void add_f16_kernel(__half2* __restrict__ y, const __half2* __restrict__ x, size_t n2) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n2; i += (size_t)gridDim.x * blockDim.x) {
        const float2 a = __half22float2(y[i]), b = __half22float2(x[i]);
        y[i] = __floats2half2_rn(sat_f16(a.x + b.x), sat_f16(a.y + b.y));
    }
}

// y[r, c] += mask * x[r, c] / keep on 16-bit 2-D tensors, mask = element (mask_base + r*ldm + c) of the (seed) stream that
// lcb_gemm16_dropout uses for the same output: the residual connection of ResidualWrapper under a DropoutWrapper
// (nnet/lstm.py:236-260: out = dropout(x + cell(x))) -- the GEMM epilogue already wrote dropout(cell(x)), this adds dropout(x)
// with the SAME mask; in backward it adds the masked output gradient to the input gradient.
template <bool F16>
__global__ void masked_add16_kernel(uint16_t* __restrict__ y, int ldy, const uint16_t* __restrict__ x, int ldx, long long rows, int cols,
                                    float inv_keep, uint32_t thr16, uint64_t seed, unsigned long long mask_base, int ldm) {
    const long long total = rows * cols;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long r = i / cols;
        const int c = (int)(i - r * cols);
        const unsigned long long k = mask_base + (unsigned long long)r * ldm + c;
        const bool keep = thr16 >= 65536u || rng_keep16(rng_u64(seed, k >> 2), (int)(k & 3), thr16);
        if (!keep) continue;
        float a, b;
        if constexpr (F16) {
            a = __half2float(reinterpret_cast<const __half*>(y)[r * ldy + c]);
            b = __half2float(reinterpret_cast<const __half*>(x)[r * ldx + c]);
            reinterpret_cast<__half*>(y)[r * ldy + c] = __float2half_rn(sat_f16(a + b * inv_keep));
        } else {
            a = __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(y)[r * ldy + c]);
            b = __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(x)[r * ldx + c]);
            reinterpret_cast<__nv_bfloat16*>(y)[r * ldy + c] = __float2bfloat16(a + b * inv_keep);
        }
    }
}

// label-smoothing regulariser (nnet/bilstm.py:254-269): per row p = softmax(logits),
//   loss += w * sum_v p_v (log p_v - q_v),   q = log(1/V) (uniform) or the given log prior;
//   dlogits_v += w * p_v * ((log p_v - q_v) - sum_u p_u (log p_u - q_u)).   Over ALL B*T rows, padding included,
// exactly as the reference sums it.  One warp per row.
__global__ void __launch_bounds__(256)
label_smooth_kernel(const float* __restrict__ logits, float* __restrict__ dlogits, long long rows, int V, float w,
                    const float* __restrict__ log_prior, float* __restrict__ loss_out)
{
    const int lane = threadIdx.x & 31;
    float local = 0.f;
    const float qu = -logf((float)V);
    for (long long r = (long long)blockIdx.x * 8 + (threadIdx.x >> 5); r < rows; r += (long long)gridDim.x * 8) {
        const float* x = logits + (size_t)r * V;
        float mx = -INFINITY;
        for (int v = lane; v < V; v += 32) mx = fmaxf(mx, x[v]);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        float s = 0.f;
        for (int v = lane; v < V; v += 32) s += __expf(x[v] - mx);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        const float lse = mx + logf(s);
        float kl = 0.f;
        for (int v = lane; v < V; v += 32) {
            const float lp = x[v] - lse;
            kl += __expf(lp) * (lp - (log_prior ? log_prior[v] : qu));
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) kl += __shfl_xor_sync(0xffffffffu, kl, o);
        if (dlogits) {
            float* g = dlogits + (size_t)r * V;
            for (int v = lane; v < V; v += 32) {
                const float lp = x[v] - lse;
                g[v] += w * __expf(lp) * ((lp - (log_prior ? log_prior[v] : qu)) - kl);
            }
        }
        local += kl;
    }
    if (lane == 0 && loss_out) atomicAdd(loss_out, w * local);
}

// hi = bf16(x), lo = bf16(x - hi): x ~= hi + lo to ~16 mantissa bits (split-bf16 GEMMs for weight folding)
__global__ void split_f32_bf16_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ hi,
                                      __nv_bfloat16* __restrict__ lo, size_t n) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const float x = src[i];
        const __nv_bfloat16 h = __float2bfloat16(x);
        hi[i] = h;
        lo[i] = __float2bfloat16(x - __bfloat162float(h));
    }
}

// nnet_input [B,T,D] f32 (zero padded, batch-major)  ->  X0 [T,B,Dp] fp16 (time-major, Dp >= D, pad = 0)
__global__ void pack_input_kernel(const float* __restrict__ x, __half* __restrict__ out, int B, int T, int D, int Dp) {
    const size_t total = (size_t)T * B * Dp;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int d = (int)(i % Dp);
        const size_t tb = i / Dp;
        const int b = (int)(tb % B), t = (int)(tb / B);
        out[i] = (d < D) ? __float2half_rn(sat_f16(x[((size_t)b * T + t) * D + d])) : __float2half_rn(0.f);
    }
}

// column sums of a row-major matrix: out[c] (+)= sum_r src[r, c]   (bias gradients)
template <typename TIn>
__global__ void colsum_kernel(const TIn* __restrict__ src, int rows, int cols, int ld, float* __restrict__ out, int accumulate) {
    __shared__ float sm[8][33];
    const int c = blockIdx.x * 32 + threadIdx.x;
    float acc = 0.f;
    const int r0 = blockIdx.y * 8 + threadIdx.y;
    if (c < cols)
        for (int r = r0; r < rows; r += gridDim.y * 8) {
            if constexpr (sizeof(TIn) == 2) acc += __bfloat162float(src[(size_t)r * ld + c]);
            else acc += src[(size_t)r * ld + c];
        }
    sm[threadIdx.y][threadIdx.x] = acc;
    __syncthreads();
    if (threadIdx.y == 0 && c < cols) {
        float s = 0.f;
#pragma unroll
        for (int k = 0; k < 8; ++k) s += sm[k][threadIdx.x];
        atomicAdd(out + c, s);
    }
    (void)accumulate;
}

static inline int grid_for(size_t n, int per_thread, int threads) {
    size_t b = (n + (size_t)per_thread * threads - 1) / ((size_t)per_thread * threads);
    if (b < 1) b = 1;
    if (b > num_sms() * 16) b = num_sms() * 16;
    return (int)b;
}

}  // namespace lcb

using namespace lcb;

// dst_dtype 1: bf16, 2: fp16 (saturating)
extern "C" int lcb_cast_f32_16(const float* src, void* dst, int dst_dtype, size_t n, void* stream) {
    if (!src || !dst) return LCB_ERR_NULL_POINTER;
    if (dst_dtype != 1 && dst_dtype != 2) return LCB_ERR_BAD_SHAPE;
    if (((uintptr_t)src & 15) || ((uintptr_t)dst & 7)) return LCB_ERR_MISALIGNED;
    if (n == 0) return LCB_OK;
    g_launches += 1;
    if (dst_dtype == 2) cast_f32_16_kernel<true><<<grid_for(n, 4, 256), 256, 0, (cudaStream_t)stream>>>(src, (uint16_t*)dst, n);
    else cast_f32_16_kernel<false><<<grid_for(n, 4, 256), 256, 0, (cudaStream_t)stream>>>(src, (uint16_t*)dst, n);
    return cudaGetLastError() == cudaSuccess ? LCB_OK : LCB_ERR_CUDA;
}

extern "C" int lcb_f16_to_bf16(const void* src, void* dst, size_t n, void* stream) {
    if (!src || !dst) return LCB_ERR_NULL_POINTER;
    if ((n & 1) || ((uintptr_t)src & 3) || ((uintptr_t)dst & 3)) return LCB_ERR_MISALIGNED;
    if (n == 0) return LCB_OK;
    g_launches += 1;
    f16_to_bf16_kernel<<<grid_for(n / 2, 1, 256), 256, 0, (cudaStream_t)stream>>>((const __half2*)src, (__nv_bfloat162*)dst, n / 2);
    return cudaGetLastError() == cudaSuccess ? LCB_OK : LCB_ERR_CUDA;
}

extern "C" int lcb_dropout16(void* x, int dtype, size_t n, float keep_prob, unsigned long long seed, void* stream) {
    if (!x) return LCB_ERR_NULL_POINTER;
    if ((dtype != 1 && dtype != 2) || !(keep_prob > 0.f) || keep_prob > 1.f) return LCB_ERR_BAD_SHAPE;
    if (n == 0 || keep_prob == 1.f) return LCB_OK;
    if ((uintptr_t)x & 15) return LCB_ERR_MISALIGNED;
    g_launches += 1;
    if (dtype == 2) dropout16_kernel<true><<<grid_for(n, 16, 256), 256, 0, (cudaStream_t)stream>>>((uint16_t*)x, n, 1.f / keep_prob, keep_threshold16(keep_prob), seed);
    else dropout16_kernel<false><<<grid_for(n, 16, 256), 256, 0, (cudaStream_t)stream>>>((uint16_t*)x, n, 1.f / keep_prob, keep_threshold16(keep_prob), seed);
    return cudaGetLastError() == cudaSuccess ? LCB_OK : LCB_ERR_CUDA;
}

// the 0/1 mask lcb_dropout16 / the mixture kernels use for (seed, element index) -- for tests
extern "C" int lcb_dropout_mask(unsigned char* mask, size_t n, float keep_prob, unsigned long long seed, void* stream) {
    if (!mask) return LCB_ERR_NULL_POINTER;
    if (n == 0) return LCB_OK;
    g_launches += 1;
    dropout_mask_kernel<<<grid_for(n, 2, 256), 256, 0, (cudaStream_t)stream>>>(mask, n, keep_threshold16(keep_prob), seed);
    return cudaGetLastError() == cudaSuccess ? LCB_OK : LCB_ERR_CUDA;
}

extern "C" int lcb_add_f16(void* y, const void* x, size_t n, void* stream) {
    if (!y || !x) return LCB_ERR_NULL_POINTER;
    if ((n & 1) || ((uintptr_t)y & 3) || ((uintptr_t)x & 3)) return LCB_ERR_MISALIGNED;
    if (n == 0) return LCB_OK;
    g_launches += 1;
    add_f16_kernel<<<grid_for(n / 2, 1, 256), 256, 0, (cudaStream_t)stream>>>((__half2*)y, (const __half2*)x, n / 2);
    return cudaGetLastError() == cudaSuccess ? LCB_OK : LCB_ERR_CUDA;
}

extern "C" int lcb_masked_add16(void* y, int ldy, const void* x, int ldx, long long rows, int cols, int dtype, float keep_prob,
                                unsigned long long seed, unsigned long long mask_base, int ldm, void* stream) {
    if (!y || !x) return LCB_ERR_NULL_POINTER;
    if ((dtype != 1 && dtype != 2) || !(keep_prob > 0.f) || keep_prob > 1.f || rows < 0 || cols < 0 || ldy < cols || ldx < cols) return LCB_ERR_BAD_SHAPE;
    if (rows == 0 || cols == 0) return LCB_OK;
    g_launches += 1;
    const int grid = grid_for((size_t)rows * cols, 1, 256);
    if (dtype == 2) masked_add16_kernel<true><<<grid, 256, 0, (cudaStream_t)stream>>>((uint16_t*)y, ldy, (const uint16_t*)x, ldx, rows, cols, 1.f / keep_prob, keep_threshold16(keep_prob), seed, mask_base, ldm);
    else masked_add16_kernel<false><<<grid, 256, 0, (cudaStream_t)stream>>>((uint16_t*)y, ldy, (const uint16_t*)x, ldx, rows, cols, 1.f / keep_prob, keep_threshold16(keep_prob), seed, mask_base, ldm);
    return cudaGetLastError() == cudaSuccess ? LCB_OK : LCB_ERR_CUDA;
}

// loss_out (device float, caller-initialised) += weight * sum_rows KL-term; dlogits (nullable) += its gradient
extern "C" int lcb_label_smooth(const float* logits, float* dlogits, long long rows, int V, float weight,
                                const float* log_prior, float* loss_out, void* stream) {
    if (!logits) return LCB_ERR_NULL_POINTER;
    if (rows <= 0 || V <= 0) return LCB_ERR_BAD_SHAPE;
    long long blocks = (rows + 7) / 8; if (blocks > num_sms() * 8) blocks = num_sms() * 8;
    g_launches += 1;
    label_smooth_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(logits, dlogits, rows, V, weight, log_prior, loss_out);
    return cudaGetLastError() == cudaSuccess ? LCB_OK : LCB_ERR_CUDA;
}

extern "C" int lcb_split_f32_bf16(const float* src, void* hi, void* lo, size_t n, void* stream) {
    if (!src || !hi || !lo) return LCB_ERR_NULL_POINTER;
    if (n == 0) return LCB_OK;
    g_launches += 1;
    split_f32_bf16_kernel<<<grid_for(n, 1, 256), 256, 0, (cudaStream_t)stream>>>(src, (__nv_bfloat16*)hi, (__nv_bfloat16*)lo, n);
    return cudaGetLastError() == cudaSuccess ? LCB_OK : LCB_ERR_CUDA;
}

extern "C" int lcb_pack_input(const float* nnet_input, void* x0, int B, int T, int D, int Dp, void* stream) {
    if (!nnet_input || !x0) return LCB_ERR_NULL_POINTER;
    if (B <= 0 || T <= 0 || D <= 0 || Dp < D) return LCB_ERR_BAD_SHAPE;
    g_launches += 1;
    pack_input_kernel<<<grid_for((size_t)T * B * Dp, 1, 256), 256, 0, (cudaStream_t)stream>>>(nnet_input, (__half*)x0, B, T, D, Dp);
    return cudaGetLastError() == cudaSuccess ? LCB_OK : LCB_ERR_CUDA;
}

// out[c] += sum_r src[r,c];  src_dtype 0: f32, 1: bf16.  `out` must be initialised by the caller.
extern "C" int lcb_colsum(const void* src, int src_dtype, int rows, int cols, int ld, float* out, void* stream) {
    if (!src || !out) return LCB_ERR_NULL_POINTER;
    if (rows <= 0 || cols <= 0 || ld < cols) return LCB_ERR_BAD_SHAPE;
    dim3 blk(32, 8);
    int gy = (rows + 511) / 512; if (gy > 256) gy = 256; if (gy < 1) gy = 1;
    dim3 grd((cols + 31) / 32, gy);
    g_launches += 1;
    if (src_dtype) colsum_kernel<__nv_bfloat16><<<grd, blk, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)src, rows, cols, ld, out, 1);
    else colsum_kernel<float><<<grd, blk, 0, (cudaStream_t)stream>>>((const float*)src, rows, cols, ld, out, 1);
    return cudaGetLastError() == cudaSuccess ? LCB_OK : LCB_ERR_CUDA;
}
