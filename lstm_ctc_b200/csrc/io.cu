// io.cu -- the data formats either side of the hot path.
//
//   lcb_splice_subsample : frame splicing +-context with edge replication and subsampling of a zero-padded minibatch on the
//                          device (nnet/tfrecord.py:28-51 `_splice` / `_subsample`, applied per utterance by the tf.data map at
//                          :105-113).  One HBM-bound gather: reads each input frame (1+lc+rc)/factor times out of L2, writes
//                          the [B, T', D(1+lc+rc)] tensor the BiLSTM stack consumes.
//   lcb_crc32c           : CRC-32C (Castagnoli) of a HOST buffer, slicing-by-8 -- the checksum of the TFRecord framing
//                          (length, masked crc of length, payload, masked crc of payload) that tf.python_io.TFRecordWriter
//                          (nnet/tfrecord.py:132,155) writes and tf.data.TFRecordDataset (:118) verifies.
#include "ptx.cuh"
#include "lstm_ctc_b200.h"

namespace lcb {

// out[b, t', c*D + d] = in[b, clamp(t'*f + c - lc, 0, len_b - 1), d]   for t' < len_b / f ; 0 elsewhere
__global__ void splice_subsample_kernel(const float* __restrict__ in, const int* __restrict__ lens, float* __restrict__ out,
                                        int* __restrict__ lens_out, int B, int T, int D, int lc, int rc, int f, int Tout)
{
    const int C = 1 + lc + rc;
    const size_t DW = (size_t)C * D;
    const size_t total = (size_t)B * Tout * DW;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int col = (int)(i % DW);
        const size_t bt = i / DW;
        const int tp = (int)(bt % Tout), b = (int)(bt / Tout);
        int len = lens[b]; len = len < 0 ? 0 : (len > T ? T : len);
        const int nout = len / f;
        float v = 0.f;
        if (tp < nout) {
            const int c = col / D, d = col - c * D;
            int t = tp * f + c - lc;
            t = t < 0 ? 0 : (t > len - 1 ? len - 1 : t);
            v = in[((size_t)b * T + t) * D + d];
        }
        out[i] = v;
    }
    if (blockIdx.x == 0 && lens_out)
        for (int b = threadIdx.x; b < B; b += blockDim.x) {
            int len = lens[b]; len = len < 0 ? 0 : (len > T ? T : len);
            lens_out[b] = len / f;
        }
}

static uint32_t g_crc_tab[8][256];
static void crc32c_init() {
    for (uint32_t n = 0; n < 256; ++n) {
        uint32_t c = n;
        for (int k = 0; k < 8; ++k) c = (c & 1u) ? (0x82F63B78u ^ (c >> 1)) : (c >> 1);
        g_crc_tab[0][n] = c;
    }
    for (uint32_t n = 0; n < 256; ++n) {
        uint32_t c = g_crc_tab[0][n];
        for (int k = 1; k < 8; ++k) { c = g_crc_tab[0][c & 0xffu] ^ (c >> 8); g_crc_tab[k][n] = c; }
    }
}

}  // namespace lcb

using namespace lcb;

extern "C" int lcb_splice_subsample(const float* in, const int32_t* lens, float* out, int32_t* lens_out,
                                    int B, int T, int D, int left_context, int right_context, int subsample, int Tout, void* stream)
{
    if (!in || !lens || !out) return LCB_ERR_NULL_POINTER;
    const int f = subsample > 0 ? subsample : 1;
    if (B <= 0 || T <= 0 || D <= 0 || left_context < 0 || right_context < 0 || Tout <= 0 || Tout < T / f) return LCB_ERR_BAD_SHAPE;
    const size_t total = (size_t)B * Tout * (size_t)(1 + left_context + right_context) * D;
    g_launches += 1;
    splice_subsample_kernel<<<grid_for(total, 1, 256), 256, 0, (cudaStream_t)stream>>>(in, lens, out, lens_out, B, T, D,
                                                                                       left_context, right_context, f, Tout);
    return cudaGetLastError() == cudaSuccess ? LCB_OK : LCB_ERR_CUDA;
}

extern "C" uint32_t lcb_crc32c(const void* data, size_t n, uint32_t crc)
{
    static const bool ready = (crc32c_init(), true);      // C++11 static: initialised once, thread-safe (prefetch thread + main thread)
    (void)ready;
    const unsigned char* p = (const unsigned char*)data;
    uint32_t c = crc ^ 0xffffffffu;
    while (n && ((uintptr_t)p & 7)) { c = g_crc_tab[0][(c ^ *p++) & 0xffu] ^ (c >> 8); --n; }
    while (n >= 8) {
        uint64_t w; memcpy(&w, p, 8);
        const uint32_t lo = (uint32_t)w ^ c, hi = (uint32_t)(w >> 32);
        c = g_crc_tab[7][lo & 0xff] ^ g_crc_tab[6][(lo >> 8) & 0xff] ^ g_crc_tab[5][(lo >> 16) & 0xff] ^ g_crc_tab[4][lo >> 24] ^
            g_crc_tab[3][hi & 0xff] ^ g_crc_tab[2][(hi >> 8) & 0xff] ^ g_crc_tab[1][(hi >> 16) & 0xff] ^ g_crc_tab[0][hi >> 24];
        p += 8; n -= 8;
    }
    while (n--) c = g_crc_tab[0][(c ^ *p++) & 0xffu] ^ (c >> 8);
    return c ^ 0xffffffffu;
}
