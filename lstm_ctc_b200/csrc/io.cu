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

// ---------------------------------------------------------------------------------------------------------------------------
// tf.train.SequenceExample decoder (HOST code): what tf.parse_single_sequence_example does for the two FixedLenSequenceFeature
// specs of nnet/tfrecord.py:96-106 -- feature_lists["nnet_input"]: one float_list Feature per frame (packed or unpacked floats),
// feature_lists["nnet_target"]: int64_list Features.  Runs without the GIL from the pipeline's decoding threads: the Python
// restatement (tfrecord.parse_sequence_example_py) walks every frame's Feature in the interpreter, ~2.4 ms per 630-frame
// utterance, which capped the batched nnet-forward at a few hundred utterances per second.
namespace lcb {
struct PbSpan { size_t lo, hi; };
static bool pb_varint(const uint8_t* b, size_t hi, size_t& p, uint64_t& v) {
    v = 0;
    for (int shift = 0; shift < 64 && p < hi; shift += 7) {
        const uint8_t c = b[p++];
        v |= (uint64_t)(c & 0x7f) << shift;
        if (!(c & 0x80)) return true;
    }
    return false;
}
// next field of the message in [p, hi): returns false at the end or on a malformed field (ok tells which)
static bool pb_next(const uint8_t* b, size_t hi, size_t& p, int& field, int& wt, uint64_t& val, PbSpan& sp, bool& ok) {
    if (p >= hi) return false;
    uint64_t tag;
    if (!pb_varint(b, hi, p, tag)) { ok = false; return false; }
    field = (int)(tag >> 3); wt = (int)(tag & 7);
    if (wt == 2) {
        uint64_t n;
        if (!pb_varint(b, hi, p, n) || n > hi - p) { ok = false; return false; }
        sp.lo = p; sp.hi = p + (size_t)n; p += (size_t)n;
    } else if (wt == 0) {
        if (!pb_varint(b, hi, p, val)) { ok = false; return false; }
    } else if (wt == 5) {
        if (hi - p < 4) { ok = false; return false; }
        sp.lo = p; sp.hi = p + 4; p += 4;
    } else if (wt == 1) {
        if (hi - p < 8) { ok = false; return false; }
        sp.lo = p; sp.hi = p + 8; p += 8;
    } else { ok = false; return false; }
    return true;
}
}  // namespace lcb

extern "C" int lcb_parse_sequence_example(const void* buf, size_t n, float* x_out, size_t x_cap, int64_t* y_out, size_t y_cap,
                                          long long* rows, long long* cols, long long* num_labels)
{
    using namespace lcb;
    if (!buf || !rows || !cols || !num_labels) return LCB_ERR_NULL_POINTER;
    const uint8_t* b = (const uint8_t*)buf;
    long long R = 0, NY = 0;
    long long Cc = x_out ? *cols : -1;           // fill pass: the caller passes back the row width the sizing pass returned
    if (x_out && Cc < 0) return LCB_ERR_BAD_SHAPE;
    bool ok = true, have_x = false, have_y = false;
    size_t p = 0; int f, wt; uint64_t v; PbSpan sp{0, 0};
    while (pb_next(b, n, p, f, wt, v, sp, ok)) {
        if (f != 2 || wt != 2) continue;                               // context features: unused by the reference
        size_t p1 = sp.lo; int f1, wt1; uint64_t v1; PbSpan s1{0, 0};
        while (pb_next(b, sp.hi, p1, f1, wt1, v1, s1, ok)) {           // FeatureLists.feature_list map entries
            if (f1 != 1 || wt1 != 2) continue;
            PbSpan key{0, 0}, val{0, 0}; bool hk = false, hv = false;
            size_t p2 = s1.lo; int f2, wt2; uint64_t v2; PbSpan s2{0, 0};
            while (pb_next(b, s1.hi, p2, f2, wt2, v2, s2, ok)) {
                if (f2 == 1 && wt2 == 2) { key = s2; hk = true; }
                else if (f2 == 2 && wt2 == 2) { val = s2; hv = true; }
            }
            if (!ok) return LCB_ERR_BAD_SHAPE;
            if (!hk || !hv) continue;
            const size_t kl = key.hi - key.lo;
            const bool is_x = kl == 10 && memcmp(b + key.lo, "nnet_input", 10) == 0;
            const bool is_y = kl == 11 && memcmp(b + key.lo, "nnet_target", 11) == 0;
            if (!is_x && !is_y) continue;
            if (is_y) have_y = true;
            size_t p3 = val.lo; int f3, wt3; uint64_t v3; PbSpan s3{0, 0};
            while (pb_next(b, val.hi, p3, f3, wt3, v3, s3, ok)) {       // FeatureList.feature
                if (f3 != 1 || wt3 != 2) continue;
                long long c_row = 0;
                size_t p4 = s3.lo; int f4, wt4; uint64_t v4; PbSpan s4{0, 0};
                while (pb_next(b, s3.hi, p4, f4, wt4, v4, s4, ok)) {    // Feature: float_list = 2, int64_list = 3
                    if (wt4 != 2) continue;
                    if (is_x && f4 == 2) {
                        size_t p5 = s4.lo; int f5, wt5; uint64_t v5; PbSpan s5{0, 0};
                        while (pb_next(b, s4.hi, p5, f5, wt5, v5, s5, ok)) {
                            if (f5 != 1 || (wt5 != 2 && wt5 != 5)) continue;
                            const size_t cnt = (s5.hi - s5.lo) / 4;       // packed run, or one unpacked value
                            if (x_out) {
                                const size_t at = (size_t)R * (size_t)Cc + (size_t)c_row;
                                if (c_row + (long long)cnt > Cc) return LCB_ERR_BAD_SHAPE;
                                if (at + cnt > x_cap) return LCB_ERR_WORKSPACE_TOO_SMALL;
                                memcpy(x_out + at, b + s5.lo, cnt * 4);
                            }
                            c_row += (long long)cnt;
                        }
                    } else if (is_y && f4 == 3) {
                        size_t p5 = s4.lo; int f5, wt5; uint64_t v5; PbSpan s5{0, 0};
                        while (pb_next(b, s4.hi, p5, f5, wt5, v5, s5, ok)) {
                            if (f5 != 1) continue;
                            if (wt5 == 2) {
                                size_t q = s5.lo;
                                while (q < s5.hi) {
                                    uint64_t x;
                                    if (!pb_varint(b, s5.hi, q, x)) return LCB_ERR_BAD_SHAPE;
                                    if (y_out) { if ((size_t)NY >= y_cap) return LCB_ERR_WORKSPACE_TOO_SMALL; y_out[NY] = (int64_t)x; }
                                    ++NY;
                                }
                            } else if (wt5 == 0) {
                                if (y_out) { if ((size_t)NY >= y_cap) return LCB_ERR_WORKSPACE_TOO_SMALL; y_out[NY] = (int64_t)v5; }
                                ++NY;
                            }
                        }
                    }
                }
                if (is_x) {
                    if (Cc < 0) {
                        Cc = c_row;                                        // the first frame fixes num_cols ...
                    } else if (c_row != Cc) {
                        return LCB_ERR_BAD_SHAPE;                          // ... and every frame must have it (np.stack / TF shape check)
                    }
                    ++R;
                    have_x = true;
                }
            }
            if (!ok) return LCB_ERR_BAD_SHAPE;
        }
        if (!ok) return LCB_ERR_BAD_SHAPE;
    }
    if (!ok) return LCB_ERR_BAD_SHAPE;
    *rows = R; *cols = have_x ? Cc : 0; *num_labels = have_y ? NY : -1;      // -1: no "nnet_target" feature list
    return LCB_OK;
}

