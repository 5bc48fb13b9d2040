// lcb_all.cu -- unity translation unit for liblstm_ctc_b200.so (one TU so the device-side error
// word and the inline-PTX helpers are shared without relocatable device code).
#include "gemm.cu"
#include "ctc.cu"
#include "lstm_rec.cu"
#include "elementwise.cu"
#include "mos.cu"
#include "optim.cu"
#include "decode.cu"
#include "io.cu"
