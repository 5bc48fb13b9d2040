// ptx.cuh -- thin inline-PTX wrappers for sm_100a: mbarrier, TMA, tcgen05/TMEM, clusters/DSMEM.
// Everything here is hand-written for Blackwell; nothing falls back to older paths.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace lcb {

// host-side count of kernels launched by this library (bench.py reports it as gpu_launches)
static long long g_launches = 0;

// SM count of the current device (queried once per device, never a literal): grids are sized in multiples of it
static inline int num_sms() {
    static int cache[64];
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) dev = 0;
    if (cache[dev] == 0) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 1;
        cache[dev] = n;
    }
    return cache[dev];
}

// ----------------------------------------------------------------------------------------
// device-side error word: a kernel that times out on a barrier records a code here and bails
// out instead of hanging the GPU.  Host reads it through lcb_device_error().
// ----------------------------------------------------------------------------------------
__device__ int g_dev_error = 0;   // single definition: the library is one translation unit (lcb_all.cu)

enum DevErr : int {
    DEV_OK = 0,
    DEV_ERR_MBAR_TIMEOUT = 101,
    DEV_ERR_CLUSTER_TIMEOUT = 102,
    DEV_ERR_CTC_BAD_LABEL = 201,
};

__device__ __forceinline__ void dev_set_error(int code) { atomicCAS(&g_dev_error, 0, code); }
__device__ __forceinline__ bool dev_has_error() { return *((volatile int*)&g_dev_error) != 0; }

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ uint64_t globaltimer_ns() {
    uint64_t t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ int ld_acquire_gpu_s32(const int* p) {
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_gpu_s32(int* p, int v) {
    asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
// predicated read-only global loads (0 when the predicate is false): one predicated LDG each, no branch, free to be scheduled
__device__ __forceinline__ float ldg_f32_if(const float* p, bool pred) {
    float v = 0.f;
    asm("{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %2, 0;\n\t@q ld.global.nc.f32 %0, [%1];\n\t}\n" : "+f"(v) : "l"(p), "r"((int)pred));
    return v;
}
__device__ __forceinline__ uint2 ldg_u2_if(const uint2* p, bool pred) {
    uint2 v = make_uint2(0u, 0u);
    asm("{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %3, 0;\n\t@q ld.global.nc.v2.u32 {%0, %1}, [%2];\n\t}\n" : "+r"(v.x), "+r"(v.y) : "l"(p), "r"((int)pred));
    return v;
}
__device__ __forceinline__ bool elect_one() {
    uint32_t pred = 0;
    asm volatile(
        "{\n\t.reg .pred P1;\n\telect.sync _|P1, 0xffffffff;\n\tselp.b32 %0, 1, 0, P1;\n\t}\n"
        : "=r"(pred));
    return pred != 0;
}

// ----------------------------------------------------------------------------------------
// mbarrier
// ----------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_mbar_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async_smem() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async_all() {       // generic-proxy writes (incl. DSMEM) -> async proxy
    asm volatile("fence.proxy.async;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
// remote arrive on the same-offset barrier of CTA `cta` in the cluster
__device__ __forceinline__ void mbar_arrive_cluster(uint64_t* bar, uint32_t cta) {
    uint32_t remote;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(smem_u32(bar)), "r"(cta));
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(remote) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred P1;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n\t"
        "selp.b32 %0, 1, 0, P1;\n\t}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
// non-blocking phase test (try_wait may suspend the thread for a system-dependent time; test_wait never does): for polling
// SEVERAL barriers from one thread
__device__ __forceinline__ bool mbar_test_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred P1;\n\t"
        "mbarrier.test_wait.parity.shared::cta.b64 P1, [%1], %2;\n\t"
        "selp.b32 %0, 1, 0, P1;\n\t}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ bool mbar_try_wait_cluster_acq(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred P1;\n\t"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 P1, [%1], %2;\n\t"
        "selp.b32 %0, 1, 0, P1;\n\t}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
// Bounded wait: ~2 s budget, then records DEV_ERR_MBAR_TIMEOUT and returns false so the kernel
// can drain instead of wedging the GPU.
#ifndef LCB_WAIT_TIMEOUT_NS
#define LCB_WAIT_TIMEOUT_NS 2000000000ull
#endif
__device__ __forceinline__ bool mbar_wait(uint64_t* bar, uint32_t parity) {
    if (mbar_try_wait(bar, parity)) return true;
    if (dev_has_error()) return false;          // somebody already timed out: drain fast
    uint64_t t0 = 0;
    uint32_t spins = 0;
    while (!mbar_try_wait(bar, parity)) {
        if ((++spins & 0x3ff) == 0) {
            uint64_t now = globaltimer_ns();
            if (t0 == 0) t0 = now;
            if (now - t0 > LCB_WAIT_TIMEOUT_NS || dev_has_error()) {
                dev_set_error(DEV_ERR_MBAR_TIMEOUT);
                return false;
            }
        }
    }
    return true;
}
__device__ __forceinline__ bool mbar_wait_cluster_acq(uint64_t* bar, uint32_t parity) {
    if (mbar_try_wait_cluster_acq(bar, parity)) return true;
    uint64_t t0 = 0;
    uint32_t spins = 0;
    while (!mbar_try_wait_cluster_acq(bar, parity)) {
        if ((++spins & 0x3ff) == 0) {
            uint64_t now = globaltimer_ns();
            if (t0 == 0) t0 = now;
            if (now - t0 > LCB_WAIT_TIMEOUT_NS || dev_has_error()) {
                dev_set_error(DEV_ERR_MBAR_TIMEOUT);
                return false;
            }
        }
    }
    return true;
}

// ----------------------------------------------------------------------------------------
// TMA (cp.async.bulk.tensor) -- loads complete on an mbarrier via complete_tx
// ----------------------------------------------------------------------------------------
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* m) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}
// TMA tensor store: shared::cta box -> global tensor (rows/cols outside the tensor are clipped), bulk async-group tracked
__device__ __forceinline__ void tma_store_3d(const CUtensorMap* m, uint32_t smem_src, int c0, int c1, int c2) {
    asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.tile.bulk_group [%0, {%2, %3, %4}], [%1];"
                 ::"l"(reinterpret_cast<uint64_t>(m)), "r"(smem_src), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
// 2-D TMA tensor store / fp32 reduce-add: shared::cta box -> global tensor (out-of-range rows/cols are clipped)
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* m, uint32_t smem_src, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.tile.bulk_group [%0, {%2, %3}], [%1];"
                 ::"l"(reinterpret_cast<uint64_t>(m)), "r"(smem_src), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_reduce_add_2d(const CUtensorMap* m, uint32_t smem_src, int c0, int c1) {
    asm volatile("cp.reduce.async.bulk.tensor.2d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3}], [%1];"
                 ::"l"(reinterpret_cast<uint64_t>(m)), "r"(smem_src), "r"(c0), "r"(c1) : "memory");
}
template <int N>
__device__ __forceinline__ void bulk_wait_group_read_pending() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_group_pending() { asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory"); }
// explicit shared-space accesses through 32-bit addresses (no generic-address arithmetic)
__device__ __forceinline__ void sts_b16(uint32_t addr, uint16_t v) { asm volatile("st.shared.b16 [%0], %1;" ::"r"(addr), "h"(v) : "memory"); }
__device__ __forceinline__ float lds_f16(uint32_t addr) {                 // one fp16 from shared memory, widened
    uint16_t h; asm volatile("ld.shared.b16 %0, [%1];" : "=h"(h) : "r"(addr) : "memory");
    float f; asm("cvt.f32.f16 %0, %1;" : "=f"(f) : "h"(h));
    return f;
}
__device__ __forceinline__ uint32_t lds_b32(uint32_t addr) { uint32_t v; asm volatile("ld.shared.b32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory"); return v; }
__device__ __forceinline__ void sts_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
// plain 1-D bulk copy global -> shared::cta (size multiple of 16, 16B aligned)
__device__ __forceinline__ void bulk_load_1d(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
        ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(gsrc)), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}

// ----------------------------------------------------------------------------------------
// cp.async (LDGSTS), per-thread-private staging
// ----------------------------------------------------------------------------------------
template <int BYTES>
__device__ __forceinline__ void cp_async(void* smem_dst, const void* gsrc) {
    static_assert(BYTES == 4 || BYTES == 8 || BYTES == 16, "cp.async size");
    if constexpr (BYTES == 16)
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem_dst)), "l"(gsrc) : "memory");
    else
        asm volatile("cp.async.ca.shared.global [%0], [%1], %2;" ::"r"(smem_u32(smem_dst)), "l"(gsrc), "n"(BYTES) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// ----------------------------------------------------------------------------------------
// tcgen05 / TMEM
// ----------------------------------------------------------------------------------------
template <int NCOLS>
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_result) {   // one full warp
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_result)), "n"(NCOLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
template <int NCOLS>
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr) {        // same warp that allocated
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(NCOLS) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// D[tmem] (+)= A[smem desc] * B[smem desc], bf16/f16 inputs, f32 accumulate; single-thread issue
__device__ __forceinline__ void umma_f16_ss(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// SS form issued from a converged warp by the lane whose `leader` flag is set (see umma_f16_ts_elect); the accumulate flag
// is a template literal (a run-time predicate operand costs ~10 cycles per MMA)
template <bool ACCUMULATE>
__device__ __forceinline__ void umma_f16_ss_elect(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t leader) {
    if (ACCUMULATE)
        asm volatile(
            "{\n\t.reg .pred pe;\n\tsetp.ne.b32 pe, %4, 0;\n\t"
            "@pe tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, 1;\n\t}\n"
            ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(leader) : "memory");
    else
        asm volatile(
            "{\n\t.reg .pred pe;\n\tsetp.ne.b32 pe, %4, 0;\n\t"
            "@pe tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, 0;\n\t}\n"
            ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(leader) : "memory");
}
__device__ __forceinline__ void tma_load_2d_elect(uint32_t smem_dst, const CUtensorMap* m, uint32_t bar, int c0, int c1, uint32_t leader) {
    asm volatile(
        "{\n\t.reg .pred pe;\n\tsetp.ne.b32 pe, %5, 0;\n\t"
        "@pe cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];\n\t}\n"
        ::"r"(smem_dst), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar), "r"(c0), "r"(c1), "r"(leader)
        : "memory");
}
// same, descriptors given as (lo, hi) words: only `lo` (start address field) changes between k-steps, so the
// issue loop is one integer add per operand instead of rebuilding 64-bit descriptors
__device__ __forceinline__ void umma_f16_ss_lohi(uint32_t tmem_d, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi,
                                                 uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
        "mov.b64 da, {%1, %2};\n\tmov.b64 db, {%3, %4};\n\t"
        "setp.ne.b32 p, %6, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n\t}\n"
        ::"r"(tmem_d), "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate)
        : "memory");
}
// TS form: A operand resident in TMEM (lane = row, two 16-bit K elements per 32-bit column), B from smem
__device__ __forceinline__ void umma_f16_ts_lohi(uint32_t tmem_d, uint32_t tmem_a, uint32_t b_lo, uint32_t b_hi,
                                                 uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .b64 db;\n\t"
        "mov.b64 db, {%2, %3};\n\t"
        "setp.ne.b32 p, %5, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], db, %4, p;\n\t}\n"
        ::"r"(tmem_d), "r"(tmem_a), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate)
        : "memory");
}
// The same, issued from a CONVERGED warp by the lane whose `leader` flag is set: when the operands are warp-uniform
// expressions the compiler keeps them in uniform registers, instead of the ELECT / 4 x R2UR / branch waterfall it wraps
// around every MMA issued from inside an `if (lane == 0)` region (~65 cycles per MMA and thread).
template <bool ACCUMULATE = true>
__device__ __forceinline__ void umma_f16_ts_elect(uint32_t tmem_d, uint32_t tmem_a, uint32_t b_lo, uint32_t b_hi,
                                                  uint32_t idesc, uint32_t leader) {
    if (ACCUMULATE)
        asm volatile(
            "{\n\t.reg .pred pe;\n\t.reg .b64 db;\n\t"
            "mov.b64 db, {%2, %3};\n\t"
            "setp.ne.b32 pe, %5, 0;\n\t"
            "@pe tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], db, %4, 1;\n\t}\n"
            ::"r"(tmem_d), "r"(tmem_a), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(leader)
            : "memory");
    else
        asm volatile(
            "{\n\t.reg .pred pe;\n\t.reg .b64 db;\n\t"
            "mov.b64 db, {%2, %3};\n\t"
            "setp.ne.b32 pe, %5, 0;\n\t"
            "@pe tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], db, %4, 0;\n\t}\n"
            ::"r"(tmem_d), "r"(tmem_a), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(leader)
            : "memory");
}
// registers -> TMEM: thread i of the warp writes 8 consecutive 32-bit columns of lane (quarter base + i)
__device__ __forceinline__ void tmem_st_32x32b_x8(uint32_t taddr, const uint32_t (&r)[8]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
                 ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
                 : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
// 16 lanes x 256 bit: thread t of the warp receives rows (t/4, t/4+8) of the 16-lane slab at [taddr],
// columns 2*(t%4), 2*(t%4)+1:  r0,r1 = row t/4;  r2,r3 = row t/4 + 8
__device__ __forceinline__ void tmem_ld_16x256b_x1(uint32_t taddr, uint32_t (&r)[4]) {
    asm volatile("tcgen05.ld.sync.aligned.16x256b.x1.b32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(taddr) : "memory");
}
// mirror of tmem_ld_16x256b_x1: thread t writes rows (t/4, t/4+8), columns 2*(t%4), 2*(t%4)+1 of the 16-lane slab
__device__ __forceinline__ void tmem_st_16x256b_x1(uint32_t taddr, const uint32_t (&r)[4]) {
    asm volatile("tcgen05.st.sync.aligned.16x256b.x1.b32 [%0], {%1, %2, %3, %4};"
                 ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]) : "memory");
}
// commit all prior async tcgen05 ops of this thread to an mbarrier (implies fence::before_thread_sync)
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// 32 lanes x 32 consecutive fp32 columns: thread i of the warp gets row (lane base + i), cols [c, c+32)
__device__ __forceinline__ void tmem_ld_32x32b_x32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld_32x32b_x16(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld_32x32b_x8(uint32_t taddr, uint32_t (&r)[8]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
        : "r"(taddr)
        : "memory");
}

// --- descriptors (layout per cute/arch/mma_sm100_desc.hpp SmemDescriptor / InstrDescriptor) ---
// shared-memory matrix descriptor, SWIZZLE_128B canonical layouts:
//   K-major : rows of 128 B (64 bf16 of K), 8-row swizzle atoms of 1024 B; SBO = 1024 (next 8 rows)
//   MN-major: rows of 128 B (64 bf16 of M/N) indexed by k, 8-k atoms of 1024 B;
//             SBO = 1024 (next 8 k), LBO = byte distance between 64-element MN blocks
__device__ __forceinline__ uint64_t make_smem_desc_sw128(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3fff);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3fff) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3fff) << 32;
    d |= (uint64_t)1 << 46;     // descriptor version = 1 (Blackwell)
    d |= (uint64_t)2 << 61;     // layout_type = SWIZZLE_128B
    return d;
}
// SWIZZLE_NONE ("interleave") K-major canonical layout: 8 rows x 16 bytes core matrices, each a contiguous
// 128-byte block; LBO = byte distance between core matrices adjacent in K, SBO = adjacent in M/N
__device__ __forceinline__ uint64_t make_smem_desc_noswz(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3fff);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3fff) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3fff) << 32;
    d |= (uint64_t)1 << 46;     // descriptor version = 1 (Blackwell); layout_type = 0 (SWIZZLE_NONE)
    return d;
}
// bulk DSMEM copy: own shared memory -> shared memory of another CTA of the cluster, complete_tx on ITS mbarrier
__device__ __forceinline__ void bulk_copy_s2c(uint32_t dst_cluster_addr, uint32_t src_cta_addr, uint32_t bytes, uint32_t mbar_cluster_addr) {
    asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst_cluster_addr), "r"(src_cta_addr), "r"(bytes), "r"(mbar_cluster_addr) : "memory");
}
// bulk store: own shared memory -> global (async proxy), tracked by the thread's bulk async-group
__device__ __forceinline__ void bulk_store_s2g(void* gdst, uint32_t src_cta_addr, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(src_cta_addr), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit_group() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_group_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }        // writes performed
__device__ __forceinline__ void bulk_wait_group_read_all() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }  // sources read
// multicast bulk load: global -> the SAME shared-memory offset of every CTA in cta_mask, complete_tx on the mbarrier at the
// same offset of each of them (the L2 response is replicated by the cluster crossbar: one L2 read, NC deliveries)
__device__ __forceinline__ void bulk_load_multicast(uint32_t dst_cta_addr, const void* gsrc, uint32_t bytes, uint32_t mbar_cta_addr, uint16_t cta_mask) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;"
                 ::"r"(dst_cta_addr), "l"(gsrc), "r"(bytes), "r"(mbar_cta_addr), "h"(cta_mask) : "memory");
}
// instruction descriptor for kind::f16: bf16 x bf16 -> f32
__host__ __device__ constexpr uint32_t make_idesc_bf16_f32(int M, int N, int a_mn_major, int b_mn_major) {
    return (1u << 4)                       // c_format = F32
           | (1u << 7)                     // a_format = BF16
           | (1u << 10)                    // b_format = BF16
           | ((uint32_t)a_mn_major << 15)  // a_major (0 = K)
           | ((uint32_t)b_mn_major << 16)  // b_major
           | ((uint32_t)(N >> 3) << 17)    // n_dim
           | ((uint32_t)(M >> 4) << 24);   // m_dim
}

// ----------------------------------------------------------------------------------------
// clusters / DSMEM
// ----------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ uint32_t cluster_nctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ uint32_t cluster_id_x() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%clusterid.x;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_arrive_release() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
__device__ __forceinline__ void cluster_wait_acquire() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }
__device__ __forceinline__ void cluster_sync_all() { cluster_arrive_release(); cluster_wait_acquire(); }
__device__ __forceinline__ uint32_t mapa_shared(uint32_t local_smem_addr, uint32_t cta) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_smem_addr), "r"(cta));
    return r;
}
__device__ __forceinline__ void st_cluster_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    asm volatile("st.shared::cluster.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ void st_cluster_v2(uint32_t addr, uint32_t a, uint32_t b) {
    asm volatile("st.shared::cluster.v2.b32 [%0], {%1, %2};" ::"r"(addr), "r"(a), "r"(b) : "memory");
}
// asynchronous DSMEM store: lands in CTA-remote shared memory and performs complete_tx(16) on the mbarrier
// (of the SAME destination CTA) when it does -- no release fence / barrier.cluster needed on the sender
__device__ __forceinline__ void st_async_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d, uint32_t mbar) {
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.b32 [%0], {%1, %2, %3, %4}, [%5];"
                 ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d), "r"(mbar) : "memory");
}
__device__ __forceinline__ void st_cluster_b32(uint32_t addr, uint32_t a) {
    asm volatile("st.shared::cluster.b32 [%0], %1;" ::"r"(addr), "r"(a) : "memory");
}

// ----------------------------------------------------------------------------------------
// counter-based dropout RNG: keep(seed, idx) is a pure function, so the backward pass regenerates the
// forward mask instead of storing it (TF's DropoutWrapper / tf.nn.dropout streams cannot be matched
// bit-for-bit anyway; parity tests export this mask and feed it to the oracle).
// ----------------------------------------------------------------------------------------
__host__ __device__ __forceinline__ uint32_t rng_u32(uint64_t seed, uint64_t idx) {
    uint64_t z = seed + 0x9E3779B97F4A7C15ull * (idx + 1);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    z = z ^ (z >> 31);
    return (uint32_t)(z >> 32);
}
__host__ __device__ __forceinline__ uint32_t keep_threshold(float keep) {      // P(u32 < thr) = keep
    double t = (double)keep * 4294967296.0;
    return t >= 4294967295.0 ? 0xffffffffu : (uint32_t)t;
}
__host__ __device__ __forceinline__ bool rng_keep(uint64_t seed, uint64_t idx, uint32_t thr) {
    return thr == 0xffffffffu || rng_u32(seed, idx) < thr;
}
// cheaper stream for the big layer-output dropouts: ONE 64-bit hash decides four consecutive elements (16 bits each,
// keep-probability granularity 2^-16)
__host__ __device__ __forceinline__ uint64_t rng_u64(uint64_t seed, uint64_t idx) {
    uint64_t z = seed + 0x9E3779B97F4A7C15ull * (idx + 1);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
__host__ __device__ __forceinline__ uint32_t keep_threshold16(float keep) {    // P(u16 < thr) = keep; 65536 = keep everything
    double t = (double)keep * 65536.0 + 0.5;
    return t >= 65536.0 ? 65536u : (uint32_t)t;
}
__host__ __device__ __forceinline__ bool rng_keep16(uint64_t word, int e, uint32_t thr16) {   // element e (0..3) of its 4-block
    return (uint32_t)((word >> (16 * e)) & 0xffffu) < thr16;
}
__host__ __device__ __forceinline__ bool rng_keepq(uint64_t seed, uint64_t idx, uint32_t thr16) {   // same stream, one element
    return thr16 >= 65536u || rng_keep16(rng_u64(seed, idx >> 2), (int)(idx & 3), thr16);
}

// ----------------------------------------------------------------------------------------
// misc math
// ----------------------------------------------------------------------------------------
__device__ __forceinline__ float sigmoidf_fast(float x) { return __fdividef(1.0f, 1.0f + __expf(-x)); }
__device__ __forceinline__ float tanhf_fast(float x) {
    // tanh(x) = 2*sigmoid(2x) - 1 ; accurate to ~1e-6 abs with __expf
    return __fdividef(2.0f, 1.0f + __expf(-2.0f * x)) - 1.0f;
}
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
    __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&v);
}

}  // namespace lcb
