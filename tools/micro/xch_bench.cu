// xch_bench.cu -- micro-benchmark of the per-step all-gather inside a 16-CTA cluster (design input for lstm_rec.cu):
//   mode 0: 16 bulk DSMEM copies (cp.async.bulk.shared::cluster.shared::cta) per CTA per step           [what v1/v2 do]
//   mode 1: bulk store smem -> L2 scratch, wait_group, then ONE multicast bulk load into all CTAs' smem
//   mode 2: as mode 0 with half-size messages to half of the CTAs (traffic of a cta_group::2 operand split)
//   mode 3: reduce-scatter through L2: 16 bulk stores (one per owner), remote mbarrier arrive, owner bulk-loads 16 tiles
//   mode 4: tagged 16-byte chunks: st.global of the slice (bit 14 of every 16-bit element carries a per-reuse toggle), every CTA
//           polls all NC slices with ld.relaxed.gpu.v4 until each chunk shows the expected tag, copies it to smem -- no TMA, no
//           fence, no store acknowledgement on the chain                                                  [round 2 candidate]
//   mode 5: st.global of the slice + __threadfence + fence.proxy.async, then ONE multicast bulk load (mode 1 without the
//           bulk store's acknowledgement round trip through the TMA unit)
// Every step depends on the previous one (a CTA sends step s+1 only after all of step s has landed), like the recurrence.
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../lstm_ctc_b200/csrc/ptx.cuh"
using namespace lcb;

__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_load_mc(uint32_t dst_cta_addr, const void* gsrc, uint32_t bytes, uint32_t mbar_cta_addr, uint16_t mask) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;"
                 ::"r"(dst_cta_addr), "l"(gsrc), "r"(bytes), "r"(mbar_cta_addr), "h"(mask) : "memory");
}

struct P { unsigned char* scratch; long long* out; int steps; int slice; int mode; int work; };

__device__ __forceinline__ uint4 ldg_relaxed_v4(const void* p) {
    uint4 v;
    asm volatile("ld.relaxed.gpu.global.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void stg_relaxed_v4(void* p, uint4 v) {
    asm volatile("st.relaxed.gpu.global.v4.b32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ void fence_proxy_async_global() { asm volatile("fence.proxy.async.global;" ::: "memory"); }

__global__ void __launch_bounds__(256, 1) xch_kernel(P p)
{
    extern __shared__ __align__(1024) unsigned char sm[];
    const int NC = 16;
    const uint32_t cta = cluster_ctarank();
    const int cid = (int)cluster_id_x();
    unsigned char* rbuf = sm;                                   // [2][NC][slice] receive buffers
    unsigned char* sbuf = sm + 2 * NC * p.slice;                // [2][NC][slice] send staging (mode 3 needs NC tiles)
    const int stiles = p.mode == 3 ? NC : 1;
    uint64_t* bar = reinterpret_cast<uint64_t*>(sbuf + 2 * stiles * p.slice);   // [2] data, [2] signal
    if (threadIdx.x == 0) { mbar_init(&bar[0], 1); mbar_init(&bar[1], 1); mbar_init(&bar[2], NC); mbar_init(&bar[3], NC); fence_mbar_init(); }
    for (int i = threadIdx.x; i < 2 * stiles * p.slice / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(sbuf)[i] = i + cta;
    fence_proxy_async_smem();
    __syncthreads();
    cluster_sync_all();
    unsigned char* my_scr = p.scratch + (size_t)cid * 2 * NC * NC * p.slice;   // [2][NC dst/owner][NC src][slice]
    const uint32_t rb = smem_u32(rbuf), sb = smem_u32(sbuf);
    long long t0 = 0;
    float acc = 0.f;
    for (int s = 0; s < p.steps; ++s) {
        const int par = s & 1;
        if (s == 8 && threadIdx.x == 0) t0 = clock64();
        // "compute": a dependent chain of `work` FMAs on every thread (stands in for tcgen05.ld + gate math)
        for (int i = 0; i < p.work; ++i) acc = acc * 1.0001f + 0.5f;
        __syncthreads();
        if (p.mode == 0 || p.mode == 2) {
            const int nd = p.mode == 0 ? NC : NC / 2;
            const uint32_t bytes = p.mode == 0 ? p.slice : p.slice / 2;
            if (threadIdx.x == 0) mbar_arrive_expect_tx(&bar[par], (uint32_t)(nd * bytes));
            if (threadIdx.x < nd) {
                uint32_t dst = p.mode == 0 ? (cta + threadIdx.x) % NC : (((cta >> 1) + threadIdx.x) % (NC / 2)) * 2 + (cta & 1);
                fence_proxy_async_smem();
                bulk_copy_s2c(mapa_shared(rb + (par * NC + cta) * p.slice, dst), sb + par * stiles * p.slice, bytes, mapa_shared(smem_u32(&bar[par]), dst));
            }
        } else if (p.mode == 1) {
            if (threadIdx.x == 0) {
                mbar_arrive_expect_tx(&bar[par], (uint32_t)(NC * p.slice));
                fence_proxy_async_smem();
                unsigned char* g = my_scr + ((size_t)par * NC + cta) * p.slice;
                bulk_store_s2g(g, sb + par * stiles * p.slice, p.slice);
                bulk_commit();
                bulk_wait0();
                bulk_load_mc(rb + (par * NC + cta) * p.slice, g, p.slice, smem_u32(&bar[par]), (uint16_t)0xffff);
            }
        } else if (p.mode == 4) {
            const uint32_t TB = 0x40004000u;
            const uint32_t tag = (((s >> 1) & 1) ^ 1) ? TB : 0u;
            uint4* gs = reinterpret_cast<uint4*>(my_scr + (size_t)par * NC * p.slice);
            const int cps = p.slice / 16;                             // chunks per slice
            if ((int)threadIdx.x < cps) {
                uint4 v = reinterpret_cast<const uint4*>(sbuf + par * stiles * p.slice)[threadIdx.x];
                v.x = (v.x & ~TB) | tag; v.y = (v.y & ~TB) | tag; v.z = (v.z & ~TB) | tag; v.w = (v.w & ~TB) | tag;
                stg_relaxed_v4(gs + cta * cps + threadIdx.x, v);
            }
            const int nch = NC * cps, per = nch / (int)blockDim.x;   // chunks per thread (<= 8)
            uint32_t pending = (1u << per) - 1u;
            uint4* rb4 = reinterpret_cast<uint4*>(rbuf + (size_t)par * NC * p.slice);
            int guard = 0;
            while (pending && ++guard < 100000) {
                uint4 v[8];
#pragma unroll
                for (int k = 0; k < 8; ++k) if (pending >> k & 1) v[k] = ldg_relaxed_v4(gs + k * blockDim.x + threadIdx.x);
#pragma unroll
                for (int k = 0; k < 8; ++k) if (pending >> k & 1) {
                    const bool okc = (v[k].x & TB) == tag && (v[k].y & TB) == tag && (v[k].z & TB) == tag && (v[k].w & TB) == tag;
                    if (okc) { rb4[k * blockDim.x + threadIdx.x] = v[k]; pending &= ~(1u << k); }
                }
            }
            fence_proxy_async_smem();
            __syncthreads();
            acc += (float)rbuf[(par * NC + (threadIdx.x & 15)) * p.slice + (threadIdx.x >> 4)];
            continue;
        } else if (p.mode == 5) {
            uint4* gs = reinterpret_cast<uint4*>(my_scr + (size_t)par * NC * p.slice);
            const int cps = p.slice / 16;
            if ((int)threadIdx.x < cps) {
                uint4 v = reinterpret_cast<const uint4*>(sbuf + par * stiles * p.slice)[threadIdx.x];
                gs[cta * cps + threadIdx.x] = v;
                __threadfence();
                fence_proxy_async_global();
            }
            __syncthreads();
            if (threadIdx.x == 0) {
                mbar_arrive_expect_tx(&bar[par], (uint32_t)(NC * p.slice));
                bulk_load_mc(rb + (par * NC + cta) * p.slice, gs + cta * cps, p.slice, smem_u32(&bar[par]), (uint16_t)0xffff);
            }
        } else {
            // reduce-scatter through L2: tile for owner o goes to scratch[par][o][cta]; then signal o; o loads [par][o][*]
            if (threadIdx.x < NC) {
                const uint32_t o = (cta + threadIdx.x) % NC;
                fence_proxy_async_smem();
                bulk_store_s2g(my_scr + (((size_t)par * NC + o) * NC + cta) * p.slice, sb + (par * NC + o) * p.slice, p.slice);
                bulk_commit();
                bulk_wait0();
                mbar_arrive_cluster(&bar[2 + par], o);
            }
            if (threadIdx.x == 32) {
                mbar_wait_cluster_acq(&bar[2 + par], (uint32_t)((s >> 1) & 1));
                mbar_arrive_expect_tx(&bar[par], (uint32_t)(NC * p.slice));
                bulk_load_1d(rbuf + (size_t)par * NC * p.slice, my_scr + ((size_t)par * NC + cta) * NC * p.slice, NC * p.slice, &bar[par]);
            }
        }
        // everybody waits for this step's data (acquire: DSMEM / async-proxy deliveries)
        mbar_wait_cluster_acq(&bar[par], (uint32_t)((s >> 1) & 1));
        acc += (float)rbuf[(par * NC + (threadIdx.x & 15)) * p.slice + (threadIdx.x >> 4)];
    }
    if (threadIdx.x == 0) { p.out[blockIdx.x * 2] = clock64() - t0; p.out[blockIdx.x * 2 + 1] = (long long)acc; }
    cluster_sync_all();
}

int main(int argc, char** argv)
{
    const int ncl = argc > 1 ? atoi(argv[1]) : 4;
    const int steps = 1008;
    unsigned char* scratch; long long* out;
    cudaMalloc(&scratch, (size_t)ncl * 2 * 16 * 16 * 4096);
    cudaMemset(scratch, 0, (size_t)ncl * 2 * 16 * 16 * 4096);
    cudaMalloc(&out, ncl * 16 * 2 * sizeof(long long));
    cudaFuncSetAttribute(xch_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    const int nthr = argc > 2 ? atoi(argv[2]) : 256;
    for (int work : {0, 256}) for (int mode : {1, 4, 5}) for (int slice : {512, 1024, 2048}) {
        cudaMemset(scratch, 0, (size_t)ncl * 2 * 16 * 16 * 4096);
        const size_t smem = (size_t)2 * 16 * slice + (size_t)2 * (mode == 3 ? 16 : 1) * slice + 64;
        if (smem > 200 * 1024) continue;
        cudaFuncSetAttribute(xch_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        P p{scratch, out, steps, slice, mode, work};
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(ncl * 16); cfg.blockDim = dim3(nthr); cfg.dynamicSmemBytes = smem;
        cudaLaunchAttribute at[1]; at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = 16; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        cudaError_t e = cudaLaunchKernelEx(&cfg, xch_kernel, p);
        cudaError_t e2 = cudaDeviceSynchronize();
        long long h[2];
        cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
        printf("work %3d mode %d slice %4d B (per-CTA in %2d KB): %7.0f cycles/step   [%s %s]\n", work, mode, slice, 16 * slice / 1024,
               (double)h[0] / (steps - 8), cudaGetErrorString(e), cudaGetErrorString(e2));
    }
    return 0;
}
