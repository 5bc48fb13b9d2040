// pingpong.cu -- one-way latency of SM -> L2 -> SM signalling on B200, by flavour (design input for the recurrence exchange):
// two CTAs on different SMs bounce a 16-byte word through global memory; reported = cycles per round trip / 2.
//   flavour 0: st.relaxed.gpu.v4  + ld.relaxed.gpu.v4 polling
//   flavour 1: st.volatile.v4     + ld.volatile.v4 polling
//   flavour 2: atom.exch.b64 (returns after the L2 performed it) + ld.relaxed.gpu polling
//   flavour 3: red.relaxed.gpu.add.u32 (fire and forget) + ld.relaxed.gpu.b32 polling
//   flavour 4: st.global.cg.v4 (weak) + ld.global.cv.v4 polling
//   flavour 5: plain L2-hit load latency (dependent ld.relaxed.gpu chain, one CTA)
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint4 ld_rel(const void* p) { uint4 v; asm volatile("ld.relaxed.gpu.global.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory"); return v; }
__device__ __forceinline__ void st_rel(void* p, uint4 v) { asm volatile("st.relaxed.gpu.global.v4.b32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory"); }
__device__ __forceinline__ uint4 ld_vol(const void* p) { uint4 v; asm volatile("ld.volatile.global.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory"); return v; }
__device__ __forceinline__ void st_vol(void* p, uint4 v) { asm volatile("st.volatile.global.v4.b32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory"); }
__device__ __forceinline__ uint4 ld_cv(const void* p) { uint4 v; asm volatile("ld.global.cv.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory"); return v; }
__device__ __forceinline__ void st_cg(void* p, uint4 v) { asm volatile("st.global.cg.v4.b32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory"); }
__device__ __forceinline__ uint32_t ld_rel32(const void* p) { uint32_t v; asm volatile("ld.relaxed.gpu.global.b32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v; }

__global__ void pp(unsigned char* buf, long long* out, int flavour, int iters)
{
    if (threadIdx.x != 0) return;
    const int me = blockIdx.x;               // 0 or 1
    uint4* mine = reinterpret_cast<uint4*>(buf + 4096 * me);        // I write here
    uint4* theirs = reinterpret_cast<uint4*>(buf + 4096 * (me ^ 1));  // I poll here
    long long t0 = 0;
    if (flavour == 5) {
        if (me) return;
        uint32_t* p = reinterpret_cast<uint32_t*>(buf);
        uint32_t idx = 0;
        for (int i = 0; i < 64; ++i) idx = ld_rel32(p + idx);
        t0 = clock64();
        for (int i = 0; i < iters; ++i) idx = ld_rel32(p + idx);
        out[0] = (clock64() - t0) / iters; out[1] = idx;
        return;
    }
    auto wait_ge = [&](uint32_t want) {
        int guard = 0;
        while (++guard < (1 << 22)) {
            uint32_t got;
            if (flavour == 0 || flavour == 2) got = ld_rel(theirs).x;
            else if (flavour == 1) got = ld_vol(theirs).x;
            else if (flavour == 3) got = ld_rel32(theirs);
            else got = ld_cv(theirs).x;
            if (got >= want) break;
        }
    };
    auto send = [&](uint32_t i) {
        const uint4 v = make_uint4(i, 0u, 0u, 0u);
        if (flavour == 0) st_rel(mine, v);
        else if (flavour == 1) st_vol(mine, v);
        else if (flavour == 2) atomicExch(reinterpret_cast<unsigned long long*>(mine), (unsigned long long)i);
        else if (flavour == 3) atomicAdd(reinterpret_cast<unsigned int*>(mine), 1u);
        else st_cg(mine, v);
    };
    for (int i = 1; i <= iters + 16; ++i) {
        if (i == 17) t0 = clock64();
        if (me == 0) { send((uint32_t)i); wait_ge((uint32_t)i); }
        else { wait_ge((uint32_t)i); send((uint32_t)i); }
    }
    out[me * 2] = (clock64() - t0) / iters;
}

int main()
{
    unsigned char* buf; long long* out;
    cudaMalloc(&buf, 1 << 16); cudaMalloc(&out, 64);
    const char* names[] = {"st.relaxed.gpu + ld.relaxed.gpu", "st.volatile + ld.volatile", "atom.exch + ld.relaxed.gpu", "red.add + ld.relaxed.gpu.b32", "st.cg + ld.cv", "dependent L2-hit load"};
    for (int f = 0; f < 6; ++f) {
        cudaMemset(buf, 0, 1 << 16); cudaMemset(out, 0, 64);
        pp<<<2, 32>>>(buf, out, f, 2000);
        cudaError_t e = cudaDeviceSynchronize();
        long long h[4]; cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
        printf("%-36s: %lld cycles per %s   [%s]\n", names[f], f == 5 ? h[0] : h[0] / 2, f == 5 ? "load" : "one-way hop (round trip / 2)", cudaGetErrorString(e));
    }
    return 0;
}
