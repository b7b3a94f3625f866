// tmem_probe.cu -- measures tcgen05.ld / tcgen05.st (32x32b) semantics and throughput for using TMEM as a
// lane-private scratchpad (accumulator storage of the FAST phase-1 kernel).  Build: nvcc -arch=sm_100a.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ void tmem_st16(uint32_t addr, const uint32_t (&v)[16]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
                 :: "r"(addr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]),
                    "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t addr, uint32_t (&v)[16]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                   "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                 : "r"(addr) : "memory");
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

__global__ void __launch_bounds__(256, 1) k_probe(uint32_t *err, long long *cyc, int iters) {
    __shared__ uint32_t base_s;
    const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" :: "r"((uint32_t)__cvta_generic_to_shared(&base_s)));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t base = base_s;
    // warp w owns lanes 32*(w%4).. and columns 256*(w/4) .. +255
    const uint32_t my = base + ((uint32_t)(32 * (warp % 4)) << 16) + 256u * (warp / 4);
    uint32_t v[16], r[16];
    uint32_t bad = 0;
    // correctness: write a pattern to all 256 columns, read back
    for (int c = 0; c < 256; c += 16) {
#pragma unroll
        for (int i = 0; i < 16; i++) v[i] = (threadIdx.x << 16) ^ (blockIdx.x << 28) ^ (c + i) * 2654435761u;
        tmem_st16(my + c, v);
    }
    tmem_wait_st();
    for (int c = 0; c < 256; c += 16) {
        tmem_ld16(my + c, r);
        tmem_wait_ld();
#pragma unroll
        for (int i = 0; i < 16; i++) bad += r[i] != ((threadIdx.x << 16) ^ (blockIdx.x << 28) ^ (c + i) * 2654435761u);
    }
    // read-modify-write dependency through TMEM without an explicit wait::st between st and ld of the same columns
    for (int it = 0; it < 4; it++) {
        tmem_ld16(my, r); tmem_wait_ld();
#pragma unroll
        for (int i = 0; i < 16; i++) r[i] += 1;
        tmem_st16(my, r);
        tmem_wait_st();
    }
    tmem_ld16(my, r); tmem_wait_ld();
#pragma unroll
    for (int i = 0; i < 16; i++) bad += r[i] != (((threadIdx.x << 16) ^ (blockIdx.x << 28) ^ (i) * 2654435761u) + 4);
    __syncthreads();
    // throughput: all 8 warps, loads only
    long long t0 = clock64();
    uint32_t acc = 0;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int c = 0; c < 256; c += 16) { tmem_ld16(my + c, r); tmem_wait_ld(); acc += r[0] ^ r[15]; }
    }
    __syncthreads();
    long long t1 = clock64();
    // loads, 4 in flight before each wait
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int c = 0; c < 256; c += 64) {
            uint32_t a0[16], a1[16], a2[16], a3[16];
            tmem_ld16(my + c, a0); tmem_ld16(my + c + 16, a1); tmem_ld16(my + c + 32, a2); tmem_ld16(my + c + 48, a3);
            tmem_wait_ld();
            acc += a0[0] ^ a1[1] ^ a2[2] ^ a3[3];
        }
    }
    __syncthreads();
    long long t2 = clock64();
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int c = 0; c < 256; c += 16) { tmem_st16(my + c, v); }
        tmem_wait_st();
    }
    __syncthreads();
    long long t3 = clock64();
    // single warp latency: dependent ld chain
    long long t4 = t3, t5 = t3;
    if (warp == 0) {
        t4 = clock64();
        uint32_t a = 0;
        for (int it = 0; it < 64; it++) { tmem_ld16(my + (a & 0xF0), r); tmem_wait_ld(); a = r[0] & 0; }
        t5 = clock64();
        acc += a;
    }
    if (threadIdx.x == 0 && blockIdx.x == 0) { cyc[0] = t1 - t0; cyc[1] = t2 - t1; cyc[2] = t3 - t2; cyc[3] = t5 - t4; }
    if (bad || acc == 0x12345) atomicAdd(err, bad + (acc == 0x12345));
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" :: "r"(base));
}

int main() {
    uint32_t *d_err; long long *d_cyc;
    cudaMalloc(&d_err, 4); cudaMalloc(&d_cyc, 64); cudaMemset(d_err, 0, 4);
    const int iters = 200;
    k_probe<<<148, 256>>>(d_err, d_cyc, iters);
    cudaError_t e = cudaDeviceSynchronize();
    uint32_t err; long long cyc[4];
    cudaMemcpy(&err, d_err, 4, cudaMemcpyDeviceToHost); cudaMemcpy(cyc, d_cyc, 32, cudaMemcpyDeviceToHost);
    printf("status=%s mismatches=%u\n", cudaGetErrorString(e), err);
    const double bytes = 8.0 * 32 * 256 * 4 * iters;     // per SM: 8 warps x 32 lanes x 256 cols x 4 B
    printf("ld (wait each x16): %.1f B/clk/SM (%lld clk)\n", bytes / cyc[0], cyc[0]);
    printf("ld (4 in flight)  : %.1f B/clk/SM (%lld clk)\n", bytes / cyc[1], cyc[1]);
    printf("st                : %.1f B/clk/SM (%lld clk)\n", bytes / cyc[2], cyc[2]);
    printf("ld latency (1 warp, x16 + wait): %.1f clk\n", cyc[3] / 64.0);
    return err != 0;
}
