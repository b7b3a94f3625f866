// tmem_rmw_probe.cu -- how expensive is a read-modify-write sweep over TMEM-resident data, pattern by pattern?
// (The FAST phase-1 kernel keeps its RGSW sums in tensor memory and sweeps them once per gadget digit; knock-out timing showed
// that sweep to be 42 % of the kernel although it moves little data.)  8 warps per CTA (2 per scheduler), one CTA per SM, each warp
// sweeps 128 columns (8 chunks of 16) of its lane quadrant ITERS times; prints cycles per chunk and warp.
//   V0  kernel pattern: wait::ld(c) ; issue ld(c+1) ; 16 FMA ; st(c)
//   V1  same, store delayed by one iteration (st(c-1) issued after ld(c+1))
//   V2  loads only (prefetched), no store
//   V3  stores only
//   V4  x32 chunks: wait ; ld32(b+1) ; 32 FMA ; st32(b)
//   V5  all 8 loads, one wait, all arithmetic, all 8 stores
//   V6  V0 with wait::st after every store
//   V7  no TMEM at all (arithmetic only)
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 tools/tmem_rmw_probe.cu -o tools/tmem_rmw_probe
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define LD16(addr, v) asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];" \
    : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]) : "r"(addr) : "memory")
#define ST16(addr, v) asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" \
    :: "r"(addr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]) : "memory")
#define PIN16(v) asm volatile("" : "+r"(v[0]), "+r"(v[1]), "+r"(v[2]), "+r"(v[3]), "+r"(v[4]), "+r"(v[5]), "+r"(v[6]), "+r"(v[7]), "+r"(v[8]), "+r"(v[9]), "+r"(v[10]), "+r"(v[11]), "+r"(v[12]), "+r"(v[13]), "+r"(v[14]), "+r"(v[15]) :: "memory")
__device__ __forceinline__ void wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// 16 words = 8 doubles: v <- v * k + x (one dependent FMA per double, like a complex multiply-accumulate's depth of 2 when FMAS = 2)
template <int FMAS> __device__ __forceinline__ void work(uint32_t (&v)[16], double k, double x) {
#pragma unroll
    for (int i = 0; i < 8; i++) {
        double d = __hiloint2double((int)v[2 * i + 1], (int)v[2 * i]);
#pragma unroll
        for (int f = 0; f < FMAS; f++) d = fma(d, k, x);
        v[2 * i] = (uint32_t)__double2loint(d); v[2 * i + 1] = (uint32_t)__double2hiint(d);
    }
}

template <int V> __global__ void __launch_bounds__(256, 1) k(long long *cyc, double *sink, int iters) {
    __shared__ uint32_t base_s;
    const int warp = threadIdx.x / 32;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" :: "r"((uint32_t)__cvta_generic_to_shared(&base_s)));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t tm = base_s + ((uint32_t)(32 * (warp & 3)) << 16) + 256u * (uint32_t)(warp >> 2);
    uint32_t z[16];
#pragma unroll
    for (int i = 0; i < 16; i++) z[i] = (i & 1) ? 0x3ff00000u : (uint32_t)threadIdx.x;
    for (int c = 0; c < 16; c++) ST16(tm + 16 * c, z);
    wait_st();
    const double kk = 1.0000001, xx = 1e-9 * threadIdx.x;
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
        if (V == 0 || V == 6) {
            uint32_t v[2][16];
            LD16(tm, v[0]);
#pragma unroll
            for (int c = 0; c < 8; c++) {
                wait_ld(); PIN16(v[c & 1]);
                if (c < 7) LD16(tm + 16 * (c + 1), v[(c + 1) & 1]);
                work<2>(v[c & 1], kk, xx);
                ST16(tm + 16 * c, v[c & 1]);
                if (V == 6) wait_st();
            }
            wait_st();
        } else if (V == 1) {
            uint32_t v[3][16];
            LD16(tm, v[0]);
#pragma unroll
            for (int c = 0; c < 8; c++) {
                wait_ld(); PIN16(v[c % 3]);
                if (c < 7) LD16(tm + 16 * (c + 1), v[(c + 1) % 3]);
                if (c > 0) ST16(tm + 16 * (c - 1), v[(c - 1) % 3]);
                work<2>(v[c % 3], kk, xx);
            }
            ST16(tm + 16 * 7, v[7 % 3]);
            wait_st();
        } else if (V == 2) {
            uint32_t v[2][16];
            LD16(tm, v[0]);
#pragma unroll
            for (int c = 0; c < 8; c++) {
                wait_ld(); PIN16(v[c & 1]);
                if (c < 7) LD16(tm + 16 * (c + 1), v[(c + 1) & 1]);
                work<2>(v[c & 1], kk, xx);
                z[c & 15] ^= v[c & 1][3];
            }
        } else if (V == 3) {
#pragma unroll
            for (int c = 0; c < 8; c++) { work<2>(z, kk, xx); ST16(tm + 16 * c, z); }
            wait_st();
        } else if (V == 5) {
            uint32_t v[8][16];
#pragma unroll
            for (int c = 0; c < 8; c++) LD16(tm + 16 * c, v[c]);
            wait_ld();
#pragma unroll
            for (int c = 0; c < 8; c++) { PIN16(v[c]); work<2>(v[c], kk, xx); ST16(tm + 16 * c, v[c]); }
            wait_st();
        } else if (V == 7) {
#pragma unroll
            for (int c = 0; c < 8; c++) work<2>(z, kk, xx);
        }
    }
    const long long t1 = clock64();
    wait_ld(); wait_st();
    double s = 0;
#pragma unroll
    for (int i = 0; i < 16; i++) s += z[i];
    sink[blockIdx.x * 256 + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) cyc[V] = t1 - t0;
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" :: "r"(base_s));
}
template <int V> void run(long long *cyc, double *sink, const char *what) {
    const int iters = 2000;
    k<V><<<148, 256>>>(cyc, sink, iters);
    cudaError_t e = cudaDeviceSynchronize();
    long long h[8];
    cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    printf("V%d %-62s %7.1f cycles per 16-column chunk and warp  (%s)\n", V, what, (double)h[V] / iters / 8, cudaGetErrorString(e));
}
int main() {
    long long *cyc; double *sink;
    cudaMalloc(&cyc, 64); cudaMalloc(&sink, sizeof(double) * 148 * 256);
    run<7>(cyc, sink, "arithmetic only (16 dependent-pair DFMA per chunk)");
    run<2>(cyc, sink, "loads only, next chunk prefetched");
    run<3>(cyc, sink, "stores only");
    run<0>(cyc, sink, "wait(c); ld(c+1); fma; st(c)   [kernel pattern]");
    run<1>(cyc, sink, "wait(c); ld(c+1); st(c-1); fma [store delayed one iteration]");
    run<6>(cyc, sink, "kernel pattern + wait::st after every store");
    run<5>(cyc, sink, "8 loads, one wait, fma, 8 stores");
    return 0;
}
