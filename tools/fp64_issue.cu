// fp64_issue.cu -- what bounds a DFMA stream on one SM sub-partition: the FP64 pipe (one warp instruction per 2 cycles) or the
// register-file read ports (distinct even / odd source registers per instruction)?  Measures TFLOP/s for
//   V0  a = fma(a, m, c)      one new register source per instruction (m, c shared by all chains)   <- the usual "peak" loop
//   V1  a = fma(b, c, a)      three distinct register sources per instruction                         <- complex multiply-accumulate
//   V2  a = fma(b, K, a)      K in the constant bank: two register sources                             <- butterflies with constant twiddles
//   V3  a = a + b             DADD, two register sources
// at 1..4 warps per scheduler and 8 / 16 independent chains per thread.     nvcc -arch=sm_100a -O3 tools/fp64_issue.cu -o tools/fp64_issue
#include <cstdio>
#include <cuda_runtime.h>
__constant__ double cK[16];
template <int V, int ILP> __global__ void k(double *out, int iters) {
    double a[ILP], b[ILP], c[ILP];
#pragma unroll
    for (int i = 0; i < ILP; i++) { a[i] = threadIdx.x * 1e-9 + i; b[i] = 1.0 + (threadIdx.x + i) * 1e-12; c[i] = 1e-7 * (i + 1) + threadIdx.x * 1e-15; }
    const double m = 1.0000001 + threadIdx.x * 1e-13, cc = 1e-7;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < ILP; i++) {
            if (V == 0) a[i] = fma(a[i], m, cc);
            if (V == 1) a[i] = fma(b[i], c[i], a[i]);
            if (V == 2) a[i] = fma(b[i], cK[i & 15], a[i]);
            if (V == 3) a[i] = a[i] + b[i];
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; i++) s += a[i] + b[i] + c[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int V, int ILP> void run(int threads, double *d) {
    const int iters = 1 << 14, blocks = 148;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<V, ILP><<<blocks, threads>>>(d, iters);
    cudaEventRecord(e0); k<V, ILP><<<blocks, threads>>>(d, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double inst = (double)blocks * threads * iters * ILP;
    printf("V%d warps/scheduler %d ILP %2d : %6.2f T inst-flops/s (x2 for FMA)  %.3f cycles/warp-inst/scheduler @1.965GHz\n", V, threads / 128, ILP,
           (V == 3 ? 1.0 : 2.0) * inst / (ms * 1e-3) / 1e12, (ms * 1e-3 * 1.965e9) / (inst / 32 / (148 * 4)));
}
template <int V> void all(double *d) {
    for (int th : {128, 256, 384, 512}) { run<V, 8>(th, d); run<V, 16>(th, d); }
}
int main() {
    double h[16]; for (int i = 0; i < 16; i++) h[i] = 1e-7 * (i + 1);
    cudaMemcpyToSymbol(cK, h, sizeof(h));
    double *d; cudaMalloc(&d, sizeof(double) * 148 * 1024);
    all<0>(d); all<1>(d); all<2>(d); all<3>(d);
    return 0;
}
