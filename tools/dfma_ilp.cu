// dfma_ilp.cu -- FP64 FMA throughput as a function of warps per scheduler and independent chains per thread.
#include <cstdio>
#include <cuda_runtime.h>
template <int ILP> __global__ void k(double *out, int iters) {
    double a[ILP];
#pragma unroll
    for (int i = 0; i < ILP; i++) a[i] = threadIdx.x * 1e-9 + i;
    const double m = 1.0000001, c = 1e-7;
    for (int it = 0; it < iters; it++)
#pragma unroll
        for (int i = 0; i < ILP; i++) a[i] = fma(a[i], m, c);
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; i++) s += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int ILP> void run(int threads, double *d) {
    const int iters = 1 << 15, blocks = 148;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<ILP><<<blocks, threads>>>(d, iters);
    cudaEventRecord(e0); k<ILP><<<blocks, threads>>>(d, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    printf("warps/SM %2d  ILP %2d : %6.2f TFLOP/s\n", threads / 32, ILP, 2.0 * blocks * threads * (double)iters * ILP / (ms * 1e-3) / 1e12);
}
int main() {
    double *d; cudaMalloc(&d, sizeof(double) * 148 * 1024);
    for (int th : {128, 256, 512, 1024}) { run<1>(th, d); run<2>(th, d); run<4>(th, d); run<8>(th, d); run<16>(th, d); }
    return 0;
}
