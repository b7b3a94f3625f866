// common.cuh -- shared device helpers: torus arithmetic (bit-exact integer semantics of the reference),
// strict complex arithmetic, shared-memory padding.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define MK_THREADS 256            // every blind-rotation CTA
#define MK_MAXL 16                // largest gadget length in params.jl (KMS32 l_uni)

typedef double2 cplx;

// ---- strict complex arithmetic: Julia Complex{Float64} `*`, `+`, `-` with no contraction
// (file is compiled with --fmad=false; the intrinsics make the intent explicit).
__device__ __forceinline__ cplx cmul_s(cplx z, cplx w) {
    return make_double2(__dsub_rn(__dmul_rn(z.x, w.x), __dmul_rn(z.y, w.y)),
                        __dadd_rn(__dmul_rn(z.x, w.y), __dmul_rn(z.y, w.x)));
}
__device__ __forceinline__ cplx cadd_s(cplx a, cplx b) { return make_double2(__dadd_rn(a.x, b.x), __dadd_rn(a.y, b.y)); }
__device__ __forceinline__ cplx csub_s(cplx a, cplx b) { return make_double2(__dsub_rn(a.x, b.x), __dsub_rn(a.y, b.y)); }

// ---- fast complex arithmetic (explicit FMA)
__device__ __forceinline__ cplx cmul_f(cplx z, cplx w) {
    return make_double2(fma(z.x, w.x, -(z.y * w.y)), fma(z.x, w.y, z.y * w.x));
}
__device__ __forceinline__ cplx cmac_f(cplx acc, cplx z, cplx w) {      // acc + z*w, 4 FMA
    acc.x = fma(z.x, w.x, acc.x); acc.x = fma(-z.y, w.y, acc.x);
    acc.y = fma(z.x, w.y, acc.y); acc.y = fma(z.y, w.x, acc.y);
    return acc;
}

// ---- torus traits
template <class T> struct Torus;
template <> struct Torus<uint32_t> {
    typedef int32_t S;
    static constexpr int BITS = 32;
    // native (arithmetic.jl:1-4)
    static __device__ __forceinline__ uint32_t native(double x) {
        x = __dsub_rn(x, __dmul_rn(floor(__dmul_rn(x, 2.3283064365386963e-10)), 4.294967296e9));
        return x == 4.294967296e9 ? 0u : __double2uint_rz(x);
    }
    static __device__ __forceinline__ double to_double(uint32_t v) { return (double)(int32_t)v; }
};
template <> struct Torus<uint64_t> {
    typedef int64_t S;
    static constexpr int BITS = 64;
    // native (arithmetic.jl:6-9)
    static __device__ __forceinline__ uint64_t native(double x) {
        x = __dsub_rn(x, __dmul_rn(floor(__dmul_rn(x, 5.421010862427522e-20)), 1.8446744073709552e19));
        return x == 1.8446744073709552e19 ? 0ull : __double2ull_rz(x);
    }
    static __device__ __forceinline__ double to_double(uint64_t v) { return __ll2double_rn((long long)v); }
};

// divbits (arithmetic.jl:23-27): round-to-nearest shift, not reduced; shift >= width gives 0 in Julia.
template <class T> __device__ __forceinline__ T divbits(T a, int bit) {
    if (bit == 0) return a;
    constexpr int W = sizeof(T) * 8;
    const T carry = (T)((T)(a << (W - bit)) >> (W - 1));
    return (T)((T)(a >> bit) + carry);
}

// Digit j (0 = most significant) of the balanced base-2^logB decomposition
// (gsw.jl:86-96 = gsw.jl:98-110 = unienc.jl:4-18), returned as a signed value in [-B/2, B/2).
// The carry chain runs from the least significant digit up, so digit j needs the digits below it.
template <class T> __device__ __forceinline__ int32_t gadget_digit(T a, int j, int l, int logB) {
    const T mask = (T)(((T)1 << logB) - 1);
    T a1 = divbits<T>(a, (int)(sizeof(T) * 8) - l * logB);
    T d = 0;
    for (int i = l - 1; i >= j; i--) {
        d = a1 & mask;
        a1 >>= logB;
        a1 = (T)(a1 + (d >> (logB - 1)));
    }
    // d is the unbalanced digit j with the carry from below already folded in
    const int32_t sd = (int32_t)d;
    return sd - (int32_t)(((uint32_t)d >> (logB - 1)) << logB);
}

// Shared-memory complex buffers are padded by one element per 8 so that the stride-8 / stride-2
// butterfly passes stay bank-conflict free for 16-byte accesses.
__host__ __device__ __forceinline__ constexpr int PAD(int i) { return i + (i >> 3); }
__host__ __device__ constexpr int padded_len(int H) { return H + (H >> 3); }

struct FftTables {
    const cplx *psi, *psiinv, *roots, *rootsinv;     // reference tables (fft.jl:26-44), H entries each
};
