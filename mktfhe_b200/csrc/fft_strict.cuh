// fft_strict.cuh -- the reference's twisted negacyclic transform, butterfly for butterfly.
//
//   forward  fftto!  + fft!  : /root/reference/src/ring/fft.jl:57-63, 105-155  (Cooley-Tukey, bit-reversed Psi)
//   inverse  ifftto! + ifft! : /root/reference/src/ring/fft.jl:74-81, 159-210  (Gentleman-Sande, Psiinv)
//
// A group of H/8 threads transforms one polynomial held in a padded shared-memory buffer; every thread
// owns 8 points per pass and runs up to three consecutive stages in registers (the reference's stages
// are unchanged, only regrouped, so every butterfly sees the same operands, the same twiddle and the
// same operation order: results are bit-identical to the scalar loop nest).
// Groups of one CTA run in lock step; the passes are separated by __syncthreads().
#pragma once
#include "common.cuh"

// One pass = NST consecutive stages whose smallest span is K3.  t = thread index inside the group.
// INV = false: forward butterflies (t, u*w) -> (t+u, t-u), stages in decreasing span.
// INV = true : inverse butterflies (t, u) -> (t+u, (t-u)*w), stages in increasing span.
template <int H, int K3, int NST, bool INV>
__device__ __forceinline__ void fft_pass_regs(cplx (&x)[8], const cplx *__restrict__ tw, int t) {
    constexpr int R = 1 << NST, NSG = 8 / R;
#pragma unroll
    for (int u = 0; u < NSG; u++) {
        const int sg = t * NSG + u;
        const int gg = sg / K3;
#pragma unroll
        for (int ss = 0; ss < NST; ss++) {
            const int s = INV ? (NST - 1 - ss) : ss;
            const int half = R >> (s + 1);                    // partner distance in units of K3
            const int mS = H / (2 * K3 * half);               // groups at this stage = first twiddle index
#pragma unroll
            for (int m = 0; m < R; m++) {
                if (m & half) continue;
                const cplx w = __ldg(&tw[mS + (gg << s) + (m >> (NST - s))]);
                cplx &a = x[u * R + m], &b = x[u * R + m + half];
                if (!INV) {
                    const cplx t0 = a, u0 = cmul_s(b, w);
                    a = cadd_s(t0, u0);
                    b = csub_s(t0, u0);
                } else {
                    const cplx t0 = a, u0 = b;
                    a = cadd_s(t0, u0);
                    b = cmul_s(csub_s(t0, u0), w);
                }
            }
        }
    }
}

// Index of point m (0..7) of thread t in a pass: sub-group sg covers base + q*K3, q < R.
template <int K3, int NST> __device__ __forceinline__ int pass_index(int t, int m) {
    constexpr int R = 1 << NST, NSG = 8 / R;
    const int u = m / R, q = m % R;
    const int sg = t * NSG + u;
    return (sg / K3) * R * K3 + (sg % K3) + q * K3;
}

template <int H, int K3, int NST, bool INV>
__device__ __forceinline__ void fft_pass_smem(cplx *buf, const cplx *__restrict__ tw, int t) {
    cplx x[8];
#pragma unroll
    for (int m = 0; m < 8; m++) x[m] = buf[PAD(pass_index<K3, NST>(t, m))];
    fft_pass_regs<H, K3, NST, INV>(x, tw, t);
#pragma unroll
    for (int m = 0; m < 8; m++) buf[PAD(pass_index<K3, NST>(t, m))] = x[m];
}

// Pass schedules.  H = 1024: spans (512,256,128) (64,32,16) (8,4,2) (1).  H = 512: (256,128,64) (32,16,8) (4,2,1).
template <int H> struct Sched;
template <> struct Sched<1024> { static constexpr int NP = 4; };
template <> struct Sched<512> { static constexpr int NP = 3; };

// Forward transform of the polynomial whose twisted coefficients `load(i)` returns for point i
// (i < H): load must return (signed(p[i]) - im*signed(p[i+H])) * roots[i]  (fft.jl:60).
// active = false lets idle groups take part in the barriers only.
template <int H, class Load>
__device__ __forceinline__ void fft_forward_strict(cplx *buf, const FftTables &tb, int t, bool active, Load load) {
    {   // first pass reads its operands straight from the loader: thread t owns points t + m*H/8
        cplx x[8];
        if (active) {
#pragma unroll
            for (int m = 0; m < 8; m++) x[m] = load(t + m * (H / 8));
            fft_pass_regs<H, H / 8, 3, false>(x, tb.psi, t);
#pragma unroll
            for (int m = 0; m < 8; m++) buf[PAD(t + m * (H / 8))] = x[m];
        }
    }
    __syncthreads();
    if (H == 1024) {
        if (active) fft_pass_smem<H, 16, 3, false>(buf, tb.psi, t);
        __syncthreads();
        if (active) fft_pass_smem<H, 2, 3, false>(buf, tb.psi, t);
        __syncthreads();
        if (active) fft_pass_smem<H, 1, 1, false>(buf, tb.psi, t);
    } else {
        if (active) fft_pass_smem<H, 8, 3, false>(buf, tb.psi, t);
        __syncthreads();
        if (active) fft_pass_smem<H, 1, 3, false>(buf, tb.psi, t);
    }
    __syncthreads();
}

// Inverse transform of the spectrum in buf (slot order).  store(i, z) receives, for point i < H,
// z = ifft(...)[i] * rootsinv[i]  (fft.jl:76-77); the caller applies `native` to (re, -im).
template <int H, class Store>
__device__ __forceinline__ void fft_inverse_strict(cplx *buf, const FftTables &tb, int t, bool active, Store store) {
    if (H == 1024) {
        if (active) fft_pass_smem<H, 1, 1, true>(buf, tb.psiinv, t);
        __syncthreads();
        if (active) fft_pass_smem<H, 2, 3, true>(buf, tb.psiinv, t);
        __syncthreads();
        if (active) fft_pass_smem<H, 16, 3, true>(buf, tb.psiinv, t);
    } else {
        if (active) fft_pass_smem<H, 1, 3, true>(buf, tb.psiinv, t);
        __syncthreads();
        if (active) fft_pass_smem<H, 8, 3, true>(buf, tb.psiinv, t);
    }
    __syncthreads();
    if (active) {
        cplx x[8];
#pragma unroll
        for (int m = 0; m < 8; m++) x[m] = buf[PAD(t + m * (H / 8))];
        fft_pass_regs<H, H / 8, 3, true>(x, tb.psiinv, t);
#pragma unroll
        for (int m = 0; m < 8; m++) {
            const int i = t + m * (H / 8);
            store(i, cmul_s(x[m], __ldg(&tb.rootsinv[i])));
        }
    }
    __syncthreads();
}

// Loader helpers ------------------------------------------------------------------------------------
// Twisted input of gadget digit j of polynomial p (coefficient form, N = 2H torus values).
template <class T, int H> struct DigitLoad {
    const T *p; const cplx *roots; int j, l, logB;
    __device__ __forceinline__ cplx operator()(int i) const {
        const int32_t dr = gadget_digit<T>(p[i], j, l, logB);
        const int32_t di = gadget_digit<T>(p[i + H], j, l, logB);
        // signed(d_i) - im*signed(d_{i+H}); digits are in [-B/2, B/2) so the negation cannot wrap
        return cmul_s(make_double2((double)dr, (double)(-di)), __ldg(&roots[i]));
    }
};
// Twisted input of a full-width polynomial: fftto!(tacc, acc) at bootstrapping.jl:441 / fft.jl:60.
template <class T, int H> struct RawLoad {
    const T *p; const cplx *roots;
    __device__ __forceinline__ cplx operator()(int i) const {
        const double re = Torus<T>::to_double(p[i]);
        const double im = Torus<T>::to_double((T)((T)0 - p[i + H]));     // wraps in W bits like Complex{IntW}
        return cmul_s(make_double2(re, im), __ldg(&roots[i]));
    }
};
