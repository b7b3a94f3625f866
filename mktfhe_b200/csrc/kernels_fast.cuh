// kernels_fast.cuh -- FAST-mode KMS phase 1 (the >= 98 % of the flops): register-resident FP64 transform.
//
// What it computes: /root/reference/src/tfhe/bootstrapping.jl:389-443 (phase_1 of KMS) -- per RLEV row and LWE
// index: gadget-decompose (gsw.jl:86-96) -> forward transforms -> RGSW multiply-accumulate (polynomial.jl:105-108)
// -> x (X^a - 1) -> inverse transforms -> round (arithmetic.jl:6-9) -> accumulate.  Integer stages are the
// reference's exactly; floating-point stages use a different (mathematically equal) operation schedule:
//
//   * Transform: the reference twists by exp(i*pi*j/N) and then runs a negacyclic Cooley-Tukey network
//     (fft.jl:57-63,105-155).  Both steps together evaluate c(Y) = sum (p_j - i*p_{j+H}) Y^j at the roots of
//     Y^H + i, in bit-reversed order.  Here one radix-2 network does that directly: node (s, i) of the product
//     tree Y^(H/2^s) - rho(s,i), rho(0,0) = -i, splits with twiddle sqrt(rho(s,i)).  No twist pass, the
//     stage-1..4 twiddles are thread-invariant (constant bank), siblings differ by a factor -i, and slot n
//     holds the same evaluation point as the reference's slot n, so keys need no slot permutation.
//   * One unit = one (gate, party, row) = 64 threads x 16 points; passes of 4 + 4 + 2 stages in registers,
//     two shared-memory exchanges per transform with XOR swizzles (bank-conflict free for 16-byte accesses).
//   * Forward butterfly (t + w*u, t - w*u) in 6 FMA-pipe instructions: hi = 2t - lo.
//   * The 2 x 16 complex accumulators of the RGSW product stay in registers across the 2*l digit transforms;
//     bootstrapping-key polynomials are read straight from L2 with 16-byte coalesced loads in "thread order"
//     ([e][t] = slot 16t + e), each value used once per unit.
//   * (X^a - 1)/H is rebuilt per slot from two small tables: exp(-i*pi*(4*brv6(t)+1)*a/N) (per thread) times a
//     16th root of unity (per register slot); the 1/H of the inverse transform rides on it.
#pragma once
#include <cstdlib>
#include <string>
#include <vector>
#include "common.cuh"
#include "../../include/mktfhe_params.h"

namespace fast {

constexpr int H = 1024, N = 2048, UT = 64, U = 4, CTA = UT * U;

__constant__ double2 c_tw1[16];      // TW[1..15]: stages 1..4 (index 2^s + i), entry 0 unused
__constant__ double2 c_e16[16];      // exp(-i*pi*j/8)

struct Tables {
    const cplx *t2;        // [8][16]  pass-2 twiddles per 64-point block: w4, w5, w6a, w6b, w7a..w7d (even nodes)
    const cplx *t8;        // [2][64]  TW[256 + 4t + g], g = 0, 2
    const cplx *t9;        // [4][64]  TW[512 + 2(4t + g)], g = 0..3
    const cplx *emono;     // [4096] exp(-i*pi*m/2048) / H
    const cplx *t2w;       // [16][32] pass-2 twiddles of the one-warp transform (kernels_fast_w.cuh)
};

// ---- butterflies -------------------------------------------------------------------------------------
// forward: (a, b) -> (a + w*b, a - w*b); 6 instructions (hi = 2a - lo)
__device__ __forceinline__ void bf(cplx &a, cplx &b, const cplx w) {
    const double lx = fma(-w.y, b.y, fma(w.x, b.x, a.x));
    const double ly = fma(w.y, b.x, fma(w.x, b.y, a.y));
    b.x = fma(2.0, a.x, -lx); b.y = fma(2.0, a.y, -ly);
    a.x = lx; a.y = ly;
}
// forward with twiddle -i*w (sibling node): w*b rotated by -i
__device__ __forceinline__ void bf_mi(cplx &a, cplx &b, const cplx w) {
    const double lx = fma(w.y, b.x, fma(w.x, b.y, a.x));        // a.x + Im(w*b)
    const double ly = fma(w.y, b.y, fma(-w.x, b.x, a.y));       // a.y - Re(w*b)
    b.x = fma(2.0, a.x, -lx); b.y = fma(2.0, a.y, -ly);
    a.x = lx; a.y = ly;
}
// inverse (unscaled): (A, B) -> (A + B, (A - B) * conj(w))
__device__ __forceinline__ void bi(cplx &a, cplx &b, const cplx w) {
    const double dx = a.x - b.x, dy = a.y - b.y;
    a.x += b.x; a.y += b.y;
    b.x = fma(w.y, dy, w.x * dx);            // Re(d * conj(w)) = dx*wx + dy*wy
    b.y = fma(-w.y, dx, w.x * dy);           // Im(d * conj(w)) = dy*wx - dx*wy
}
// inverse with twiddle -i*w: conj(-i*w) = i*conj(w)
__device__ __forceinline__ void bi_mi(cplx &a, cplx &b, const cplx w) {
    const double dx = a.x - b.x, dy = a.y - b.y;
    a.x += b.x; a.y += b.y;
    b.x = -fma(-w.y, dx, w.x * dy);          // Re(i*z) = -Im(z)
    b.y = fma(w.y, dy, w.x * dx);            // Im(i*z) =  Re(z)
}

// ---- in-register passes over x[16] ---------------------------------------------------------------------
// stages 1..4 (spans 512..64): element m of thread t is point t + 64m; twiddles TW[2^s + (m >> (4-s))]
__device__ __forceinline__ void pass1_fwd(cplx (&x)[16]) {
#pragma unroll
    for (int m = 0; m < 8; m++) bf(x[m], x[m + 8], c_tw1[1]);
#pragma unroll
    for (int m = 0; m < 16; m++) if (!(m & 4)) { if (m & 8) bf_mi(x[m], x[m + 4], c_tw1[2]); else bf(x[m], x[m + 4], c_tw1[2]); }
#pragma unroll
    for (int m = 0; m < 16; m++) if (!(m & 2)) { if (m & 4) bf_mi(x[m], x[m + 2], c_tw1[4 + (m >> 3) * 2]); else bf(x[m], x[m + 2], c_tw1[4 + (m >> 3) * 2]); }
#pragma unroll
    for (int m = 0; m < 16; m += 2) { if (m & 2) bf_mi(x[m], x[m + 1], c_tw1[8 + (m >> 2) * 2]); else bf(x[m], x[m + 1], c_tw1[8 + (m >> 2) * 2]); }
}
__device__ __forceinline__ void pass1_inv(cplx (&x)[16]) {
#pragma unroll
    for (int m = 0; m < 16; m += 2) { if (m & 2) bi_mi(x[m], x[m + 1], c_tw1[8 + (m >> 2) * 2]); else bi(x[m], x[m + 1], c_tw1[8 + (m >> 2) * 2]); }
#pragma unroll
    for (int m = 0; m < 16; m++) if (!(m & 2)) { if (m & 4) bi_mi(x[m], x[m + 2], c_tw1[4 + (m >> 3) * 2]); else bi(x[m], x[m + 2], c_tw1[4 + (m >> 3) * 2]); }
#pragma unroll
    for (int m = 0; m < 16; m++) if (!(m & 4)) { if (m & 8) bi_mi(x[m], x[m + 4], c_tw1[2]); else bi(x[m], x[m + 4], c_tw1[2]); }
#pragma unroll
    for (int m = 0; m < 8; m++) bi(x[m], x[m + 8], c_tw1[1]);
}
// stages 5..8 inside 64-point block `blk`: element q of the thread is point 64*blk + o + 4q;
// twiddles TW[2^(4+ss) + (blk << ss) + (q >> (4-ss))]; odd nodes come from their even sibling.
__device__ __forceinline__ void pass2_fwd(cplx (&x)[16], const cplx *__restrict__ tw, int blk) {
    const cplx w4 = tw[blk], w5 = tw[16 + blk];
    const cplx w6a = tw[32 + blk], w6b = tw[48 + blk];
    const cplx w7a = tw[64 + blk], w7b = tw[80 + blk], w7c = tw[96 + blk], w7d = tw[112 + blk];
#pragma unroll
    for (int q = 0; q < 8; q++) bf(x[q], x[q + 8], w4);
#pragma unroll
    for (int q = 0; q < 16; q++) if (!(q & 4)) { if (q & 8) bf_mi(x[q], x[q + 4], w5); else bf(x[q], x[q + 4], w5); }
#pragma unroll
    for (int q = 0; q < 16; q++) if (!(q & 2)) {
        const cplx w = (q & 8) ? w6b : w6a;
        if (q & 4) bf_mi(x[q], x[q + 2], w); else bf(x[q], x[q + 2], w);
    }
#pragma unroll
    for (int q = 0; q < 16; q += 2) {
        const cplx w = (q >> 2) == 0 ? w7a : (q >> 2) == 1 ? w7b : (q >> 2) == 2 ? w7c : w7d;
        if (q & 2) bf_mi(x[q], x[q + 1], w); else bf(x[q], x[q + 1], w);
    }
}
__device__ __forceinline__ void pass2_inv(cplx (&x)[16], const cplx *__restrict__ tw, int blk) {
    const cplx w4 = tw[blk], w5 = tw[16 + blk];
    const cplx w6a = tw[32 + blk], w6b = tw[48 + blk];
    const cplx w7a = tw[64 + blk], w7b = tw[80 + blk], w7c = tw[96 + blk], w7d = tw[112 + blk];
#pragma unroll
    for (int q = 0; q < 16; q += 2) {
        const cplx w = (q >> 2) == 0 ? w7a : (q >> 2) == 1 ? w7b : (q >> 2) == 2 ? w7c : w7d;
        if (q & 2) bi_mi(x[q], x[q + 1], w); else bi(x[q], x[q + 1], w);
    }
#pragma unroll
    for (int q = 0; q < 16; q++) if (!(q & 2)) {
        const cplx w = (q & 8) ? w6b : w6a;
        if (q & 4) bi_mi(x[q], x[q + 2], w); else bi(x[q], x[q + 2], w);
    }
#pragma unroll
    for (int q = 0; q < 16; q++) if (!(q & 4)) { if (q & 8) bi_mi(x[q], x[q + 4], w5); else bi(x[q], x[q + 4], w5); }
#pragma unroll
    for (int q = 0; q < 8; q++) bi(x[q], x[q + 8], w4);
}
// stages 9, 10 on the four 4-point groups 16t + 4g .. +3 of thread t: nodes 4t+g (depth 8) and 2(4t+g)+{0,1}
__device__ __forceinline__ void pass3_fwd(cplx (&x)[16], const cplx *__restrict__ tw8, const cplx *__restrict__ tw9e, int t) {
#pragma unroll
    for (int g = 0; g < 4; g++) {
        const cplx w8 = tw8[(g >> 1) * UT + t], w9 = tw9e[g * UT + t];
        if (g & 1) { bf_mi(x[4 * g], x[4 * g + 2], w8); bf_mi(x[4 * g + 1], x[4 * g + 3], w8); }
        else { bf(x[4 * g], x[4 * g + 2], w8); bf(x[4 * g + 1], x[4 * g + 3], w8); }
        bf(x[4 * g], x[4 * g + 1], w9);
        bf_mi(x[4 * g + 2], x[4 * g + 3], w9);
    }
}
__device__ __forceinline__ void pass3_inv(cplx (&x)[16], const cplx *__restrict__ tw8, const cplx *__restrict__ tw9e, int t) {
#pragma unroll
    for (int g = 0; g < 4; g++) {
        const cplx w8 = tw8[(g >> 1) * UT + t], w9 = tw9e[g * UT + t];
        bi(x[4 * g], x[4 * g + 1], w9);
        bi_mi(x[4 * g + 2], x[4 * g + 3], w9);
        if (g & 1) { bi_mi(x[4 * g], x[4 * g + 2], w8); bi_mi(x[4 * g + 1], x[4 * g + 3], w8); }
        else { bi(x[4 * g], x[4 * g + 2], w8); bi(x[4 * g + 1], x[4 * g + 3], w8); }
    }
}

__device__ __forceinline__ void unit_bar(int unit) { asm volatile("bar.sync %0, %1;" ::"r"(unit + 1), "r"(UT) : "memory"); }

// Exchange layouts (16-byte elements; a quarter-warp must hit 8 distinct 16-byte bank groups).  Additive padding
// keeps every access at [per-thread base + compile-time offset]:
//   exchange 1: pos1(n) = n + 4*(n >> 6)   (64-point blocks skewed by 4)
//   exchange 2: pos2(n) = n + (n >> 4)     (16-slot runs skewed by 1)
constexpr int XB_LEN = H + 64;
__device__ __forceinline__ constexpr int p1(int n) { return n + 4 * (n >> 6); }
__device__ __forceinline__ constexpr int p2(int n) { return n + (n >> 4); }

// x (layout A: point t + 64m) -> pass 1..3 -> x (layout C: slot 16t + e)
__device__ __forceinline__ void fft_fwd(cplx (&x)[16], cplx *xb, const cplx *tw2, const cplx *tw8, const cplx *tw9e, int t, int unit) {
    pass1_fwd(x);
    unit_bar(unit);                                              // previous readers of xb are done
#pragma unroll
    for (int m = 0; m < 16; m++) xb[t + 68 * m] = x[m];                       // pos1(t + 64m)
    unit_bar(unit);
    const int blk = t >> 2, o = t & 3;
    cplx *xq = xb + 68 * blk + o;
#pragma unroll
    for (int q = 0; q < 16; q++) x[q] = xq[4 * q];                            // pos1(64blk + o + 4q)
    pass2_fwd(x, tw2, blk);
    unit_bar(unit);
#pragma unroll
    for (int q = 0; q < 16; q++) xq[4 * q + (q >> 2)] = x[q];                 // pos2(64blk + o + 4q)
    unit_bar(unit);
#pragma unroll
    for (int e = 0; e < 16; e++) x[e] = xb[17 * t + e];                       // pos2(16t + e)
    pass3_fwd(x, tw8, tw9e, t);
}
// x (layout C) -> x (layout A), unscaled inverse
__device__ __forceinline__ void fft_inv(cplx (&x)[16], cplx *xb, const cplx *tw2, const cplx *tw8, const cplx *tw9e, int t, int unit) {
    pass3_inv(x, tw8, tw9e, t);
    unit_bar(unit);
#pragma unroll
    for (int e = 0; e < 16; e++) xb[17 * t + e] = x[e];
    unit_bar(unit);
    const int blk = t >> 2, o = t & 3;
    cplx *xq = xb + 68 * blk + o;
#pragma unroll
    for (int q = 0; q < 16; q++) x[q] = xq[4 * q + (q >> 2)];
    pass2_inv(x, tw2, blk);
    unit_bar(unit);
#pragma unroll
    for (int q = 0; q < 16; q++) xq[4 * q] = x[q];
    unit_bar(unit);
#pragma unroll
    for (int m = 0; m < 16; m++) x[m] = xb[t + 68 * m];
    pass1_inv(x);
}

// int32 -> double without the conversion pipe
__device__ __forceinline__ double i2d(int32_t d) {
    return __hiloint2double(0x43300000, (int)((uint32_t)d ^ 0x80000000u)) - 4503601774854144.0;   // 2^52 + 2^31
}
// x mod 2^64 as uint64, rounding toward -inf like `native` (arithmetic.jl:6-9)
__device__ __forceinline__ uint64_t d2torus(double x) {
    const double q = fma(x, 5.421010862427522e-20, 6755399441055744.0) - 6755399441055744.0;   // rint(x / 2^64)
    const double y = fma(-18446744073709551616.0, q, x);
    return (uint64_t)__double2ll_rd(y);
}

struct Args {
    const uint32_t *tilde;        // [B][lwe_words] (phase 1) / [units] rotations (step mode)
    const cplx *const *brk;       // [k] FAST-layout keys: [idx][dg][comp][e][t]
    Tables tb;
    cplx *lev_out;                // [B][R][2][H], reference slot order (lev_fast: thread order [e][t], scaled by 1/H)
    int lev_fast;
    uint64_t *acc_io;             // step mode: [units][2][N]
    int step_mode, step_party, step_idx;
    int n, d, k, l, logB, l_lev, logB_lev, R, lwe_words;
    size_t units;
};

constexpr size_t SMEM_UNIT = (size_t)2 * N * 8 + (size_t)XB_LEN * 16;         // acc b, a + exchange buffer
constexpr size_t SMEM_BYTES = U * SMEM_UNIT + (size_t)(128 + 128 + 256) * 16;  // + t2, t8, t9

template <int ELL>
__global__ void __launch_bounds__(CTA, 1) k_phase1(const Args a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int tid = threadIdx.x, unit_l = tid / UT, t = tid % UT;
    cplx *tw2 = reinterpret_cast<cplx *>(smem_raw + U * SMEM_UNIT), *tw8 = tw2 + 128, *tw9e = tw8 + 128;
    for (int i = tid; i < 256; i += CTA) { if (i < 128) { tw2[i] = a.tb.t2[i]; tw8[i] = a.tb.t8[i]; } tw9e[i] = a.tb.t9[i]; }
    uint64_t *accb = reinterpret_cast<uint64_t *>(smem_raw + unit_l * SMEM_UNIT), *acca = accb + N;
    cplx *xb = reinterpret_cast<cplx *>(accb + 2 * N);
    __syncthreads();

    const size_t unit = (size_t)blockIdx.x * U + unit_l;
    const bool live = unit < a.units;            // dead units still take part in nothing: they own their barrier id
    if (!live) return;

    int gate, party, row = 0;
    if (!a.step_mode) {
        gate = (int)(unit / a.R);
        const int r = (int)(unit % a.R);
        party = r == 0 ? 0 : 1 + (r - 1) / a.l_lev;
        row = r == 0 ? 0 : (r - 1) % a.l_lev;
#pragma unroll
        for (int m = 0; m < 32; m++) { accb[t + 64 * m] = 0; acca[t + 64 * m] = 0; }
        if (t == 0) accb[0] = (uint64_t)1 << (64 - (row + 1) * a.logB_lev);          // bootstrapping.jl:402-408
    } else {
        gate = (int)unit; party = a.step_party;
        const uint64_t *src = a.acc_io + unit * 2 * N;
#pragma unroll
        for (int m = 0; m < 32; m++) { accb[t + 64 * m] = src[t + 64 * m]; acca[t + 64 * m] = src[N + t + 64 * m]; }
    }
    // thread-private ownership: thread t only ever touches coefficients t + 64m (+H) of its unit, so no barrier is
    // needed around the accumulator itself.

    const int l = a.l, logB = a.logB;
    const int bit = 64 - l * logB;
    // divbits rounding (arithmetic.jl:23-27) and the balanced-digit carry chain (gsw.jl:86-96) in one 64-bit add:
    // adding 2^(bit-1) rounds, adding B/2 at every digit position turns the carry chain into plain field extraction
    // (field - B/2 is the balanced digit); everything is mod 2^64, i.e. mod 2^(l*logB) on the digits, like the reference.
    uint64_t cadd = (uint64_t)1 << (bit - 1);
    for (int j = 0; j < l; j++) cadd += (uint64_t)1 << (bit + j * logB + logB - 1);
    const uint32_t mask = (1u << logB) - 1;
    const double dbias = 4503599627370496.0 + (double)(1 << (logB - 1));       // 2^52 + B/2
    const cplx *brk = a.brk[party];
    const size_t per_idx = (size_t)4 * l * H;
    const uint32_t *at_src = a.step_mode ? a.tilde + unit * ELL : a.tilde + (size_t)gate * a.lwe_words + 1 + (size_t)party * a.n;
    const int nsteps = a.step_mode ? 1 : (ELL == 1 ? a.n : a.d);
    int brv6t = (int)(__brev((unsigned)t) >> 26);

    for (int step = 0; step < nsteps; step++) {
        // ELL == 1: one LWE index per step (:412-438).  ELL > 1: one block of ELL key bits per step (:624-655);
        // the block's monomials are folded into the keys so one accumulator pair serves the whole block:
        //   sum_bit mono_bit * (sum_dg D_dg * K_bit,dg)  =  sum_dg D_dg * (sum_bit mono_bit * K_bit,dg)
        uint32_t atv[ELL];
        bool any = false;
#pragma unroll
        for (int b = 0; b < ELL; b++) { atv[b] = at_src[(a.step_mode ? 0 : step * ELL) + b]; any |= atv[b] > 0; }
        if (!any) continue;                                                           // :413 / whole-block no-op
        const uint32_t at = atv[0];
        const int idx = (a.step_mode ? a.step_idx : step) * ELL;
        const cplx *kidx = brk + (size_t)idx * per_idx + t;
        cplx m1v[ELL];
#pragma unroll
        for (int b = 0; b < ELL; b++) m1v[b] = __ldg(&a.tb.emono[((4 * brv6t + 1) * atv[b]) & 4095]);

        cplx tb_[16], ta_[16];
#pragma unroll
        for (int e = 0; e < 16; e++) tb_[e] = ta_[e] = make_double2(0.0, 0.0);

        for (int dg = 0; dg < 2 * l; dg++) {
            const uint64_t *src = dg < l ? accb : acca;
            const int sh = bit + (l - 1 - (dg < l ? dg : dg - l)) * logB;
            cplx x[16];
#pragma unroll
            for (int m = 0; m < 16; m++) {
                const uint64_t v0 = src[t + 64 * m] + cadd, v1 = src[t + 64 * m + H] + cadd;
                const uint32_t f0 = (uint32_t)(v0 >> sh) & mask, f1 = (uint32_t)(v1 >> sh) & mask;
                // signed(d_j) - im*signed(d_{j+H}); 2^52 + field is exact in the double's mantissa
                x[m] = make_double2(__hiloint2double(0x43300000, (int)f0) - dbias, dbias - __hiloint2double(0x43300000, (int)f1));
            }
            fft_fwd(x, xb, tw2, tw8, tw9e, t, unit_l);
            const cplx *kb = kidx + (size_t)(dg * 2) * H, *ka = kb + H;
            if (ELL == 1) {
#pragma unroll
                for (int e = 0; e < 16; e++) {
                    const cplx wb = __ldg(kb + e * UT), wa = __ldg(ka + e * UT);
                    tb_[e] = cmac_f(tb_[e], x[e], wb);
                    ta_[e] = cmac_f(ta_[e], x[e], wa);
                }
            } else {
#pragma unroll
                for (int e = 0; e < 16; e++) {
                    const int b4 = ((e & 1) << 3) | ((e & 2) << 1) | ((e & 4) >> 1) | ((e & 8) >> 3);
                    cplx kcb = make_double2(0.0, 0.0), kca = kcb;
#pragma unroll
                    for (int b = 0; b < ELL; b++) {
                        if (atv[b] == 0) continue;
                        cplx mo = cmul_f(m1v[b], c_e16[(atv[b] * b4) & 15]);
                        mo.x -= 1.0 / H;
                        kcb = cmac_f(kcb, mo, __ldg(kb + b * per_idx + e * UT));
                        kca = cmac_f(kca, mo, __ldg(ka + b * per_idx + e * UT));
                    }
                    tb_[e] = cmac_f(tb_[e], x[e], kcb);
                    ta_[e] = cmac_f(ta_[e], x[e], kca);
                }
            }
        }
        // (X^a - 1) / H per slot: slot n = 16t + e evaluates at exp(-i*pi*(4*brv10(n)+1)/N), brv10(n) = 64*brv4(e) + brv6(t)
        if (ELL == 1) {
            const cplx m1 = m1v[0];
#pragma unroll
            for (int e = 0; e < 16; e++) {
                const int b4 = ((e & 1) << 3) | ((e & 2) << 1) | ((e & 4) >> 1) | ((e & 8) >> 3);
                cplx mo = cmul_f(m1, c_e16[(at * b4) & 15]);
                mo.x -= 1.0 / H;
                tb_[e] = cmul_f(mo, tb_[e]);
                ta_[e] = cmul_f(mo, ta_[e]);
            }
        }
        fft_inv(tb_, xb, tw2, tw8, tw9e, t, unit_l);
#pragma unroll
        for (int m = 0; m < 16; m++) {
            accb[t + 64 * m] += d2torus(tb_[m].x);
            accb[t + 64 * m + H] += d2torus(-tb_[m].y);
        }
        fft_inv(ta_, xb, tw2, tw8, tw9e, t, unit_l);
#pragma unroll
        for (int m = 0; m < 16; m++) {
            acca[t + 64 * m] += d2torus(ta_[m].x);
            acca[t + 64 * m + H] += d2torus(-ta_[m].y);
        }
    }

    if (!a.step_mode) {            // fftto!(tacc, acc): bootstrapping.jl:441 -- full-width coefficients
        cplx *out = a.lev_out + unit * 2 * H;
#pragma unroll 1
        for (int c = 0; c < 2; c++) {
            const uint64_t *src = c == 0 ? accb : acca;
            cplx x[16];
#pragma unroll
            for (int m = 0; m < 16; m++)
                x[m] = make_double2(__ll2double_rn((long long)src[t + 64 * m]),
                                    __ll2double_rn((long long)((uint64_t)0 - src[t + 64 * m + H])));
            fft_fwd(x, xb, tw2, tw8, tw9e, t, unit_l);
#pragma unroll
            for (int e = 0; e < 16; e++) out[(size_t)c * H + 16 * t + e] = x[e];
        }
    } else {
        uint64_t *dst = a.acc_io + unit * 2 * N;
#pragma unroll
        for (int m = 0; m < 32; m++) { dst[t + 64 * m] = accb[t + 64 * m]; dst[N + t + 64 * m] = acca[t + 64 * m]; }
    }
}

// ======================================================================================================
// TMEM variant (MKTFHE_FAST_KERNEL=tmem).  Blackwell's tensor memory is lane-private per warp quadrant, which is exactly the
// ownership pattern of this kernel: thread t only ever touches RLWE-accumulator coefficients t + 64m (+H) and
// RGSW-accumulator slots 16t + e.  Both accumulators therefore live in TMEM (tcgen05.ld / tcgen05.st,
// measured ~850 B/clk/SM against 128 B/clk/SM for shared memory):
//   per thread 256 columns: [0,64) acc.b  [64,128) acc.a  [128,192) tacc.b  [192,256) tacc.a
//   warp w -> lanes 32*(w%4).., columns 256*(w/4)..   (8 warps fill the 512-column allocation of the CTA)
// That frees 128 registers per thread (bootstrapping-key values are prefetched before the pass that precedes
// their use) and 128 KiB of shared memory (the exchange buffer is double-buffered: 2 barriers per transform).
__device__ __forceinline__ void tm_st16(uint32_t addr, const uint32_t (&v)[16]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
                 :: "r"(addr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]),
                    "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]) : "memory");
}
__device__ __forceinline__ void tm_ld16(uint32_t addr, uint32_t (&v)[16]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                   "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                 : "r"(addr) : "memory");
}
__device__ __forceinline__ void tm_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// tcgen05.ld is asynchronous: its destination registers are only valid after tcgen05.wait::ld.  The compiler sees no
// data dependency between the wait and those registers, so every consumer first passes them through this empty
// volatile asm (volatile asms keep their order), which pins all later uses behind the wait.
__device__ __forceinline__ void tm_pin16(uint32_t (&v)[16]) {
    asm volatile("" : "+r"(v[0]), "+r"(v[1]), "+r"(v[2]), "+r"(v[3]), "+r"(v[4]), "+r"(v[5]), "+r"(v[6]), "+r"(v[7]),
                      "+r"(v[8]), "+r"(v[9]), "+r"(v[10]), "+r"(v[11]), "+r"(v[12]), "+r"(v[13]), "+r"(v[14]), "+r"(v[15]) :: "memory");
}
__device__ __forceinline__ void tm_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
// 4 complex doubles <-> 16 TMEM columns
__device__ __forceinline__ void tm_ld_c4(uint32_t addr, cplx (&z)[4]) {
    uint32_t v[16];
    tm_ld16(addr, v);
    tm_wait_ld();
    tm_pin16(v);
#pragma unroll
    for (int i = 0; i < 4; i++) z[i] = make_double2(__hiloint2double((int)v[4 * i + 1], (int)v[4 * i]), __hiloint2double((int)v[4 * i + 3], (int)v[4 * i + 2]));
}
__device__ __forceinline__ void tm_st_c4(uint32_t addr, const cplx (&z)[4]) {
    uint32_t v[16];
#pragma unroll
    for (int i = 0; i < 4; i++) {
        v[4 * i] = (uint32_t)__double2loint(z[i].x); v[4 * i + 1] = (uint32_t)__double2hiint(z[i].x);
        v[4 * i + 2] = (uint32_t)__double2loint(z[i].y); v[4 * i + 3] = (uint32_t)__double2hiint(z[i].y);
    }
    tm_st16(addr, v);
}

constexpr int TM_ACC_B = 0, TM_ACC_A = 64, TM_TACC_B = 128, TM_TACC_A = 192;
constexpr size_t SMEM_UNIT_TM = (size_t)2 * XB_LEN * 16;                                   // two exchange buffers
constexpr size_t SMEM_BYTES_TM = U * SMEM_UNIT_TM + (size_t)(128 + 128 + 256) * 16 + 16;

// Per-thread twiddles of passes 2 and 3 (they depend on t only), loaded once per kernel: 14 shared-memory loads per
// transform less, for 56 registers that the TMA kernel has to spare.
struct TwRegs {
    cplx w4, w5, w6a, w6b, w7a, w7b, w7c, w7d, w8[2], w9[4];
    __device__ __forceinline__ void load(const cplx *tw2, const cplx *tw8, const cplx *tw9e, int t) {
        const int blk = t >> 2;
        w4 = tw2[blk]; w5 = tw2[16 + blk]; w6a = tw2[32 + blk]; w6b = tw2[48 + blk];
        w7a = tw2[64 + blk]; w7b = tw2[80 + blk]; w7c = tw2[96 + blk]; w7d = tw2[112 + blk];
        w8[0] = tw8[t]; w8[1] = tw8[UT + t];
#pragma unroll
        for (int g = 0; g < 4; g++) w9[g] = tw9e[g * UT + t];
    }
};
__device__ __forceinline__ void pass2_fwd_r(cplx (&x)[16], const TwRegs &w) {
#pragma unroll
    for (int q = 0; q < 8; q++) bf(x[q], x[q + 8], w.w4);
#pragma unroll
    for (int q = 0; q < 16; q++) if (!(q & 4)) { if (q & 8) bf_mi(x[q], x[q + 4], w.w5); else bf(x[q], x[q + 4], w.w5); }
#pragma unroll
    for (int q = 0; q < 16; q++) if (!(q & 2)) {
        const cplx ww = (q & 8) ? w.w6b : w.w6a;
        if (q & 4) bf_mi(x[q], x[q + 2], ww); else bf(x[q], x[q + 2], ww);
    }
#pragma unroll
    for (int q = 0; q < 16; q += 2) {
        const cplx ww = (q >> 2) == 0 ? w.w7a : (q >> 2) == 1 ? w.w7b : (q >> 2) == 2 ? w.w7c : w.w7d;
        if (q & 2) bf_mi(x[q], x[q + 1], ww); else bf(x[q], x[q + 1], ww);
    }
}
__device__ __forceinline__ void pass2_inv_r(cplx (&x)[16], const TwRegs &w) {
#pragma unroll
    for (int q = 0; q < 16; q += 2) {
        const cplx ww = (q >> 2) == 0 ? w.w7a : (q >> 2) == 1 ? w.w7b : (q >> 2) == 2 ? w.w7c : w.w7d;
        if (q & 2) bi_mi(x[q], x[q + 1], ww); else bi(x[q], x[q + 1], ww);
    }
#pragma unroll
    for (int q = 0; q < 16; q++) if (!(q & 2)) {
        const cplx ww = (q & 8) ? w.w6b : w.w6a;
        if (q & 4) bi_mi(x[q], x[q + 2], ww); else bi(x[q], x[q + 2], ww);
    }
#pragma unroll
    for (int q = 0; q < 16; q++) if (!(q & 4)) { if (q & 8) bi_mi(x[q], x[q + 4], w.w5); else bi(x[q], x[q + 4], w.w5); }
#pragma unroll
    for (int q = 0; q < 8; q++) bi(x[q], x[q + 8], w.w4);
}
__device__ __forceinline__ void pass3_fwd_r(cplx (&x)[16], const TwRegs &w) {
#pragma unroll
    for (int g = 0; g < 4; g++) {
        const cplx w8 = w.w8[g >> 1], w9 = w.w9[g];
        if (g & 1) { bf_mi(x[4 * g], x[4 * g + 2], w8); bf_mi(x[4 * g + 1], x[4 * g + 3], w8); }
        else { bf(x[4 * g], x[4 * g + 2], w8); bf(x[4 * g + 1], x[4 * g + 3], w8); }
        bf(x[4 * g], x[4 * g + 1], w9);
        bf_mi(x[4 * g + 2], x[4 * g + 3], w9);
    }
}
__device__ __forceinline__ void pass3_inv_r(cplx (&x)[16], const TwRegs &w) {
#pragma unroll
    for (int g = 0; g < 4; g++) {
        const cplx w8 = w.w8[g >> 1], w9 = w.w9[g];
        bi(x[4 * g], x[4 * g + 1], w9);
        bi_mi(x[4 * g + 2], x[4 * g + 3], w9);
        if (g & 1) { bi_mi(x[4 * g], x[4 * g + 2], w8); bi_mi(x[4 * g + 1], x[4 * g + 3], w8); }
        else { bi(x[4 * g], x[4 * g + 2], w8); bi(x[4 * g + 1], x[4 * g + 3], w8); }
    }
}
// fft_fwd2 / fft_inv2 with register-resident twiddles
__device__ __forceinline__ void fft_fwd2r(cplx (&x)[16], cplx *xa, cplx *xc, const TwRegs &w, int t, int unit) {
    pass1_fwd(x);
#pragma unroll
    for (int m = 0; m < 16; m++) xa[t + 68 * m] = x[m];
    unit_bar(unit);
    const int blk = t >> 2, o = t & 3;
#pragma unroll
    for (int q = 0; q < 16; q++) x[q] = xa[68 * blk + o + 4 * q];
    pass2_fwd_r(x, w);
#pragma unroll
    for (int q = 0; q < 16; q++) xc[68 * blk + o + 4 * q + (q >> 2)] = x[q];
    unit_bar(unit);
#pragma unroll
    for (int e = 0; e < 16; e++) x[e] = xc[17 * t + e];
    pass3_fwd_r(x, w);
}
__device__ __forceinline__ void fft_inv2r(cplx (&x)[16], cplx *xa, cplx *xc, const TwRegs &w, int t, int unit) {
    pass3_inv_r(x, w);
#pragma unroll
    for (int e = 0; e < 16; e++) xa[17 * t + e] = x[e];
    unit_bar(unit);
    const int blk = t >> 2, o = t & 3;
#pragma unroll
    for (int q = 0; q < 16; q++) x[q] = xa[68 * blk + o + 4 * q + (q >> 2)];
    pass2_inv_r(x, w);
#pragma unroll
    for (int q = 0; q < 16; q++) xc[68 * blk + o + 4 * q] = x[q];
    unit_bar(unit);
#pragma unroll
    for (int m = 0; m < 16; m++) x[m] = xc[t + 68 * m];
    pass1_inv(x);
}

// forward transform with double-buffered exchanges; `mid1` / `mid2` run after the exchange reads, i.e. while the
// pass that follows is still ahead: the caller issues its key prefetches there.
template <class F1, class F2>
__device__ __forceinline__ void fft_fwd2(cplx (&x)[16], cplx *xa, cplx *xc, const cplx *tw2, const cplx *tw8, const cplx *tw9e,
                                         int t, int unit, F1 mid1, F2 mid2) {
    pass1_fwd(x);
#pragma unroll
    for (int m = 0; m < 16; m++) xa[t + 68 * m] = x[m];
    unit_bar(unit);
    const int blk = t >> 2, o = t & 3;
#pragma unroll
    for (int q = 0; q < 16; q++) x[q] = xa[68 * blk + o + 4 * q];
    mid1();
    pass2_fwd(x, tw2, blk);
#pragma unroll
    for (int q = 0; q < 16; q++) xc[68 * blk + o + 4 * q + (q >> 2)] = x[q];
    unit_bar(unit);
#pragma unroll
    for (int e = 0; e < 16; e++) x[e] = xc[17 * t + e];
    mid2();
    pass3_fwd(x, tw8, tw9e, t);
}
// The two buffers strictly alternate over the whole kernel (forward: xa then xc; inverse: xa then xc as well), so a
// buffer is only rewritten after a barrier that every reader of its previous contents has passed.
__device__ __forceinline__ void fft_inv2(cplx (&x)[16], cplx *xa, cplx *xc, const cplx *tw2, const cplx *tw8, const cplx *tw9e, int t, int unit) {
    pass3_inv(x, tw8, tw9e, t);
#pragma unroll
    for (int e = 0; e < 16; e++) xa[17 * t + e] = x[e];
    unit_bar(unit);
    const int blk = t >> 2, o = t & 3;
#pragma unroll
    for (int q = 0; q < 16; q++) x[q] = xa[68 * blk + o + 4 * q + (q >> 2)];
    pass2_inv(x, tw2, blk);
#pragma unroll
    for (int q = 0; q < 16; q++) xc[68 * blk + o + 4 * q] = x[q];
    unit_bar(unit);
#pragma unroll
    for (int m = 0; m < 16; m++) x[m] = xc[t + 68 * m];
    pass1_inv(x);
}

template <int ELL>
__global__ void __launch_bounds__(CTA, 1) k_phase1_tm(const Args a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int tid = threadIdx.x, warp = tid >> 5, unit_l = tid / UT, t = tid % UT;
    cplx *tw2 = reinterpret_cast<cplx *>(smem_raw + U * SMEM_UNIT_TM), *tw8 = tw2 + 128, *tw9e = tw8 + 128;
    uint32_t *tm_base_s = reinterpret_cast<uint32_t *>(tw9e + 256);
    for (int i = tid; i < 256; i += CTA) { if (i < 128) { tw2[i] = a.tb.t2[i]; tw8[i] = a.tb.t8[i]; } tw9e[i] = a.tb.t9[i]; }
    cplx *xa = reinterpret_cast<cplx *>(smem_raw + unit_l * SMEM_UNIT_TM), *xc = xa + XB_LEN;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" :: "r"((uint32_t)__cvta_generic_to_shared(tm_base_s)));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t tm = *tm_base_s + ((uint32_t)(32 * (warp & 3)) << 16) + 256u * (uint32_t)(warp >> 2);

    const size_t unit = (size_t)blockIdx.x * U + unit_l;
    if (unit < a.units) {
        int gate, party, row = 0;
        if (!a.step_mode) {
            gate = (int)(unit / a.R);
            const int r = (int)(unit % a.R);
            party = r == 0 ? 0 : 1 + (r - 1) / a.l_lev;
            row = r == 0 ? 0 : (r - 1) % a.l_lev;
            uint32_t z[16];
#pragma unroll
            for (int i = 0; i < 16; i++) z[i] = 0u;
            const uint64_t gv = t == 0 ? (uint64_t)1 << (64 - (row + 1) * a.logB_lev) : 0;   // bootstrapping.jl:402-408
#pragma unroll
            for (int c = 0; c < 128; c += 16) {                // tcgen05.st is warp-collective: same instruction on every lane
                z[0] = c == 0 ? (uint32_t)gv : 0u;
                z[1] = c == 0 ? (uint32_t)(gv >> 32) : 0u;
                tm_st16(tm + c, z);
            }
        } else {
            gate = (int)unit; party = a.step_party;
            const uint64_t *src = a.acc_io + unit * 2 * N;
#pragma unroll
            for (int pz = 0; pz < 2; pz++)
#pragma unroll
                for (int c = 0; c < 4; c++) {                                             // 8 coefficients per 16 columns
                    uint32_t v[16];
#pragma unroll
                    for (int i = 0; i < 8; i++) {
                        const int ci = 8 * c + i;                                          // 0..31: m = ci & 15, upper half = ci >> 4
                        const uint64_t w = src[pz * N + t + 64 * (ci & 15) + (ci >> 4) * H];
                        v[2 * i] = (uint32_t)w; v[2 * i + 1] = (uint32_t)(w >> 32);
                    }
                    tm_st16(tm + pz * 64 + 16 * c, v);
                }
        }
        tm_wait_st();

        const int l = a.l, logB = a.logB;
        const int bit = 64 - l * logB;
        uint64_t cadd = (uint64_t)1 << (bit - 1);
        for (int j = 0; j < l; j++) cadd += (uint64_t)1 << (bit + j * logB + logB - 1);
        const uint32_t mask = (1u << logB) - 1;
        const double dbias = 4503599627370496.0 + (double)(1 << (logB - 1));
        const cplx *brk = a.brk[party];
        const size_t per_idx = (size_t)4 * l * H;
        const uint32_t *at_src = a.step_mode ? a.tilde + unit * ELL : a.tilde + (size_t)gate * a.lwe_words + 1 + (size_t)party * a.n;
        const int nsteps = a.step_mode ? 1 : (ELL == 1 ? a.n : a.d);
        const int brv6t = (int)(__brev((unsigned)t) >> 26);

        for (int step = 0; step < nsteps; step++) {
            uint32_t atv[ELL];
            bool any = false;
#pragma unroll
            for (int b = 0; b < ELL; b++) { atv[b] = at_src[(a.step_mode ? 0 : step * ELL) + b]; any |= atv[b] > 0; }
            if (!any) continue;
            const int idx = (a.step_mode ? a.step_idx : step) * ELL;
            const cplx *kidx = brk + (size_t)idx * per_idx + t;
            cplx m1v[ELL];
#pragma unroll
            for (int b = 0; b < ELL; b++) m1v[b] = __ldg(&a.tb.emono[((4 * brv6t + 1) * atv[b]) & 4095]);

            for (int dg = 0; dg < 2 * l; dg++) {
                const uint32_t src = tm + (dg < l ? TM_ACC_B : TM_ACC_A);
                const int sh = bit + (l - 1 - (dg < l ? dg : dg - l)) * logB;
                cplx x[16];
                {   // gadget digit of 32 coefficients -> 16 complex points (decomposition as in the shared-memory variant)
                    uint32_t lo[32], hi[32];                   // coefficient m at columns 2m, 2m+1; m + 16 = upper half (n + H)
#pragma unroll
                    for (int c = 0; c < 4; c++) {
                        uint32_t v[16];
                        tm_ld16(src + 16 * c, v);
                        tm_wait_ld();
                        tm_pin16(v);
#pragma unroll
                        for (int i = 0; i < 8; i++) { lo[8 * c + i] = v[2 * i]; hi[8 * c + i] = v[2 * i + 1]; }
                    }
#pragma unroll
                    for (int m = 0; m < 16; m++) {
                        const uint64_t v0 = (((uint64_t)hi[m] << 32) | lo[m]) + cadd, v1 = (((uint64_t)hi[m + 16] << 32) | lo[m + 16]) + cadd;
                        const uint32_t f0 = (uint32_t)(v0 >> sh) & mask, f1 = (uint32_t)(v1 >> sh) & mask;
                        x[m] = make_double2(__hiloint2double(0x43300000, (int)f0) - dbias, dbias - __hiloint2double(0x43300000, (int)f1));
                    }
                }
                const cplx *kb = kidx + (size_t)(dg * 2) * H, *ka = kb + H;
                if (ELL == 1) {
                    cplx wb[16], wa[16];
                    // key values are requested before the passes that precede their use
                    fft_fwd2(x, xa, xc, tw2, tw8, tw9e, t, unit_l,
                             [&]() {
#pragma unroll
                                 for (int e = 0; e < 16; e++) wb[e] = __ldg(kb + e * UT);
                             },
                             [&]() {
#pragma unroll
                                 for (int e = 0; e < 16; e++) wa[e] = __ldg(ka + e * UT);
                             });
#pragma unroll
                    for (int c = 0; c < 4; c++) {
                        cplx zb[4], za[4];
                        if (dg == 0) {
#pragma unroll
                            for (int i = 0; i < 4; i++) { zb[i] = cmul_f(x[4 * c + i], wb[4 * c + i]); za[i] = cmul_f(x[4 * c + i], wa[4 * c + i]); }
                        } else {
                            tm_ld_c4(tm + TM_TACC_B + 16 * c, zb);
                            tm_ld_c4(tm + TM_TACC_A + 16 * c, za);
#pragma unroll
                            for (int i = 0; i < 4; i++) { zb[i] = cmac_f(zb[i], x[4 * c + i], wb[4 * c + i]); za[i] = cmac_f(za[i], x[4 * c + i], wa[4 * c + i]); }
                        }
                        tm_st_c4(tm + TM_TACC_B + 16 * c, zb);
                        tm_st_c4(tm + TM_TACC_A + 16 * c, za);
                    }
                } else {
                    fft_fwd2(x, xa, xc, tw2, tw8, tw9e, t, unit_l, []() {}, []() {});
                    // block: fold the monomials of the block's key bits into the keys (see k_phase1)
#pragma unroll
                    for (int c = 0; c < 4; c++) {
                        cplx kv[ELL][2][4];
#pragma unroll
                        for (int b = 0; b < ELL; b++)
#pragma unroll
                            for (int i = 0; i < 4; i++) {
                                kv[b][0][i] = __ldg(kb + b * per_idx + (4 * c + i) * UT);
                                kv[b][1][i] = __ldg(ka + b * per_idx + (4 * c + i) * UT);
                            }
                        cplx zb[4], za[4];
                        if (dg != 0) { tm_ld_c4(tm + TM_TACC_B + 16 * c, zb); tm_ld_c4(tm + TM_TACC_A + 16 * c, za); }
#pragma unroll
                        for (int i = 0; i < 4; i++) {
                            const int e = 4 * c + i;
                            const int b4 = ((e & 1) << 3) | ((e & 2) << 1) | ((e & 4) >> 1) | ((e & 8) >> 3);
                            cplx kcb = make_double2(0.0, 0.0), kca = kcb;
#pragma unroll
                            for (int b = 0; b < ELL; b++) {
                                if (atv[b] == 0) continue;
                                cplx mo = cmul_f(m1v[b], c_e16[(atv[b] * b4) & 15]);
                                mo.x -= 1.0 / H;
                                kcb = cmac_f(kcb, mo, kv[b][0][i]);
                                kca = cmac_f(kca, mo, kv[b][1][i]);
                            }
                            if (dg == 0) { zb[i] = cmul_f(x[e], kcb); za[i] = cmul_f(x[e], kca); }
                            else { zb[i] = cmac_f(zb[i], x[e], kcb); za[i] = cmac_f(za[i], x[e], kca); }
                        }
                        tm_st_c4(tm + TM_TACC_B + 16 * c, zb);
                        tm_st_c4(tm + TM_TACC_A + 16 * c, za);
                    }
                }
                tm_wait_st();
            }
            // both outputs: (x (X^a - 1)/H for ELL == 1) -> inverse transform -> round -> acc +=
#pragma unroll 1
            for (int pz = 0; pz < 2; pz++) {
                cplx y[16];
#pragma unroll
                for (int c = 0; c < 4; c++) {
                    cplx z[4];
                    tm_ld_c4(tm + (pz == 0 ? TM_TACC_B : TM_TACC_A) + 16 * c, z);
#pragma unroll
                    for (int i = 0; i < 4; i++) y[4 * c + i] = z[i];
                }
                if (ELL == 1) {
#pragma unroll
                    for (int e = 0; e < 16; e++) {
                        const int b4 = ((e & 1) << 3) | ((e & 2) << 1) | ((e & 4) >> 1) | ((e & 8) >> 3);
                        cplx mo = cmul_f(m1v[0], c_e16[(atv[0] * b4) & 15]);
                        mo.x -= 1.0 / H;
                        y[e] = cmul_f(mo, y[e]);
                    }
                }
                fft_inv2(y, xa, xc, tw2, tw8, tw9e, t, unit_l);
                const uint32_t dst = tm + (pz == 0 ? TM_ACC_B : TM_ACC_A);
#pragma unroll
                for (int c = 0; c < 4; c++) {              // columns 16c..: coefficients 8c..8c+7 (m = ci & 15, upper half = ci >> 4)
                    uint32_t v[16];
                    tm_ld16(dst + 16 * c, v);
                    tm_wait_ld();
                    tm_pin16(v);
#pragma unroll
                    for (int i = 0; i < 8; i++) {
                        const int ci = 8 * c + i, m = ci & 15;
                        const uint64_t add = d2torus(ci < 16 ? y[m].x : -y[m].y);
                        const uint64_t w = (((uint64_t)v[2 * i + 1] << 32) | v[2 * i]) + add;
                        v[2 * i] = (uint32_t)w; v[2 * i + 1] = (uint32_t)(w >> 32);
                    }
                    tm_st16(dst + 16 * c, v);
                }
                tm_wait_st();
            }
        }

        if (!a.step_mode) {            // fftto!(tacc, acc): bootstrapping.jl:441
            cplx *out = a.lev_out + unit * 2 * H;
#pragma unroll 1
            for (int pz = 0; pz < 2; pz++) {
                cplx x[16];
                uint32_t lo[32], hi[32];
#pragma unroll
                for (int c = 0; c < 4; c++) {
                    uint32_t v[16];
                    tm_ld16(tm + pz * 64 + 16 * c, v);
                    tm_wait_ld();
                    tm_pin16(v);
#pragma unroll
                    for (int i = 0; i < 8; i++) { lo[8 * c + i] = v[2 * i]; hi[8 * c + i] = v[2 * i + 1]; }
                }
#pragma unroll
                for (int m = 0; m < 16; m++) {
                    const uint64_t v0 = ((uint64_t)hi[m] << 32) | lo[m], v1 = ((uint64_t)hi[m + 16] << 32) | lo[m + 16];
                    x[m] = make_double2(__ll2double_rn((long long)v0), __ll2double_rn((long long)((uint64_t)0 - v1)));
                }
                fft_fwd2(x, xa, xc, tw2, tw8, tw9e, t, unit_l, []() {}, []() {});
#pragma unroll
                for (int e = 0; e < 16; e++) out[(size_t)pz * H + 16 * t + e] = x[e];
            }
        } else {
            uint64_t *dst = a.acc_io + unit * 2 * N;
#pragma unroll
            for (int pz = 0; pz < 2; pz++)
#pragma unroll
                for (int c = 0; c < 4; c++) {
                    uint32_t v[16];
                    tm_ld16(tm + pz * 64 + 16 * c, v);
                    tm_wait_ld();
                    tm_pin16(v);
#pragma unroll
                    for (int i = 0; i < 8; i++) {
                        const int ci = 8 * c + i;
                        dst[pz * N + t + 64 * (ci & 15) + (ci >> 4) * H] = ((uint64_t)v[2 * i + 1] << 32) | v[2 * i];
                    }
                }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" :: "r"(*tm_base_s));
}

// (X^a - 1)/H at the 16 slots of a thread without 16 complex multiplications.  Slot 16t + e sits at the evaluation point
// zeta_e with zeta_e^a = m1 * eps16^(a * brv4(e)), m1 = exp(-i*pi*(4*brv6(t)+1)*a/N)/H, eps16 = exp(-i*pi/8), and
//   eps16^(a*brv4(e)) = (-i)^(a*(2*e0 + e1)) * eps16^(2a*e2) * eps16^(a*e3)            (e = e0 + 2*e1 + 4*e2 + 8*e3):
// four base values (three multiplications) and, per slot, a quarter-turn rotation = swap + sign flips in the integer pipe.
// The swap happens only for odd a (and e1 = 1), which is uniform over the unit: SW selects the code version.
// Used by the non-block kernel; in the block kernel the extra live values cost more in spills than they save (measured).
struct Mono16 {
    cplx v[4];            // index e3 + 2*e2
    uint32_t sx[4], sy[4];  // sign masks per class c = 2*e0 + e1
    __device__ __forceinline__ void init(cplx m1, uint32_t a) {
        const cplx E1 = c_e16[a & 15], E2 = c_e16[(2 * a) & 15];
        v[0] = m1; v[1] = cmul_f(m1, E1); v[2] = cmul_f(m1, E2); v[3] = cmul_f(v[2], E1);
#pragma unroll
        for (int c = 0; c < 4; c++) {
            const uint32_t q = (a * c) & 3;                 // multiply by (-i)^q
            sx[c] = (q >> 1) << 31;
            sy[c] = (((q + 1) >> 1) & 1) << 31;
        }
    }
    template <bool SW> __device__ __forceinline__ cplx at(int e) const {
        const int e0 = e & 1, e1 = (e >> 1) & 1, c = 2 * e0 + e1;
        const cplx b = v[((e >> 3) & 1) + 2 * ((e >> 2) & 1)];
        double x = (SW && e1) ? b.y : b.x, y = (SW && e1) ? b.x : b.y;
        if (c) {
            x = __hiloint2double(__double2hiint(x) ^ (int)sx[c], __double2loint(x));
            y = __hiloint2double(__double2hiint(y) ^ (int)sy[c], __double2loint(y));
        }
        return make_double2(x - 1.0 / H, y);
    }
};

// ======================================================================================================
// TMA variant (default): the TMEM kernel above with the bootstrapping key streamed ONCE PER CTA.
// The four units of a CTA are chosen from the same party, so they consume the same key polynomials in the same order
// (only their rotations differ).  A ninth warp is the producer: one thread walks the tile sequence
//   (step, [bit], digit, comp)  ->  16 KiB polynomial in thread order
// and issues `cp.async.bulk` copies into a 5-slot shared-memory ring guarded by full/empty mbarriers; the 256
// consumer threads read their 16 values of a tile with conflict-free 16-byte loads and release the slot.
// Effect: L2 -> SM key traffic drops 4x (it was 1.3 TB per 4096-gate launch, 5.4 TB/s), key values no longer occupy
// 128 registers per thread, and their latency is hidden by the ring instead of by the scheduler.
constexpr int RING_BYTES = 5 * H * 16;                                     // 80 KiB of key tiles in flight
// tile = one polynomial (16 KiB, 5 slots) for the plain kernel, half a polynomial (slots e < 8 / e >= 8 of every thread,
// 8 KiB, 10 slots) for the block kernel, whose fold then needs 8 + 8 instead of 16 + 16 complex accumulators (quarter
// tiles were measured too: 786 ms against 672 ms at KMS8 block, the barrier traffic doubles)
template <int ELL> struct TileCfg { static constexpr int TILE = ELL == 1 ? H : H / 2, RING = RING_BYTES / (TILE * 16); };
constexpr int CTA_TMA = CTA + 128;                                   // 2 consumer warpgroups + 1 producer warpgroup
// `setmaxnreg` only redistributes the registers the CTA was launched with: (launch registers) x warps must cover the
// re-split, or the last `setmaxnreg.inc` waits forever.  Launch: 12 warps x 168; after: 8 x 232 + 4 x 24.
constexpr int TMA_LAUNCH_REGS = 168, TMA_CONSUMER_REGS = 232, TMA_PRODUCER_REGS = 24;
static_assert((CTA / 32) * TMA_CONSUMER_REGS + 4 * TMA_PRODUCER_REGS <= (CTA_TMA / 32) * TMA_LAUNCH_REGS, "setmaxnreg over-subscription");
static_assert(TMA_LAUNCH_REGS * CTA_TMA <= 65536, "launch registers exceed the register file");
constexpr size_t SMEM_BYTES_TMA = U * SMEM_UNIT_TM + (size_t)(128 + 128 + 256) * 16 + (size_t)RING_BYTES + 512;

__device__ __forceinline__ void mb_init(uint64_t *bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"((uint32_t)__cvta_generic_to_shared(bar)), "r"(count));
}
__device__ __forceinline__ void mb_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"((uint32_t)__cvta_generic_to_shared(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mb_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"((uint32_t)__cvta_generic_to_shared(bar)) : "memory");
}
// mbarrier wait: try_wait with a suspend-time hint parks the thread in hardware until the phase completes (or the hint expires)
// instead of spinning -- a spinning producer took a quarter of its scheduler's issue slots (the highest warp id wins
// arbitration).  Bounded: a protocol error (a tile nobody releases, a setmaxnreg over-subscription upstream) traps after
// 2^22 expired hints instead of hanging the device until the watchdog.
__device__ __forceinline__ bool mb_try(uint64_t *bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n.reg .pred p;\n"
                 "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
                 "selp.u32 %0, 1, 0, p;\n}" : "=r"(ok) : "r"((uint32_t)__cvta_generic_to_shared(bar)), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ bool mb_try_hint(uint64_t *bar, uint32_t parity, uint32_t ns) {
    uint32_t ok;
    asm volatile("{\n.reg .pred p;\n"
                 "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n"
                 "selp.u32 %0, 1, 0, p;\n}" : "=r"(ok) : "r"((uint32_t)__cvta_generic_to_shared(bar)), "r"(parity), "r"(ns) : "memory");
    return ok != 0;
}
// consumer-side wait: tight try_wait loop (try_wait + one predicated branch on the success path; measured faster than the
// suspended form when the data is usually there: key switch 9.8 vs 10.3 ms).
// Built with -DMKTFHE_DEBUG_SPIN (MKTFHE_DEBUG_SPIN=1 python -m mktfhe_b200.build --force) every wait is BOUNDED: 2^28 failed polls
// (> 1 s) trap instead of hanging the device until the watchdog -- the build to use while changing a ring or token protocol.
// The counter and the trap cost 3-6 % in the kernels that poll often (key switch 9.8 -> 10.1 ms, KMS8 block phase 1 672 -> 715 ms:
// the trap makes the loop a divergence point), so production builds spin without them.
__device__ __forceinline__ void mbs_spin(uint32_t bar, uint32_t parity) {
#ifdef MKTFHE_DEBUG_SPIN
    asm volatile("{\n.reg .pred p, q;\n.reg .u32 cnt;\nmov.u32 cnt, 0;\nWAIT_%=:\n"
                 "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
                 "@p bra DONE_%=;\n"
                 "add.u32 cnt, cnt, 1;\nsetp.gt.u32 q, cnt, 268435456;\n@q trap;\n"
                 "bra WAIT_%=;\nDONE_%=:\n}" :: "r"(bar), "r"(parity) : "memory");
#else
    asm volatile("{\n.reg .pred p;\nWAIT_%=:\n"
                 "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
                 "@p bra DONE_%=;\nbra WAIT_%=;\nDONE_%=:\n}" :: "r"(bar), "r"(parity) : "memory");
#endif
}
__device__ __forceinline__ void mb_wait(uint64_t *bar, uint32_t parity) { mbs_spin((uint32_t)__cvta_generic_to_shared(bar), parity); }
// producer-side wait: suspended in hardware between polls (a spinning producer took a quarter of its scheduler's issue slots)
__device__ __forceinline__ void mb_wait_suspend(uint64_t *bar, uint32_t parity) {
    uint32_t spins = 0;
    while (!mb_try_hint(bar, parity, 20000u))
        if (++spins > (1u << 22)) __trap();
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"((uint32_t)__cvta_generic_to_shared(dst)), "l"(src), "r"(bytes), "r"((uint32_t)__cvta_generic_to_shared(bar)) : "memory");
}

template <int ELL>
__global__ void __launch_bounds__(CTA_TMA, 1) k_phase1_tma(const Args a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int TILE = TileCfg<ELL>::TILE, RING = TileCfg<ELL>::RING, HALVES = H / TILE;
    const int tid = threadIdx.x, warp = tid >> 5, unit_l = tid / UT, t = tid % UT;
    cplx *tw2 = reinterpret_cast<cplx *>(smem_raw + U * SMEM_UNIT_TM), *tw8 = tw2 + 128, *tw9e = tw8 + 128;
    cplx *ring = tw9e + 256;
    uint64_t *full = reinterpret_cast<uint64_t *>(ring + (size_t)RING * TILE), *empty = full + RING;
    uint32_t *tm_base_s = reinterpret_cast<uint32_t *>(empty + RING);
    for (int i = tid; i < 256; i += CTA_TMA) { if (i < 128) { tw2[i] = a.tb.t2[i]; tw8[i] = a.tb.t8[i]; } tw9e[i] = a.tb.t9[i]; }
    if (tid == 0) {
        for (int s = 0; s < RING; s++) { mb_init(&full[s], 1); mb_init(&empty[s], CTA); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" :: "r"((uint32_t)__cvta_generic_to_shared(tm_base_s)));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");

    // ---- CTA -> (party, units): party 0 has one RLEV row per gate, the others l_lev (bootstrapping.jl:400)
    int party, rows;
    size_t u0;                                                     // first unit (within the party) of this CTA
    size_t gates;
    if (!a.step_mode) {
        gates = a.units / a.R;
        const size_t ctas0 = (gates + U - 1) / U, ctasp = (gates * a.l_lev + U - 1) / U;
        if (blockIdx.x < ctas0) { party = 0; rows = 1; u0 = (size_t)blockIdx.x * U; }
        else { party = 1 + (int)((blockIdx.x - ctas0) / ctasp); rows = a.l_lev; u0 = ((blockIdx.x - ctas0) % ctasp) * U; }
    } else { gates = a.units; party = a.step_party; rows = 1; u0 = (size_t)blockIdx.x * U; }
    const int l = a.l;
    const size_t per_idx = (size_t)4 * l * H;
    const int nsteps = a.step_mode ? 1 : (ELL == 1 ? a.n : a.d);
    const cplx *brk = a.brk[party];
    const uint32_t ntiles = (uint32_t)nsteps * 2 * l * HALVES * ELL * 2;

    // The register file is partitioned per scheduler (16K registers each), so a ninth warp at 200+ registers does not
    // fit: the CTA is launched with 12 warps at <= 168 registers and the warpgroups re-split the pool (setmaxnreg):
    // the producer warpgroup keeps 24 registers per thread, the two consumer warpgroups take 232.
    if (warp >= U * 2) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 24;");
        // ---- producer: tile n = (((step * 2l + dg) * ELL + b) * 2 + comp)
        if (tid == CTA) {
            for (uint32_t n = 0; n < ntiles; n++) {
                const int slot = n % RING;
#ifdef TMA_PRODUCER_SPIN
                if (n >= RING) mb_wait(&empty[slot], ((n / RING) - 1) & 1);
#else
                if (n >= RING) mb_wait_suspend(&empty[slot], ((n / RING) - 1) & 1);
#endif
                // tile n = ((((step * 2l + dg) * HALVES + half) * ELL + b) * 2 + comp)
                const uint32_t comp = n & 1, b = (n >> 1) % ELL, hf = ((n >> 1) / ELL) % HALVES;
                const uint32_t dg = ((n >> 1) / ELL / HALVES) % (2 * l), step = (n >> 1) / ELL / HALVES / (2 * l);
                const int idx = (a.step_mode ? a.step_idx : (int)step) * ELL + (int)b;
                mb_expect_tx(&full[slot], TILE * 16);
                bulk_g2s(ring + (size_t)slot * TILE, brk + (size_t)idx * per_idx + (size_t)(dg * 2 + comp) * H + (size_t)hf * TILE, TILE * 16, &full[slot]);
            }
        }
    } else {
        asm volatile("setmaxnreg.inc.sync.aligned.u32 232;");
        // ---- consumers
        const uint32_t tm = *tm_base_s + ((uint32_t)(32 * (warp & 3)) << 16) + 256u * (uint32_t)(warp >> 2);
        cplx *xa = reinterpret_cast<cplx *>(smem_raw + unit_l * SMEM_UNIT_TM), *xc = xa + XB_LEN;
        const size_t up = u0 + unit_l;                             // unit index inside the party
        const bool live = up < gates * (size_t)rows;
        const int gate = live ? (int)(up / rows) : 0, row = live ? (int)(up % rows) : 0;
        const size_t unit_out = a.step_mode ? up : (size_t)gate * a.R + (party == 0 ? 0 : 1 + (size_t)(party - 1) * a.l_lev + row);
        // the plain kernel keeps its pass-2/3 twiddles in registers; the block kernel has none to spare (measured: +23 % time)
        TwRegs twr;
        if constexpr (ELL == 1) twr.load(tw2, tw8, tw9e, t);
        auto fwd = [&](cplx (&v)[16]) {
            if constexpr (ELL == 1) fft_fwd2r(v, xa, xc, twr, t, unit_l);
            else fft_fwd2(v, xa, xc, tw2, tw8, tw9e, t, unit_l, []() {}, []() {});
        };
        auto inv = [&](cplx (&v)[16]) {
            if constexpr (ELL == 1) fft_inv2r(v, xa, xc, twr, t, unit_l);
            else fft_inv2(v, xa, xc, tw2, tw8, tw9e, t, unit_l);
        };
        uint32_t tile_n = 0;                                       // next tile this thread will consume
        auto tile_wait = [&]() -> const cplx * {
            const int slot = tile_n % RING;
            mb_wait(&full[slot], (tile_n / RING) & 1);
            return ring + (size_t)slot * TILE + t;
        };
        auto tile_done = [&]() { mb_arrive(&empty[tile_n % RING]); tile_n++; };

        if (live) {
            if (!a.step_mode) {
                uint32_t z[16];
#pragma unroll
                for (int i = 0; i < 16; i++) z[i] = 0u;
                const uint64_t gv = t == 0 ? (uint64_t)1 << (64 - (row + 1) * a.logB_lev) : 0;   // bootstrapping.jl:402-408
#pragma unroll
                for (int c = 0; c < 128; c += 16) {            // tcgen05.st is warp-collective: same instruction on every lane
                    z[0] = c == 0 ? (uint32_t)gv : 0u;
                    z[1] = c == 0 ? (uint32_t)(gv >> 32) : 0u;
                    tm_st16(tm + c, z);
                }
            } else {
                const uint64_t *src = a.acc_io + up * 2 * N;
#pragma unroll
                for (int pz = 0; pz < 2; pz++)
#pragma unroll
                    for (int c = 0; c < 4; c++) {
                        uint32_t v[16];
#pragma unroll
                        for (int i = 0; i < 8; i++) {
                            const int ci = 8 * c + i;
                            const uint64_t w = src[pz * N + t + 64 * (ci & 15) + (ci >> 4) * H];
                            v[2 * i] = (uint32_t)w; v[2 * i + 1] = (uint32_t)(w >> 32);
                        }
                        tm_st16(tm + pz * 64 + 16 * c, v);
                    }
            }
            tm_wait_st();
        }
        const int logB = a.logB;
        const int bit = 64 - l * logB;
        uint64_t cadd = (uint64_t)1 << (bit - 1);
        for (int j = 0; j < l; j++) cadd += (uint64_t)1 << (bit + j * logB + logB - 1);
        const uint32_t mask = (1u << logB) - 1;
        const double dbias = 4503599627370496.0 + (double)(1 << (logB - 1));
        const uint32_t *at_src = a.step_mode ? a.tilde + up * ELL : a.tilde + (size_t)gate * a.lwe_words + 1 + (size_t)party * a.n;
        const int brv6t = (int)(__brev((unsigned)t) >> 26);

        for (int step = 0; step < nsteps; step++) {
            uint32_t atv[ELL];
            bool any = false;
#pragma unroll
            for (int b = 0; b < ELL; b++) { atv[b] = live ? at_src[(a.step_mode ? 0 : step * ELL) + b] : 0u; any |= atv[b] > 0; }
            if (!any) {                                        // :413 / dead unit: keep the ring moving, compute nothing
                for (int i = 0; i < 2 * l * HALVES * ELL * 2; i++) { tile_wait(); tile_done(); }
                continue;
            }
            cplx m1v[ELL];
#pragma unroll
            for (int b = 0; b < ELL; b++) m1v[b] = __ldg(&a.tb.emono[((4 * brv6t + 1) * atv[b]) & 4095]);

            for (int dg = 0; dg < 2 * l; dg++) {
                const uint32_t src = tm + (dg < l ? TM_ACC_B : TM_ACC_A);
                const int sh = bit + (l - 1 - (dg < l ? dg : dg - l)) * logB;
                cplx x[16];
                {
                    uint32_t lo[32], hi[32];
                    {
                        uint32_t v[4][16];
#pragma unroll
                        for (int c = 0; c < 4; c++) tm_ld16(src + 16 * c, v[c]);       // four loads in flight, one wait
                        tm_wait_ld();
#pragma unroll
                        for (int c = 0; c < 4; c++) {
                            tm_pin16(v[c]);
#pragma unroll
                            for (int i = 0; i < 8; i++) { lo[8 * c + i] = v[c][2 * i]; hi[8 * c + i] = v[c][2 * i + 1]; }
                        }
                    }
#pragma unroll
                    for (int m = 0; m < 16; m++) {
                        const uint64_t v0 = (((uint64_t)hi[m] << 32) | lo[m]) + cadd, v1 = (((uint64_t)hi[m + 16] << 32) | lo[m + 16]) + cadd;
                        const uint32_t f0 = (uint32_t)(v0 >> sh) & mask, f1 = (uint32_t)(v1 >> sh) & mask;
                        x[m] = make_double2(__hiloint2double(0x43300000, (int)f0) - dbias, dbias - __hiloint2double(0x43300000, (int)f1));
                    }
                }
                fwd(x);
                if (ELL == 1) {
                    // both key tiles stay in the ring while the four chunks are processed: key values are read four at a
                    // time right before use instead of occupying 128 registers
                    const cplx *kb = tile_wait();
                    const int slot_b = tile_n % RING;
                    tile_n++;
                    const cplx *ka = tile_wait();
#pragma unroll
                    for (int c = 0; c < 4; c++) {
                        cplx zb[4], za[4], kcb[4], kca[4];
#pragma unroll
                        for (int i = 0; i < 4; i++) { kcb[i] = kb[(4 * c + i) * UT]; kca[i] = ka[(4 * c + i) * UT]; }
                        if (dg == 0) {
#pragma unroll
                            for (int i = 0; i < 4; i++) { zb[i] = cmul_f(x[4 * c + i], kcb[i]); za[i] = cmul_f(x[4 * c + i], kca[i]); }
                        } else {
                            uint32_t vb[16], va[16];                    // both accumulator chunks behind one wait
                            tm_ld16(tm + TM_TACC_B + 16 * c, vb);
                            tm_ld16(tm + TM_TACC_A + 16 * c, va);
                            tm_wait_ld();
                            tm_pin16(vb); tm_pin16(va);
#pragma unroll
                            for (int i = 0; i < 4; i++) {
                                zb[i] = make_double2(__hiloint2double((int)vb[4 * i + 1], (int)vb[4 * i]), __hiloint2double((int)vb[4 * i + 3], (int)vb[4 * i + 2]));
                                za[i] = make_double2(__hiloint2double((int)va[4 * i + 1], (int)va[4 * i]), __hiloint2double((int)va[4 * i + 3], (int)va[4 * i + 2]));
                                zb[i] = cmac_f(zb[i], x[4 * c + i], kcb[i]); za[i] = cmac_f(za[i], x[4 * c + i], kca[i]);
                            }
                        }
                        tm_st_c4(tm + TM_TACC_B + 16 * c, zb);
                        tm_st_c4(tm + TM_TACC_A + 16 * c, za);
                    }
                    mb_arrive(&empty[slot_b]);
                    tile_done();
                } else {
                    // block: fold the monomials of the block's key bits into the keys (see k_phase1), part of the thread's
                    // slots at a time: Sum_bit mono_bit * K_bit for e in [EP*part, EP*part + EP), then the multiply-accumulate
                    constexpr int EP = 16 / HALVES;                 // slots per key tile and thread
#pragma unroll
                    for (int hf = 0; hf < HALVES; hf++) {
                        cplx kcb[EP], kca[EP];
#pragma unroll
                        for (int e = 0; e < EP; e++) kcb[e] = kca[e] = make_double2(0.0, 0.0);
#pragma unroll
                        for (int b = 0; b < ELL; b++) {
                            const cplx *kb = tile_wait();
                            const int slot_b = tile_n % RING;
                            tile_n++;                               // hold the .b tile while the .a tile is awaited
                            const cplx *ka = tile_wait();
                            if (atv[b] != 0) {
#pragma unroll
                                for (int eh = 0; eh < EP; eh++) {
                                    const int e = EP * hf + eh;
                                    const int b4 = ((e & 1) << 3) | ((e & 2) << 1) | ((e & 4) >> 1) | ((e & 8) >> 3);
                                    cplx mo = cmul_f(m1v[b], c_e16[(atv[b] * b4) & 15]);
                                    mo.x -= 1.0 / H;
                                    kcb[eh] = cmac_f(kcb[eh], mo, kb[eh * UT]);
                                    kca[eh] = cmac_f(kca[eh], mo, ka[eh * UT]);
                                }
                            }
                            mb_arrive(&empty[slot_b]);
                            tile_done();
                        }
#pragma unroll
                        for (int c2 = 0; c2 < EP / 4; c2++) {
                            const int c = (EP / 4) * hf + c2;
                            cplx zb[4], za[4];
                            if (dg == 0) {
#pragma unroll
                                for (int i = 0; i < 4; i++) { zb[i] = cmul_f(x[4 * c + i], kcb[4 * c2 + i]); za[i] = cmul_f(x[4 * c + i], kca[4 * c2 + i]); }
                            } else {
                                uint32_t vb[16], va[16];                // both accumulator chunks behind one wait
                                tm_ld16(tm + TM_TACC_B + 16 * c, vb);
                                tm_ld16(tm + TM_TACC_A + 16 * c, va);
                                tm_wait_ld();
                                tm_pin16(vb); tm_pin16(va);
#pragma unroll
                                for (int i = 0; i < 4; i++) {
                                    zb[i] = make_double2(__hiloint2double((int)vb[4 * i + 1], (int)vb[4 * i]), __hiloint2double((int)vb[4 * i + 3], (int)vb[4 * i + 2]));
                                    za[i] = make_double2(__hiloint2double((int)va[4 * i + 1], (int)va[4 * i]), __hiloint2double((int)va[4 * i + 3], (int)va[4 * i + 2]));
                                    zb[i] = cmac_f(zb[i], x[4 * c + i], kcb[4 * c2 + i]); za[i] = cmac_f(za[i], x[4 * c + i], kca[4 * c2 + i]);
                                }
                            }
                            tm_st_c4(tm + TM_TACC_B + 16 * c, zb);
                            tm_st_c4(tm + TM_TACC_A + 16 * c, za);
                        }
                    }
                }
                tm_wait_st();
            }
            // both outputs: (x (X^a - 1)/H for ELL == 1) -> inverse transform -> round -> acc +=
#pragma unroll 1
            for (int pz = 0; pz < 2; pz++) {
                cplx y[16];
                {
                    uint32_t v[4][16];
#pragma unroll
                    for (int c = 0; c < 4; c++) tm_ld16(tm + (pz == 0 ? TM_TACC_B : TM_TACC_A) + 16 * c, v[c]);
                    tm_wait_ld();
#pragma unroll
                    for (int c = 0; c < 4; c++) {
                        tm_pin16(v[c]);
#pragma unroll
                        for (int i = 0; i < 4; i++)
                            y[4 * c + i] = make_double2(__hiloint2double((int)v[c][4 * i + 1], (int)v[c][4 * i]), __hiloint2double((int)v[c][4 * i + 3], (int)v[c][4 * i + 2]));
                    }
                }
                if (ELL == 1) {
                    Mono16 mg;
                    mg.init(m1v[0], atv[0]);
                    if (atv[0] & 1) {
#pragma unroll
                        for (int e = 0; e < 16; e++) y[e] = cmul_f(mg.at<true>(e), y[e]);
                    } else {
#pragma unroll
                        for (int e = 0; e < 16; e++) y[e] = cmul_f(mg.at<false>(e), y[e]);
                    }
                }
                inv(y);
                const uint32_t dst = tm + (pz == 0 ? TM_ACC_B : TM_ACC_A);
                uint32_t v[4][16];
                tm_ld16(dst, v[0]);
#pragma unroll
                for (int c = 0; c < 4; c++) {
                    tm_wait_ld();
                    tm_pin16(v[c]);
                    if (c < 3) tm_ld16(dst + 16 * (c + 1), v[c + 1]);           // next chunk in flight during the rounding
#pragma unroll
                    for (int i = 0; i < 8; i++) {
                        const int ci = 8 * c + i, m = ci & 15;
                        const uint64_t add = d2torus(ci < 16 ? y[m].x : -y[m].y);
                        const uint64_t w = (((uint64_t)v[c][2 * i + 1] << 32) | v[c][2 * i]) + add;
                        v[c][2 * i] = (uint32_t)w; v[c][2 * i + 1] = (uint32_t)(w >> 32);
                    }
                    tm_st16(dst + 16 * c, v[c]);
                }
                tm_wait_st();
            }
        }

        if (live) {
            if (!a.step_mode) {            // fftto!(tacc, acc): bootstrapping.jl:441
                cplx *out = a.lev_out + unit_out * 2 * H;
#pragma unroll 1
                for (int pz = 0; pz < 2; pz++) {
                    cplx x[16];
                    uint32_t lo[32], hi[32];
#pragma unroll
                    for (int c = 0; c < 4; c++) {
                        uint32_t v[16];
                        tm_ld16(tm + pz * 64 + 16 * c, v);
                        tm_wait_ld();
                        tm_pin16(v);
#pragma unroll
                        for (int i = 0; i < 8; i++) { lo[8 * c + i] = v[2 * i]; hi[8 * c + i] = v[2 * i + 1]; }
                    }
#pragma unroll
                    for (int m = 0; m < 16; m++) {
                        const uint64_t v0 = ((uint64_t)hi[m] << 32) | lo[m], v1 = ((uint64_t)hi[m + 16] << 32) | lo[m + 16];
                        x[m] = make_double2(__ll2double_rn((long long)v0), __ll2double_rn((long long)((uint64_t)0 - v1)));
                    }
                    fwd(x);
                    if (a.lev_fast) {          // for k_phase2: thread order, and the 1/H of phase 2's inverse transforms (exact)
#pragma unroll
                        for (int e = 0; e < 16; e++) out[(size_t)pz * H + e * UT + t] = make_double2(x[e].x * (1.0 / H), x[e].y * (1.0 / H));
                    } else {
#pragma unroll
                        for (int e = 0; e < 16; e++) out[(size_t)pz * H + 16 * t + e] = x[e];
                    }
                }
            } else {
                uint64_t *dst = a.acc_io + up * 2 * N;
#pragma unroll
                for (int pz = 0; pz < 2; pz++)
#pragma unroll
                    for (int c = 0; c < 4; c++) {
                        uint32_t v[16];
                        tm_ld16(tm + pz * 64 + 16 * c, v);
                        tm_wait_ld();
                        tm_pin16(v);
#pragma unroll
                        for (int i = 0; i < 8; i++) {
                            const int ci = 8 * c + i;
                            dst[pz * N + t + 64 * (ci & 15) + (ci >> 4) * H] = ((uint64_t)v[2 * i + 1] << 32) | v[2 * i];
                        }
                    }
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" :: "r"(*tm_base_s));
}


// ---- FAST phase 2 (bootstrapping.jl:448-558) ------------------------------------------------------------------------
// One gate per 64-thread unit, four gates per CTA, parties in sequence as in the reference.  Per party idx and
// component c <= idx:  (tx, ty) = Sum_j D_j(acc_c) * levkey[idx][j];  y_c = ifft(ty);  u_c = Sum_j D_j(y_c) * rlk.d[j];
// v -/+= Sum_j D_j(y_c) * (crs | pubb[c-1])[j];  then w = Sum_j D_j(ifft(v)) * rlk.f[j] and
// acc_c = ifft(tx_c + u_c (+ w.b for c = 0)), acc_{idx+1} = ifft(w.a).
// The four running sums (tx, ty | wb, wa, u, v) live in TMEM (64 columns each per thread), the polynomial being
// re-decomposed stays in registers (the thread that produced coefficient t + 64m also decomposes it), tx_c and u_c wait in
// a global scratch in thread order.  All keys are in thread order [e][t] and carry the 1/H of the inverse transform.
struct P2Args {
    const uint32_t *tilde;        // [B][lwe_words]; only b~ is read
    const cplx *lev;              // [B][R][2][H] from k_phase1_tma with lev_fast
    const cplx *const *rlk;       // [k]: [l_uni][3][H]
    const cplx *const *pubb;      // [k]: [l_uni][H]
    const cplx *crs;              // [l_uni][H]
    Tables tb;
    uint64_t *acc;                // [B][(k+1)][N] out
    cplx *tx, *ty;                // scratch [B][(k+1)][H] each
    int k, l_lev, logB_lev, l_uni, logB_uni, R, lwe_words;
    size_t gates;
};

constexpr uint32_t TM2_TX = 0, TM2_TY = 64, TM2_U = 128, TM2_V = 192;

// digit dg (0 = most significant) of the 32 coefficients held as (lo, hi) words: one add does the rounding and the
// + B/2 at every digit position, then the digit is a bit field (gsw.jl:86-96)
__device__ __forceinline__ void p2_digits(const uint32_t (&lo)[32], const uint32_t (&hi)[32], int dg, int l, int logB, cplx (&x)[16]) {
    const int bit = 64 - l * logB;
    uint64_t cadd = (uint64_t)1 << (bit - 1);
    for (int j = 0; j < l; j++) cadd += (uint64_t)1 << (bit + j * logB + logB - 1);
    const uint32_t mask = (1u << logB) - 1;
    const double dbias = 4503599627370496.0 + (double)(1 << (logB - 1));
    const int sh = bit + (l - 1 - dg) * logB;
#pragma unroll
    for (int m = 0; m < 16; m++) {
        const uint64_t v0 = (((uint64_t)hi[m] << 32) | lo[m]) + cadd, v1 = (((uint64_t)hi[m + 16] << 32) | lo[m + 16]) + cadd;
        const uint32_t f0 = (uint32_t)(v0 >> sh) & mask, f1 = (uint32_t)(v1 >> sh) & mask;
        x[m] = make_double2(__hiloint2double(0x43300000, (int)f0) - dbias, dbias - __hiloint2double(0x43300000, (int)f1));
    }
}
// two TMEM accumulators += x * (ka, kb) (thread-order keys); first_a / first_b overwrite; sb = -1 subtracts the second product
__device__ __forceinline__ void p2_mac(uint32_t tm_a, uint32_t tm_b, const cplx (&x)[16], const cplx *ka, const cplx *kb,
                                       bool first_a, bool first_b, double sb) {
#pragma unroll
    for (int c = 0; c < 4; c++) {
        cplx za[4], zb[4], va[4], vb[4];
#pragma unroll
        for (int i = 0; i < 4; i++) {
            va[i] = __ldg(ka + (4 * c + i) * UT); vb[i] = __ldg(kb + (4 * c + i) * UT);
            za[i] = zb[i] = make_double2(0.0, 0.0);
        }
        uint32_t ra[16], rb[16];
        if (!first_a) tm_ld16(tm_a + 16 * c, ra);
        if (!first_b) tm_ld16(tm_b + 16 * c, rb);
        if (!first_a || !first_b) tm_wait_ld();
        if (!first_a) {
            tm_pin16(ra);
#pragma unroll
            for (int i = 0; i < 4; i++) za[i] = make_double2(__hiloint2double((int)ra[4 * i + 1], (int)ra[4 * i]), __hiloint2double((int)ra[4 * i + 3], (int)ra[4 * i + 2]));
        }
        if (!first_b) {
            tm_pin16(rb);
#pragma unroll
            for (int i = 0; i < 4; i++) zb[i] = make_double2(__hiloint2double((int)rb[4 * i + 1], (int)rb[4 * i]), __hiloint2double((int)rb[4 * i + 3], (int)rb[4 * i + 2]));
        }
#pragma unroll
        for (int i = 0; i < 4; i++) {
            za[i] = cmac_f(za[i], x[4 * c + i], va[i]);
            zb[i] = cmac_f(zb[i], make_double2(sb * x[4 * c + i].x, sb * x[4 * c + i].y), vb[i]);
        }
        tm_st_c4(tm_a + 16 * c, za);
        tm_st_c4(tm_b + 16 * c, zb);
    }
    tm_wait_st();
}
__device__ __forceinline__ void p2_tm_load(uint32_t tm_src, cplx (&y)[16]) {
    uint32_t v[4][16];
#pragma unroll
    for (int c = 0; c < 4; c++) tm_ld16(tm_src + 16 * c, v[c]);
    tm_wait_ld();
#pragma unroll
    for (int c = 0; c < 4; c++) {
        tm_pin16(v[c]);
#pragma unroll
        for (int i = 0; i < 4; i++)
            y[4 * c + i] = make_double2(__hiloint2double((int)v[c][4 * i + 1], (int)v[c][4 * i]), __hiloint2double((int)v[c][4 * i + 3], (int)v[c][4 * i + 2]));
    }
}
__device__ __forceinline__ void p2_tm_store(uint32_t tm_dst, const cplx (&y)[16]) {
#pragma unroll
    for (int c = 0; c < 4; c++) {
        cplx z[4];
#pragma unroll
        for (int i = 0; i < 4; i++) z[i] = y[4 * c + i];
        tm_st_c4(tm_dst + 16 * c, z);
    }
    tm_wait_st();
}

__global__ void __launch_bounds__(CTA, 1) k_phase2(const P2Args a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int tid = threadIdx.x, warp = tid >> 5, unit_l = tid / UT, t = tid % UT;
    cplx *tw2 = reinterpret_cast<cplx *>(smem_raw + U * SMEM_UNIT_TM), *tw8 = tw2 + 128, *tw9e = tw8 + 128;
    uint32_t *tm_base_s = reinterpret_cast<uint32_t *>(tw9e + 256);
    for (int i = tid; i < 256; i += CTA) { if (i < 128) { tw2[i] = a.tb.t2[i]; tw8[i] = a.tb.t8[i]; } tw9e[i] = a.tb.t9[i]; }
    cplx *xa = reinterpret_cast<cplx *>(smem_raw + unit_l * SMEM_UNIT_TM), *xc = xa + XB_LEN;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" :: "r"((uint32_t)__cvta_generic_to_shared(tm_base_s)));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t tm = *tm_base_s + ((uint32_t)(32 * (warp & 3)) << 16) + 256u * (uint32_t)(warp >> 2);

    const size_t gate = (size_t)blockIdx.x * U + unit_l;
    if (gate < a.gates) {
        const int k = a.k, ll = a.l_lev, lu = a.l_uni;
        uint64_t *ACC = a.acc + gate * (size_t)(k + 1) * N;
        cplx *TX = a.tx + gate * (size_t)(k + 1) * H + t, *TY = a.ty + gate * (size_t)(k + 1) * H + t;
        const cplx *LEV = a.lev + gate * (size_t)a.R * 2 * H + t;
        uint32_t lo[32], hi[32];
        cplx x[16];
        // coefficient held in slot m (< 16) is t + 64m, in slot 16 + m it is t + 64m + H
        auto load_poly = [&](const uint64_t *src) {
#pragma unroll
            for (int m = 0; m < 32; m++) { const uint64_t v = src[t + 64 * (m & 15) + (m >> 4) * H]; lo[m] = (uint32_t)v; hi[m] = (uint32_t)(v >> 32); }
        };
        auto store_poly = [&](uint64_t *dst, const cplx (&y)[16]) {
#pragma unroll
            for (int m = 0; m < 16; m++) { dst[t + 64 * m] = d2torus(y[m].x); dst[t + 64 * m + H] = d2torus(-y[m].y); }
        };
        auto to_regs = [&](const cplx (&y)[16]) {                      // native(): arithmetic.jl:6-9
#pragma unroll
            for (int m = 0; m < 16; m++) {
                const uint64_t v0 = d2torus(y[m].x), v1 = d2torus(-y[m].y);
                lo[m] = (uint32_t)v0; hi[m] = (uint32_t)(v0 >> 32); lo[m + 16] = (uint32_t)v1; hi[m + 16] = (uint32_t)(v1 >> 32);
            }
        };
        {   // acc = (test vector, 0 ... 0): bootstrapping.jl:11-23
            const uint32_t tb = a.tilde[gate * a.lwe_words];
            const uint64_t e8 = (uint64_t)1 << 61;
#pragma unroll
            for (int m = 0; m < 32; m++) {
                const int i0 = t + 64 * (m & 15) + (m >> 4) * H;
                const uint32_t i1 = (uint32_t)i0 + 1;
                ACC[i0] = tb <= (uint32_t)N ? (i1 <= tb ? e8 : (uint64_t)0 - e8) : (i1 <= tb - (uint32_t)N ? (uint64_t)0 - e8 : e8);
            }
            // components 1 .. k are written before they are read (component c is first read at idx = c, written at idx = c - 1)
        }
        for (int idx = 0; idx < k; idx++) {
            const cplx *lk = LEV + (size_t)(idx == 0 ? 0 : 1 + (idx - 1) * ll) * 2 * H;
            const cplx *rlk = a.rlk[idx] + t;
            const int iter = idx == 0 ? 1 : ll;                         // :481
            for (int c = 0; c <= idx; c++) {                            // components b, a_1 .. a_idx
                load_poly(ACC + (size_t)c * N);
                for (int j = 0; j < iter; j++) {                        // LEV product with levkey[idx]: :483-499
                    p2_digits(lo, hi, j, ll, a.logB_lev, x);
                    fft_fwd2(x, xa, xc, tw2, tw8, tw9e, t, unit_l, []() {}, []() {});
                    p2_mac(tm + TM2_TX, tm + TM2_TY, x, lk + (size_t)(j * 2) * H, lk + (size_t)(j * 2 + 1) * H, j == 0, j == 0, 1.0);
                }
                p2_tm_load(tm + TM2_TX, x);
#pragma unroll
                for (int e = 0; e < 16; e++) TX[(size_t)c * H + e * UT] = x[e];
                p2_tm_load(tm + TM2_TY, x);
                fft_inv2(x, xa, xc, tw2, tw8, tw9e, t, unit_l);         // y_c: :501-504
                to_regs(x);
                const cplx *kv = c == 0 ? a.crs + t : a.pubb[c - 1] + t;
                for (int j = 0; j < lu; j++) {                          // u and v: :520-535
                    p2_digits(lo, hi, j, lu, a.logB_uni, x);
                    fft_fwd2(x, xa, xc, tw2, tw8, tw9e, t, unit_l, []() {}, []() {});
                    // u restarts for every component; v runs over all components of this party (first term: c = 0, j = 0)
                    p2_mac(tm + TM2_U, tm + TM2_V, x, rlk + (size_t)(j * 3) * H, kv + (size_t)j * H, j == 0, j == 0 && c == 0, c == 0 ? -1.0 : 1.0);
                }
                p2_tm_load(tm + TM2_U, x);
#pragma unroll
                for (int e = 0; e < 16; e++) TY[(size_t)c * H + e * UT] = x[e];
            }
            // v -> coefficient form -> digits -> w: :538-550
            p2_tm_load(tm + TM2_V, x);
            fft_inv2(x, xa, xc, tw2, tw8, tw9e, t, unit_l);
            to_regs(x);
            for (int j = 0; j < lu; j++) {
                p2_digits(lo, hi, j, lu, a.logB_uni, x);
                fft_fwd2(x, xa, xc, tw2, tw8, tw9e, t, unit_l, []() {}, []() {});
                p2_mac(tm + TM2_TX, tm + TM2_TY, x, rlk + (size_t)(j * 3 + 1) * H, rlk + (size_t)(j * 3 + 2) * H, j == 0, j == 0, 1.0);   // wb, wa
            }
            // acc_c = ifft(tx_c + u_c (+ wb)); acc_{idx+1} = ifft(wa): :553-556
            for (int c = 0; c <= idx + 1; c++) {
                if (c == idx + 1) p2_tm_load(tm + TM2_TY, x);
                else {
                    if (c == 0) p2_tm_load(tm + TM2_TX, x);
                    else {
#pragma unroll
                        for (int e = 0; e < 16; e++) x[e] = make_double2(0.0, 0.0);
                    }
#pragma unroll
                    for (int e = 0; e < 16; e++) {
                        const cplx p = TX[(size_t)c * H + e * UT], q = TY[(size_t)c * H + e * UT];
                        x[e].x += p.x + q.x; x[e].y += p.y + q.y;
                    }
                }
                fft_inv2(x, xa, xc, tw2, tw8, tw9e, t, unit_l);
                store_poly(ACC + (size_t)c * N, x);
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" :: "r"(*tm_base_s));
}

// ---- gadget product hook: the building block of FAST phase 2 on caller-supplied inputs --------------------------------------
// out[c] = native(ifft( Sum_j fft(D_j(poly)) (.) key[j][c] ))  for c < ncomp, with the SAME device functions k_phase2 is made of
// (p2_digits, fft_fwd2, multiply-accumulate, fft_inv2, d2torus).  Phase 2 chains such products and re-decomposes their
// outputs, so whole-phase coefficients cannot be compared between two roundings; this hook lets a test feed the ORACLE's
// intermediate polynomial into one product and compare the result within a per-product tolerance
// (bootstrapping.jl:483-499 LEV product, :520-535 u / v, :538-550 w).  keys: [l][ncomp][H] in the reference slot order.
struct GpArgs {
    const uint64_t *polys;      // [B][N]
    const cplx *keys;           // [l][ncomp][H]
    uint64_t *out;              // [B][ncomp][N]
    Tables tb;
    int l, logB, ncomp;
    size_t units;
};
__global__ void __launch_bounds__(CTA, 1) k_gadget_product(const GpArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int tid = threadIdx.x, unit_l = tid / UT, t = tid % UT;
    cplx *tw2 = reinterpret_cast<cplx *>(smem_raw + U * SMEM_UNIT_TM), *tw8 = tw2 + 128, *tw9e = tw8 + 128;
    for (int i = tid; i < 256; i += CTA) { if (i < 128) { tw2[i] = a.tb.t2[i]; tw8[i] = a.tb.t8[i]; } tw9e[i] = a.tb.t9[i]; }
    cplx *xa = reinterpret_cast<cplx *>(smem_raw + unit_l * SMEM_UNIT_TM), *xc = xa + XB_LEN;
    __syncthreads();
    const size_t unit = (size_t)blockIdx.x * U + unit_l;
    if (unit >= a.units) return;
    uint32_t lo[32], hi[32];
    const uint64_t *src = a.polys + unit * N;
#pragma unroll
    for (int m = 0; m < 32; m++) { const uint64_t v = src[t + 64 * (m & 15) + (m >> 4) * H]; lo[m] = (uint32_t)v; hi[m] = (uint32_t)(v >> 32); }
    for (int c = 0; c < a.ncomp; c++) {
        cplx acc[16];
#pragma unroll
        for (int e = 0; e < 16; e++) acc[e] = make_double2(0.0, 0.0);
        for (int j = 0; j < a.l; j++) {
            cplx x[16];
            p2_digits(lo, hi, j, a.l, a.logB, x);
            fft_fwd2(x, xa, xc, tw2, tw8, tw9e, t, unit_l, []() {}, []() {});
            const cplx *k = a.keys + ((size_t)j * a.ncomp + c) * H + 16 * t;          // slot 16t + e
#pragma unroll
            for (int e = 0; e < 16; e++) acc[e] = cmac_f(acc[e], x[e], k[e]);
        }
#pragma unroll
        for (int e = 0; e < 16; e++) acc[e] = make_double2(acc[e].x * (1.0 / H), acc[e].y * (1.0 / H));
        fft_inv2(acc, xa, xc, tw2, tw8, tw9e, t, unit_l);
        uint64_t *dst = a.out + (unit * a.ncomp + c) * N;
#pragma unroll
        for (int m = 0; m < 16; m++) { dst[t + 64 * m] = d2torus(acc[m].x); dst[t + 64 * m + H] = d2torus(-acc[m].y); }
    }
}

// reference slot order [poly][16t + e] -> thread order [poly][e][t], scaled (1/H for the phase-2 keys)
__global__ void k_permute_scale(const cplx *__restrict__ in, cplx *__restrict__ out, size_t polys, double scale) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= polys * H) return;
    const size_t p = i / H;
    const int r = (int)(i % H), e = r / UT, t = r % UT;
    const cplx v = in[p * H + 16 * t + e];
    out[i] = make_double2(v.x * scale, v.y * scale);
}

// reference slot order [poly][16t + e] -> thread order [poly][e][t]
__global__ void k_permute_brk(const cplx *__restrict__ in, cplx *__restrict__ out, size_t polys) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= polys * H) return;
    const size_t p = i / H;
    const int r = (int)(i % H), e = r / UT, t = r % UT;
    out[i] = in[p * H + 16 * t + e];
}

}  // namespace fast

#include "kernels_fast_w.cuh"

struct FastKeys {
    std::vector<cplx *> brk;        // per party, FAST layout
    cplx **d_brk = nullptr;
    cplx *t2 = nullptr, *t8 = nullptr, *t9 = nullptr, *emono = nullptr, *t2w = nullptr;
    bool brk_w = false;             // brk is in the one-warp kernel's thread order [e < 32][t < 32]
    // phase-2 keys in thread order, scaled by 1/H
    std::vector<cplx *> rlk, pubb;
    cplx **d_rlk = nullptr, **d_pubb = nullptr, *crs = nullptr;
    bool built = false;
};

// MKTFHE_FAST_KERNEL = w32 (default: one warp per transform, kernels_fast_w.cuh; KMS_block keeps the tma kernel) | tma | tmem | smem
static inline const std::string &fast_variant() {
    static const std::string v = []() { const char *e = getenv("MKTFHE_FAST_KERNEL"); return std::string(e && *e ? e : "w32"); }();
    return v;
}
// the variants that can hand the RLEV rows to k_phase2 in its thread order
static inline bool fast_variant_tma() { return fast_variant() == "tma" || fast_variant() == "w32"; }
static inline bool fast_use_w(const mktfhe_params &p) { return fast_variant() == "w32" && p.scheme == MKTFHE_KMS; }

static inline bool fast_supported(const mktfhe_params &p) {
    // the 64-bit kernels form the rounding constant 1 << (64 - l*logB - 1): a gadget that fills the whole word (no named set of
    // params.jl does; the tightest is 22 spare bits) runs the STRICT kernels instead
    if (p.l_gsw * p.logB_gsw >= 64 || p.l_lev * p.logB_lev >= 64 || p.l_uni * p.logB_uni >= 64) return false;
    return p.N == 2048 && (p.scheme == MKTFHE_KMS || (p.scheme == MKTFHE_KMS_BLOCK && p.ell == 3));
}

static inline void fast_free(FastKeys &f) {
    for (auto &q : f.brk) if (q) cudaFree(q);
    f.brk.clear();
    if (f.d_brk) cudaFree(f.d_brk);
    if (f.t2) cudaFree(f.t2);
    if (f.t8) cudaFree(f.t8);
    if (f.t9) cudaFree(f.t9);
    if (f.emono) cudaFree(f.emono);
    if (f.t2w) cudaFree(f.t2w);
    for (auto &q : f.rlk) if (q) cudaFree(q);
    for (auto &q : f.pubb) if (q) cudaFree(q);
    f.rlk.clear(); f.pubb.clear();
    if (f.d_rlk) cudaFree(f.d_rlk);
    if (f.d_pubb) cudaFree(f.d_pubb);
    if (f.crs) cudaFree(f.crs);
    f.d_brk = nullptr; f.t2 = f.t8 = f.t9 = f.emono = f.t2w = nullptr; f.brk_w = false; f.d_rlk = f.d_pubb = nullptr; f.crs = nullptr; f.built = false;
}

#define FCK(call)                                                                        \
    do {                                                                                 \
        cudaError_t e_ = (call);                                                         \
        if (e_ != cudaSuccess) { err = std::string(#call) + ": " + cudaGetErrorString(e_); return MKTFHE_ERR_CUDA_; } \
    } while (0)
#define MKTFHE_ERR_CUDA_ (-2)

// Twiddles sqrt(rho(s, i)) = exp(-i*pi*theta(s,i)/2): theta(0,0) = 1/2, theta(s+1, 2i+b) = theta(s,i)/2 + b.
static inline int fast_build(FastKeys &f, const mktfhe_params &p, const std::vector<cplx *> &brk_ref, const std::vector<cplx *> &rlk_ref,
                             const std::vector<cplx *> &pubb_ref, const cplx *crs_ref, cudaStream_t stream, std::string &err) {
    using namespace fast;
    fast_free(f);
    std::vector<__float128> theta(1, (__float128)0.5);
    std::vector<cplx> tw(1024, make_double2(0.0, 0.0)), t2(128), t8(128), t9(256), emono(4096), e16(16), tw1(16, make_double2(0.0, 0.0));
    const __float128 pi = acosq((__float128)-1);
    for (int s = 0; s < 10; s++) {
        std::vector<__float128> nxt(theta.size() * 2);
        for (size_t i = 0; i < theta.size(); i++) {
            const __float128 ang = pi * theta[i] / 2;
            tw[((size_t)1 << s) + i] = make_double2((double)cosq(ang), (double)-sinq(ang));
            nxt[2 * i] = theta[i] / 2; nxt[2 * i + 1] = theta[i] / 2 + 1;
        }
        theta.swap(nxt);
    }
    for (int blk = 0; blk < 16; blk++) {
        t2[blk] = tw[16 + blk]; t2[16 + blk] = tw[32 + 2 * blk];
        t2[32 + blk] = tw[64 + 4 * blk]; t2[48 + blk] = tw[64 + 4 * blk + 2];
        for (int c = 0; c < 4; c++) t2[64 + 16 * c + blk] = tw[128 + 8 * blk + 2 * c];
    }
    for (int t = 0; t < 64; t++) {
        t8[t] = tw[256 + 4 * t]; t8[64 + t] = tw[256 + 4 * t + 2];
        for (int g = 0; g < 4; g++) t9[g * 64 + t] = tw[512 + 2 * (4 * t + g)];
    }
    for (int m = 0; m < 4096; m++) {
        const __float128 ang = pi * m / 2048;
        emono[m] = make_double2((double)(cosq(ang) / H), (double)(-sinq(ang) / H));
    }
    for (int j = 0; j < 16; j++) { const __float128 ang = pi * j / 8; e16[j] = make_double2((double)cosq(ang), (double)-sinq(ang)); }
    for (int i = 1; i < 16; i++) tw1[i] = tw[i];
    {   // one-warp transform (kernels_fast_w.cuh): stages 1..5 in the constant bank, stages 6..10 per thread [16][32]
        // t2w: [16][32] forward pass 2 | [16][32] inverse (decimation in time) stages 5..9 | untwist exp(i*pi*k/2048), k = 0..512
        std::vector<cplx> tw1w(32, make_double2(0.0, 0.0)), e32(32), t2w(512 + 512 + 514, make_double2(0.0, 0.0)), w32i(16);
        auto unit = [&](__float128 num, __float128 den) { const __float128 ang = 2 * pi * num / den; return make_double2((double)cosq(ang), (double)sinq(ang)); };
        for (int j = 0; j < 16; j++) w32i[j] = unit(j, 32);
        for (int t = 0; t < 32; t++) {
            cplx *ti = t2w.data() + 512;
            ti[t] = unit(t, 64);
            ti[32 + t] = unit(t, 128);
            for (int g = 0; g < 2; g++) ti[(2 + g) * 32 + t] = unit(t + 32 * g, 256);
            for (int g = 0; g < 4; g++) ti[(4 + g) * 32 + t] = unit(t + 32 * g, 512);
            for (int g = 0; g < 8; g++) ti[(8 + g) * 32 + t] = unit(t + 32 * g, 1024);
        }
        for (int k = 0; k <= 512; k++) t2w[1024 + k] = unit(k, 4096);
        for (int i = 1; i < 32; i++) tw1w[i] = tw[i];
        for (int j = 0; j < 32; j++) { const __float128 ang = pi * j / 16; e32[j] = make_double2((double)cosq(ang), (double)-sinq(ang)); }
        for (int t = 0; t < 32; t++) {
            t2w[t] = tw[32 + t];
            t2w[32 + t] = tw[64 + 2 * t];
            for (int g = 0; g < 2; g++) t2w[(2 + g) * 32 + t] = tw[128 + 4 * t + 2 * g];
            for (int g = 0; g < 4; g++) t2w[(4 + g) * 32 + t] = tw[256 + 8 * t + 2 * g];
            for (int g = 0; g < 8; g++) t2w[(8 + g) * 32 + t] = tw[512 + 16 * t + 2 * g];
        }
        FCK(cudaMalloc(&f.t2w, sizeof(cplx) * t2w.size()));
        FCK(cudaMemcpy(f.t2w, t2w.data(), sizeof(cplx) * t2w.size(), cudaMemcpyHostToDevice));
        FCK(cudaMemcpyToSymbol(fastw::c_w32i, w32i.data(), sizeof(cplx) * 16));
        FCK(cudaMemcpyToSymbol(fastw::c_tw1w, tw1w.data(), sizeof(cplx) * 32));
        FCK(cudaMemcpyToSymbol(fastw::c_e32, e32.data(), sizeof(cplx) * 32));
        FCK(cudaFuncSetAttribute(fastw::k_phase1_w, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fastw::SMEM_BYTES_W));

    }
    FCK(cudaMalloc(&f.t2, sizeof(cplx) * 128));
    FCK(cudaMalloc(&f.t8, sizeof(cplx) * 128));
    FCK(cudaMalloc(&f.t9, sizeof(cplx) * 256));
    FCK(cudaMalloc(&f.emono, sizeof(cplx) * 4096));
    FCK(cudaMemcpy(f.t2, t2.data(), sizeof(cplx) * 128, cudaMemcpyHostToDevice));
    FCK(cudaMemcpy(f.t8, t8.data(), sizeof(cplx) * 128, cudaMemcpyHostToDevice));
    FCK(cudaMemcpy(f.t9, t9.data(), sizeof(cplx) * 256, cudaMemcpyHostToDevice));
    FCK(cudaMemcpy(f.emono, emono.data(), sizeof(cplx) * 4096, cudaMemcpyHostToDevice));
    FCK(cudaMemcpyToSymbol(c_tw1, tw1.data(), sizeof(cplx) * 16));
    FCK(cudaMemcpyToSymbol(c_e16, e16.data(), sizeof(cplx) * 16));
    const size_t polys = (size_t)p.n * 4 * p.l_gsw;
    f.brk.assign(brk_ref.size(), nullptr);
    for (size_t i = 0; i < brk_ref.size(); i++) {
        FCK(cudaMalloc(&f.brk[i], polys * H * sizeof(cplx)));
        if (fast_use_w(p)) fastw::k_permute_brk_w<<<(unsigned)((polys * H + 255) / 256), 256, 0, stream>>>(brk_ref[i], f.brk[i], polys);
        else k_permute_brk<<<(unsigned)((polys * H + 255) / 256), 256, 0, stream>>>(brk_ref[i], f.brk[i], polys);
        FCK(cudaGetLastError());
    }
    f.brk_w = fast_use_w(p);
    FCK(cudaMalloc(&f.d_brk, sizeof(cplx *) * f.brk.size()));
    FCK(cudaMemcpy(f.d_brk, f.brk.data(), sizeof(cplx *) * f.brk.size(), cudaMemcpyHostToDevice));
    FCK(cudaFuncSetAttribute(k_phase1<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES));
    FCK(cudaFuncSetAttribute(k_phase1<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES));
    FCK(cudaFuncSetAttribute(k_phase1_tm<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES_TM));
    FCK(cudaFuncSetAttribute(k_phase1_tm<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES_TM));
    FCK(cudaFuncSetAttribute(k_phase1_tma<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES_TMA));
    FCK(cudaFuncSetAttribute(k_phase1_tma<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES_TMA));
    FCK(cudaFuncSetAttribute(k_phase2, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES_TM));
    // phase-2 keys: thread order, 1/H folded in (a power of two: exact)
    f.rlk.assign(rlk_ref.size(), nullptr); f.pubb.assign(pubb_ref.size(), nullptr);
    auto permute = [&](const cplx *src, cplx *&dst, size_t npoly) -> int {
        FCK(cudaMalloc(&dst, npoly * H * sizeof(cplx)));
        k_permute_scale<<<(unsigned)((npoly * H + 255) / 256), 256, 0, stream>>>(src, dst, npoly, 1.0 / H);
        FCK(cudaGetLastError());
        return 0;
    };
    for (size_t i = 0; i < rlk_ref.size(); i++) {
        int rc;
        if ((rc = permute(rlk_ref[i], f.rlk[i], (size_t)p.l_uni * 3))) return rc;
        if ((rc = permute(pubb_ref[i], f.pubb[i], (size_t)p.l_uni))) return rc;
    }
    { int rc; if ((rc = permute(crs_ref, f.crs, (size_t)p.l_uni))) return rc; }
    FCK(cudaMalloc(&f.d_rlk, sizeof(cplx *) * f.rlk.size()));
    FCK(cudaMalloc(&f.d_pubb, sizeof(cplx *) * f.pubb.size()));
    FCK(cudaMemcpy(f.d_rlk, f.rlk.data(), sizeof(cplx *) * f.rlk.size(), cudaMemcpyHostToDevice));
    FCK(cudaMemcpy(f.d_pubb, f.pubb.data(), sizeof(cplx *) * f.pubb.size(), cudaMemcpyHostToDevice));
    f.built = true;
    return 0;
}

static inline int fast_launch(FastKeys &f, const mktfhe_params &p, fast::Args a, cudaStream_t stream, int *launches, std::string &err) {
    using namespace fast;
    if (!f.built) { err = "FAST keys not built"; return -3; }
    a.brk = f.d_brk; a.tb = Tables{f.t2, f.t8, f.t9, f.emono, f.t2w};
    a.n = p.n; a.k = p.k; a.l = p.l_gsw; a.logB = p.logB_gsw; a.l_lev = p.l_lev; a.logB_lev = p.logB_lev;
    a.R = 1 + (p.k - 1) * p.l_lev; a.lwe_words = (int)mktfhe_lwe_words(&p);
    const unsigned grid = (unsigned)((a.units + U - 1) / U);
    a.d = p.d;
    const std::string variant = fast_use_w(p) ? "w32" : (fast_variant() == "w32" ? "tma" : fast_variant());      // w32 | tma | tmem | smem
    if (variant == "w32") {
        if (!f.brk_w) { err = "FAST keys are not in the one-warp kernel's order"; return -3; }
        size_t ctas;
        if (a.step_mode) ctas = (a.units + fastw::WU - 1) / fastw::WU;
        else {
            const size_t gates = a.units / a.R;
            ctas = (gates + fastw::WU - 1) / fastw::WU + (size_t)(p.k - 1) * ((gates * p.l_lev + fastw::WU - 1) / fastw::WU);
        }
        fastw::k_phase1_w<<<(unsigned)ctas, fastw::CTA_W, fastw::SMEM_BYTES_W, stream>>>(a);
    } else if (variant == "tma") {
        // CTAs are grouped by party (one key stream per CTA): ceil(gates/U) for party 0 plus ceil(gates*l_lev/U) per other party
        size_t ctas;
        if (a.step_mode) ctas = (a.units + U - 1) / U;
        else {
            const size_t gates = a.units / a.R;
            ctas = (gates + U - 1) / U + (size_t)(p.k - 1) * ((gates * p.l_lev + U - 1) / U);
        }
        if (p.scheme == MKTFHE_KMS_BLOCK) k_phase1_tma<3><<<(unsigned)ctas, CTA_TMA, SMEM_BYTES_TMA, stream>>>(a);
        else k_phase1_tma<1><<<(unsigned)ctas, CTA_TMA, SMEM_BYTES_TMA, stream>>>(a);
    } else if (variant == "tmem") {
        if (p.scheme == MKTFHE_KMS_BLOCK) k_phase1_tm<3><<<grid, CTA, SMEM_BYTES_TM, stream>>>(a);
        else k_phase1_tm<1><<<grid, CTA, SMEM_BYTES_TM, stream>>>(a);
    } else {
        if (p.scheme == MKTFHE_KMS_BLOCK) k_phase1<3><<<grid, CTA, SMEM_BYTES, stream>>>(a);
        else k_phase1<1><<<grid, CTA, SMEM_BYTES, stream>>>(a);
    }
    if (launches) (*launches)++;
    {
        cudaError_t e_ = cudaGetLastError();
        if (e_ != cudaSuccess) {
            cudaFuncAttributes fa{};
            cudaFuncGetAttributes(&fa, k_phase1_tma<1>);
            err = std::string("phase-1 launch (") + variant + "): " + cudaGetErrorString(e_) + " [regs " + std::to_string(fa.numRegs) +
                  ", static smem " + std::to_string(fa.sharedSizeBytes) + ", max dyn smem " + std::to_string(fa.maxDynamicSharedSizeBytes) +
                  ", max threads " + std::to_string(fa.maxThreadsPerBlock) + "]";
            return MKTFHE_ERR_CUDA_;
        }
    }
    return 0;
}

// lev_fast: write the RLEV rows for k_phase2 (thread order, scaled by 1/H); only the TMA kernel implements it
static inline int fast_phase1(FastKeys &f, const mktfhe_params &p, const uint32_t *tilde, cplx *lev, size_t gates,
                              cudaStream_t stream, int *launches, std::string &err, bool lev_fast = false) {
    fast::Args a{};
    a.tilde = tilde; a.lev_out = lev; a.step_mode = 0; a.lev_fast = lev_fast && fast_variant_tma();
    a.units = gates * (size_t)(1 + (p.k - 1) * p.l_lev);
    return fast_launch(f, p, a, stream, launches, err);
}

static inline int fast_phase2(FastKeys &f, const mktfhe_params &p, const uint32_t *tilde, const cplx *lev, uint64_t *acc, cplx *tx, cplx *ty,
                              size_t gates, cudaStream_t stream, int *launches, std::string &err) {
    using namespace fast;
    if (!f.built) { err = "FAST keys not built"; return -3; }
    P2Args a{};
    a.tilde = tilde; a.lev = lev; a.rlk = f.d_rlk; a.pubb = f.d_pubb; a.crs = f.crs; a.tb = Tables{f.t2, f.t8, f.t9, f.emono, f.t2w};
    a.acc = acc; a.tx = tx; a.ty = ty;
    a.k = p.k; a.l_lev = p.l_lev; a.logB_lev = p.logB_lev; a.l_uni = p.l_uni; a.logB_uni = p.logB_uni;
    a.R = 1 + (p.k - 1) * p.l_lev; a.lwe_words = (int)mktfhe_lwe_words(&p); a.gates = gates;
    k_phase2<<<(unsigned)((gates + U - 1) / U), CTA, SMEM_BYTES_TM, stream>>>(a);
    if (launches) (*launches)++;
    FCK(cudaGetLastError());
    return 0;
}

static inline int fast_gadget_product(FastKeys &f, const uint64_t *polys, const cplx *keys, uint64_t *out, int l, int logB, int ncomp,
                                      size_t batch, cudaStream_t stream, int *launches, std::string &err) {
    using namespace fast;
    if (!f.built) { err = "FAST keys not built"; return -3; }
    GpArgs a{};
    a.polys = polys; a.keys = keys; a.out = out; a.tb = Tables{f.t2, f.t8, f.t9, f.emono, f.t2w};
    a.l = l; a.logB = logB; a.ncomp = ncomp; a.units = batch;
    FCK(cudaFuncSetAttribute(k_gadget_product, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES_TM));
    k_gadget_product<<<(unsigned)((batch + U - 1) / U), CTA, SMEM_BYTES_TM, stream>>>(a);
    if (launches) (*launches)++;
    FCK(cudaGetLastError());
    return 0;
}

static inline int fast_cmux_step(FastKeys &f, const mktfhe_params &p, int party, int idx, const uint32_t *atilde, void *rows,
                                 size_t batch, cudaStream_t stream, int *launches, std::string &err) {
    fast::Args a{};
    a.tilde = atilde; a.acc_io = (uint64_t *)rows; a.step_mode = 1; a.step_party = party; a.step_idx = idx;
    a.units = batch;
    return fast_launch(f, p, a, stream, launches, err);
}
