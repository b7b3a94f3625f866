// kernels_fast32.cuh -- FAST-mode blind rotation for the N = 1024, UInt32 single-key schemes:
//   CGGI  blindrotate!  /root/reference/src/tfhe/bootstrapping.jl:32-76
//   LMSS  blindrotate!  /root/reference/src/tfhe/bootstrapping.jl:114-165
// Same construction as kernels_fast.cuh (twist-free product-tree transform, 6-instruction butterflies, fused gadget
// decomposition, monomial rebuilt per slot, RGSW accumulators in TMEM), sized for H = 512:
//   unit = one gate = 64 threads x 8 points, passes of 3 + 3 + 3 stages, two double-buffered exchanges;
//   8 units (16 warps) per CTA, 1 CTA per SM; the RLWE accumulator (8 KiB) stays in shared memory,
//   the 2 x 8 complex RGSW accumulators per thread in TMEM (32 columns per thread).
#pragma once
#include "kernels_fast.cuh"

namespace fast32 {

using fast::bf; using fast::bf_mi; using fast::bi; using fast::bi_mi; using fast::c_e16;
using fast::tm_ld16; using fast::tm_st16; using fast::tm_wait_ld; using fast::tm_wait_st;

constexpr int H = 512, N = 1024, UT = 64, U = 8, CTA = UT * U;
constexpr int XB_LEN = H + 64;                       // exchange 2 layout: n + (n >> 3)

__constant__ double2 c_tw1[8];                       // TW[1..7]: stages 1..3

struct Tables {
    const cplx *t2;        // [4][8]   per 64-point block: w4, w5, w6a, w6b      (TW[8+blk], TW[16+2blk], TW[32+4blk], TW[32+4blk+2])
    const cplx *t3;        // [4][64]  per thread: TW[64+t], TW[128+2t], TW[256+4t], TW[256+4t+2]
    const cplx *emono;     // [2048]   exp(-i*pi*m/1024) / H
};

// stages 1..3: element m of thread t is point t + 64m
__device__ __forceinline__ void pass1_fwd(cplx (&x)[8]) {
#pragma unroll
    for (int m = 0; m < 4; m++) bf(x[m], x[m + 4], c_tw1[1]);
#pragma unroll
    for (int m = 0; m < 8; m++) if (!(m & 2)) { if (m & 4) bf_mi(x[m], x[m + 2], c_tw1[2]); else bf(x[m], x[m + 2], c_tw1[2]); }
#pragma unroll
    for (int m = 0; m < 8; m += 2) { if (m & 2) bf_mi(x[m], x[m + 1], c_tw1[4 + (m >> 2) * 2]); else bf(x[m], x[m + 1], c_tw1[4 + (m >> 2) * 2]); }
}
__device__ __forceinline__ void pass1_inv(cplx (&x)[8]) {
#pragma unroll
    for (int m = 0; m < 8; m += 2) { if (m & 2) bi_mi(x[m], x[m + 1], c_tw1[4 + (m >> 2) * 2]); else bi(x[m], x[m + 1], c_tw1[4 + (m >> 2) * 2]); }
#pragma unroll
    for (int m = 0; m < 8; m++) if (!(m & 2)) { if (m & 4) bi_mi(x[m], x[m + 2], c_tw1[2]); else bi(x[m], x[m + 2], c_tw1[2]); }
#pragma unroll
    for (int m = 0; m < 4; m++) bi(x[m], x[m + 4], c_tw1[1]);
}
// stages 4..6 inside 64-point block blk: element q is point 64*blk + o + 8q
__device__ __forceinline__ void pass2_fwd(cplx (&x)[8], const cplx *__restrict__ tw, int blk) {
    const cplx w4 = tw[blk], w5 = tw[8 + blk], w6a = tw[16 + blk], w6b = tw[24 + blk];
#pragma unroll
    for (int q = 0; q < 4; q++) bf(x[q], x[q + 4], w4);
#pragma unroll
    for (int q = 0; q < 8; q++) if (!(q & 2)) { if (q & 4) bf_mi(x[q], x[q + 2], w5); else bf(x[q], x[q + 2], w5); }
#pragma unroll
    for (int q = 0; q < 8; q += 2) { const cplx w = (q & 4) ? w6b : w6a; if (q & 2) bf_mi(x[q], x[q + 1], w); else bf(x[q], x[q + 1], w); }
}
__device__ __forceinline__ void pass2_inv(cplx (&x)[8], const cplx *__restrict__ tw, int blk) {
    const cplx w4 = tw[blk], w5 = tw[8 + blk], w6a = tw[16 + blk], w6b = tw[24 + blk];
#pragma unroll
    for (int q = 0; q < 8; q += 2) { const cplx w = (q & 4) ? w6b : w6a; if (q & 2) bi_mi(x[q], x[q + 1], w); else bi(x[q], x[q + 1], w); }
#pragma unroll
    for (int q = 0; q < 8; q++) if (!(q & 2)) { if (q & 4) bi_mi(x[q], x[q + 2], w5); else bi(x[q], x[q + 2], w5); }
#pragma unroll
    for (int q = 0; q < 4; q++) bi(x[q], x[q + 4], w4);
}
// stages 7..9 on the 8 contiguous points 8t .. 8t+7: nodes t (depth 6), 2t+{0,1}, 4t+{0..3}
__device__ __forceinline__ void pass3_fwd(cplx (&x)[8], const cplx *__restrict__ tw, int t) {
    const cplx w7 = tw[t], w8 = tw[UT + t], w9a = tw[2 * UT + t], w9b = tw[3 * UT + t];
#pragma unroll
    for (int e = 0; e < 4; e++) bf(x[e], x[e + 4], w7);
#pragma unroll
    for (int e = 0; e < 8; e++) if (!(e & 2)) { if (e & 4) bf_mi(x[e], x[e + 2], w8); else bf(x[e], x[e + 2], w8); }
#pragma unroll
    for (int e = 0; e < 8; e += 2) { const cplx w = (e & 4) ? w9b : w9a; if (e & 2) bf_mi(x[e], x[e + 1], w); else bf(x[e], x[e + 1], w); }
}
__device__ __forceinline__ void pass3_inv(cplx (&x)[8], const cplx *__restrict__ tw, int t) {
    const cplx w7 = tw[t], w8 = tw[UT + t], w9a = tw[2 * UT + t], w9b = tw[3 * UT + t];
#pragma unroll
    for (int e = 0; e < 8; e += 2) { const cplx w = (e & 4) ? w9b : w9a; if (e & 2) bi_mi(x[e], x[e + 1], w); else bi(x[e], x[e + 1], w); }
#pragma unroll
    for (int e = 0; e < 8; e++) if (!(e & 2)) { if (e & 4) bi_mi(x[e], x[e + 2], w8); else bi(x[e], x[e + 2], w8); }
#pragma unroll
    for (int e = 0; e < 4; e++) bi(x[e], x[e + 4], w7);
}

__device__ __forceinline__ void unit_bar(int unit) { asm volatile("bar.sync %0, %1;" ::"r"(unit + 1), "r"(UT) : "memory"); }

// Exchange 1 is conflict-free as is (8 consecutive threads touch 8 consecutive elements); exchange 2 is skewed by n >> 3.
// The two buffers strictly alternate over the whole kernel (xa then xc, forward and inverse alike).
template <class F>
__device__ __forceinline__ void fft_fwd(cplx (&x)[8], cplx *xa, cplx *xc, const cplx *tw2, const cplx *tw3, int t, int unit, F mid) {
    pass1_fwd(x);
#pragma unroll
    for (int m = 0; m < 8; m++) xa[t + 64 * m] = x[m];
    unit_bar(unit);
    const int blk = t >> 3, o = t & 7;
#pragma unroll
    for (int q = 0; q < 8; q++) x[q] = xa[64 * blk + o + 8 * q];
    mid();
    pass2_fwd(x, tw2, blk);
#pragma unroll
    for (int q = 0; q < 8; q++) xc[72 * blk + o + 9 * q] = x[q];            // n + (n >> 3), n = 64blk + o + 8q
    unit_bar(unit);
#pragma unroll
    for (int e = 0; e < 8; e++) x[e] = xc[9 * t + e];                        // n = 8t + e
    pass3_fwd(x, tw3, t);
}
__device__ __forceinline__ void fft_inv(cplx (&x)[8], cplx *xa, cplx *xc, const cplx *tw2, const cplx *tw3, int t, int unit) {
    pass3_inv(x, tw3, t);
#pragma unroll
    for (int e = 0; e < 8; e++) xa[9 * t + e] = x[e];
    unit_bar(unit);
    const int blk = t >> 3, o = t & 7;
#pragma unroll
    for (int q = 0; q < 8; q++) x[q] = xa[72 * blk + o + 9 * q];
    pass2_inv(x, tw2, blk);
#pragma unroll
    for (int q = 0; q < 8; q++) xc[64 * blk + o + 8 * q] = x[q];
    unit_bar(unit);
#pragma unroll
    for (int m = 0; m < 8; m++) x[m] = xc[t + 64 * m];
    pass1_inv(x);
}

// x mod 2^32 as uint32, rounding toward -inf like `native` (arithmetic.jl:1-4)
__device__ __forceinline__ uint32_t d2torus32(double x) {
    const double q = fma(x, 2.3283064365386963e-10, 6755399441055744.0) - 6755399441055744.0;     // rint(x / 2^32)
    const double y = fma(-4294967296.0, q, x);                                                   // in [-2^31, 2^31]
    return (uint32_t)__double2ll_rd(y);
}

struct Args {
    const uint32_t *tilde;        // [B][lwe_words] / step modes: [units][ELL] rotations
    const cplx *brk;              // FAST layout [idx][dg][comp][e < 8][t < 64]
    Tables tb;
    uint32_t *acc_io;             // out [B][2][N] (step modes: in/out)
    int step_mode, step_idx;
    int n, d, l, logB, lwe_words;
    size_t units;
};

constexpr size_t SMEM_UNIT = (size_t)2 * XB_LEN * 16;                           // two exchange buffers (accumulators live in TMEM)
constexpr size_t SMEM_BYTES = U * SMEM_UNIT + (size_t)(32 + 256) * 16 + 16;

template <int ELL>
__global__ void __launch_bounds__(CTA, 1) k_rgsw_tm(const Args a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int tid = threadIdx.x, warp = tid >> 5, unit_l = tid / UT, t = tid % UT;
    cplx *tw2 = reinterpret_cast<cplx *>(smem_raw + U * SMEM_UNIT), *tw3 = tw2 + 32;
    uint32_t *tm_base_s = reinterpret_cast<uint32_t *>(tw3 + 256);
    for (int i = tid; i < 256; i += CTA) { if (i < 32) tw2[i] = a.tb.t2[i]; tw3[i] = a.tb.t3[i]; }
    cplx *xa = reinterpret_cast<cplx *>(smem_raw + unit_l * SMEM_UNIT), *xc = xa + XB_LEN;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" :: "r"((uint32_t)__cvta_generic_to_shared(tm_base_s)));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    // warp w -> lanes 32*(w%4).., columns 96*(w/4)..; per thread: tacc.b = columns [0,32), tacc.a = [32,64),
    // acc.b = [64,80), acc.a = [80,96): word m of a row is coefficient t + 64m (only this thread ever touches it)
    const uint32_t tm = *tm_base_s + ((uint32_t)(32 * (warp & 3)) << 16) + 96u * (uint32_t)(warp >> 2);
    constexpr uint32_t TM_ACCB = 64, TM_ACCA = 80;

    const size_t unit = (size_t)blockIdx.x * U + unit_l;
    if (unit < a.units) {
        const int gate = (int)unit;
        {
            uint32_t vb[16], va[16];
            if (!a.step_mode) {                    // test vector: bootstrapping.jl:11-23
                const uint32_t tb = a.tilde[(size_t)gate * a.lwe_words];
                const uint32_t e8 = 1u << 29;
#pragma unroll
                for (int m = 0; m < 16; m++) {
                    const uint32_t i1 = (uint32_t)(t + 64 * m) + 1;      // 1-based coefficient index
                    vb[m] = tb <= (uint32_t)N ? (i1 <= tb ? e8 : 0u - e8) : (i1 <= tb - (uint32_t)N ? 0u - e8 : e8);
                    va[m] = 0u;
                }
            } else {
                const uint32_t *src = a.acc_io + unit * 2 * N;
#pragma unroll
                for (int m = 0; m < 16; m++) { vb[m] = src[t + 64 * m]; va[m] = src[N + t + 64 * m]; }
            }
            fast::tm_st16(tm + TM_ACCB, vb);
            fast::tm_st16(tm + TM_ACCA, va);
            tm_wait_st();
        }

        const int l = a.l, logB = a.logB;
        const int bit = 32 - l * logB;
        uint32_t cadd = bit > 0 ? 1u << (bit - 1) : 0u;                 // divbits rounding (arithmetic.jl:23-27)
        for (int j = 0; j < l; j++) cadd += 1u << (bit + j * logB + logB - 1);   // + B/2 at every digit position (gsw.jl:86-96)
        const uint32_t mask = (1u << logB) - 1;
        const double dbias = 4503599627370496.0 + (double)(1 << (logB - 1));
        const size_t per_idx = (size_t)4 * l * H;
        const uint32_t *at_src = a.step_mode ? a.tilde + unit * ELL : a.tilde + (size_t)gate * a.lwe_words + 1;
        const int nsteps = a.step_mode ? 1 : (ELL == 1 ? a.n : a.d);
        const int brv6t = (int)(__brev((unsigned)t) >> 26);

        for (int step = 0; step < nsteps; step++) {
            uint32_t atv[ELL];
            bool any = false;
#pragma unroll
            for (int b = 0; b < ELL; b++) { atv[b] = at_src[(a.step_mode ? 0 : step * ELL) + b]; any |= atv[b] > 0; }
            if (!any) continue;                                           // :48 / whole-block no-op
            const int idx = (a.step_mode ? a.step_idx : step) * ELL;
            const cplx *kidx = a.brk + (size_t)idx * per_idx + t;
            // slot n = 8t + e evaluates at exp(-i*pi*(4*brv9(n)+1)/N), brv9(n) = 64*brv3(e) + brv6(t)
            cplx m1v[ELL];
#pragma unroll
            for (int b = 0; b < ELL; b++) m1v[b] = __ldg(&a.tb.emono[((4 * brv6t + 1) * atv[b]) & 2047]);

            for (int dg = 0; dg < 2 * l; dg++) {
                const int sh = bit + (l - 1 - (dg < l ? dg : dg - l)) * logB;
                cplx x[8];
                {
                    uint32_t v[16];
                    fast::tm_ld16(tm + (dg < l ? TM_ACCB : TM_ACCA), v);
                    fast::tm_wait_ld();
                    fast::tm_pin16(v);
#pragma unroll
                    for (int m = 0; m < 8; m++) {
                        const uint32_t f0 = ((v[m] + cadd) >> sh) & mask, f1 = ((v[m + 8] + cadd) >> sh) & mask;
                        x[m] = make_double2(__hiloint2double(0x43300000, (int)f0) - dbias, dbias - __hiloint2double(0x43300000, (int)f1));
                    }
                }
                const cplx *kb = kidx + (size_t)(dg * 2) * H, *ka = kb + H;
                cplx kcb[8], kca[8];
                if (ELL == 1) {
                    fft_fwd(x, xa, xc, tw2, tw3, t, unit_l, [&]() {
#pragma unroll
                        for (int e = 0; e < 8; e++) { kcb[e] = __ldg(kb + e * UT); kca[e] = __ldg(ka + e * UT); }
                    });
                } else {
                    // block (LMSS): fold the monomials of the block's key bits into the keys,
                    //   sum_bit mono_bit * (sum_dg D_dg * K_bit,dg) = sum_dg D_dg * (sum_bit mono_bit * K_bit,dg)
                    fft_fwd(x, xa, xc, tw2, tw3, t, unit_l, []() {});
#pragma unroll
                    for (int e = 0; e < 8; e++) {
                        const int b3 = ((e & 1) << 2) | (e & 2) | ((e & 4) >> 2);
                        kcb[e] = kca[e] = make_double2(0.0, 0.0);
#pragma unroll
                        for (int b = 0; b < ELL; b++) {
                            if (atv[b] == 0) continue;
                            cplx mo = cmul_f(m1v[b], c_e16[((atv[b] * b3) & 7) * 2]);
                            mo.x -= 1.0 / H;
                            kcb[e] = cmac_f(kcb[e], mo, __ldg(kb + b * per_idx + e * UT));
                            kca[e] = cmac_f(kca[e], mo, __ldg(ka + b * per_idx + e * UT));
                        }
                    }
                }
#pragma unroll
                for (int c = 0; c < 2; c++) {
                    cplx zb[4], za[4];
                    if (dg == 0) {
#pragma unroll
                        for (int i = 0; i < 4; i++) { zb[i] = cmul_f(x[4 * c + i], kcb[4 * c + i]); za[i] = cmul_f(x[4 * c + i], kca[4 * c + i]); }
                    } else {
                        fast::tm_ld_c4(tm + 16 * c, zb);
                        fast::tm_ld_c4(tm + 32 + 16 * c, za);
#pragma unroll
                        for (int i = 0; i < 4; i++) { zb[i] = cmac_f(zb[i], x[4 * c + i], kcb[4 * c + i]); za[i] = cmac_f(za[i], x[4 * c + i], kca[4 * c + i]); }
                    }
                    fast::tm_st_c4(tm + 16 * c, zb);
                    fast::tm_st_c4(tm + 32 + 16 * c, za);
                }
                tm_wait_st();
            }
            // both outputs: (x (X^a - 1)/H for ELL == 1) -> inverse transform -> round -> acc +=
#pragma unroll 1
            for (int pz = 0; pz < 2; pz++) {
                cplx y[8];
#pragma unroll
                for (int c = 0; c < 2; c++) {
                    cplx z[4];
                    fast::tm_ld_c4(tm + 32 * pz + 16 * c, z);
#pragma unroll
                    for (int i = 0; i < 4; i++) y[4 * c + i] = z[i];
                }
                if (ELL == 1) {
#pragma unroll
                    for (int e = 0; e < 8; e++) {
                        const int b3 = ((e & 1) << 2) | (e & 2) | ((e & 4) >> 2);
                        cplx mo = cmul_f(m1v[0], c_e16[((atv[0] * b3) & 7) * 2]);      // 8th roots = even 16th roots
                        mo.x -= 1.0 / H;
                        y[e] = cmul_f(mo, y[e]);
                    }
                }
                fft_inv(y, xa, xc, tw2, tw3, t, unit_l);
                {
                    uint32_t v[16];
                    fast::tm_ld16(tm + (pz == 0 ? TM_ACCB : TM_ACCA), v);
                    fast::tm_wait_ld();
                    fast::tm_pin16(v);
#pragma unroll
                    for (int m = 0; m < 8; m++) { v[m] += d2torus32(y[m].x); v[m + 8] += d2torus32(-y[m].y); }
                    fast::tm_st16(tm + (pz == 0 ? TM_ACCB : TM_ACCA), v);
                    tm_wait_st();
                }
            }
        }
        uint32_t *out = a.acc_io + unit * 2 * N;
        {
            uint32_t vb[16], va[16];
            fast::tm_ld16(tm + TM_ACCB, vb);
            fast::tm_ld16(tm + TM_ACCA, va);
            fast::tm_wait_ld();
            fast::tm_pin16(vb); fast::tm_pin16(va);
#pragma unroll
            for (int m = 0; m < 16; m++) { out[t + 64 * m] = vb[m]; out[N + t + 64 * m] = va[m]; }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" :: "r"(*tm_base_s));
}

// setmaxnreg only redistributes the registers the CTA was launched with: 16 x 112 + 4 x 24 <= 20 x 96
static_assert(16 * 112 + 4 * 24 <= 20 * 96, "setmaxnreg over-subscription");
// ---- the same kernel with the keys delivered by TMA ----------------------------------------------------------------
// All gates of a launch use the same key (one party) in the same order, so one key stream serves the eight units of a CTA:
// a fifth warpgroup issues `cp.async.bulk` copies into a 64 KiB shared-memory ring (the space the RLWE accumulators left
// when they moved to TMEM) guarded by full/empty mbarriers; the consumers read their key values four at a time right before
// the multiply-accumulate.  Key-load latency (48 dependent L2 loads per digit in the LMSS fold) leaves the critical path,
// and the key values no longer occupy 64 registers.  Launched with 20 warps at <= 96 registers and re-split with
// `setmaxnreg` (producer 24, consumers 112).
constexpr int RING_BYTES32 = 64 * 1024;
template <int ELL> struct TileCfg32 { static constexpr int TILE = ELL == 1 ? H : H / 2, RING = RING_BYTES32 / (TILE * 16); };
constexpr int CTA_TMA32 = CTA + 128;
constexpr size_t SMEM_BYTES_TMA32 = U * SMEM_UNIT + (size_t)(32 + 256) * 16 + (size_t)RING_BYTES32 + 512;

template <int ELL>
__global__ void __launch_bounds__(CTA_TMA32, 1) k_rgsw_tma(const Args a) {
    using fast::mb_init; using fast::mb_expect_tx; using fast::mb_arrive; using fast::mb_wait; using fast::mb_wait_suspend; using fast::bulk_g2s;
    constexpr int TILE = TileCfg32<ELL>::TILE, RING = TileCfg32<ELL>::RING, PARTS = H / TILE, EP = 8 / PARTS;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int tid = threadIdx.x, warp = tid >> 5, unit_l = tid / UT, t = tid % UT;
    cplx *tw2 = reinterpret_cast<cplx *>(smem_raw + U * SMEM_UNIT), *tw3 = tw2 + 32;
    cplx *ring = tw3 + 256;
    uint64_t *full = reinterpret_cast<uint64_t *>(ring + (size_t)RING * TILE), *empty = full + RING;
    uint32_t *tm_base_s = reinterpret_cast<uint32_t *>(empty + RING);
    for (int i = tid; i < 256; i += CTA_TMA32) { if (i < 32) tw2[i] = a.tb.t2[i]; tw3[i] = a.tb.t3[i]; }
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

    const int l = a.l, logB = a.logB;
    const size_t per_idx = (size_t)4 * l * H;
    const int nsteps = a.step_mode ? 1 : (ELL == 1 ? a.n : a.d);
    const uint32_t tiles_per_step = (uint32_t)(2 * l * PARTS * ELL * 2);
    const uint32_t ntiles = (uint32_t)nsteps * tiles_per_step;

    if (warp >= U * 2) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 24;");
        if (tid == CTA) {
            // tile n = ((((step * 2l + dg) * PARTS + part) * ELL + b) * 2 + comp)
            for (uint32_t n = 0; n < ntiles; n++) {
                const int slot = n % RING;
                if (n >= RING) mb_wait_suspend(&empty[slot], ((n / RING) - 1) & 1);
                const uint32_t comp = n & 1, b = (n >> 1) % ELL, part = ((n >> 1) / ELL) % PARTS;
                const uint32_t dg = ((n >> 1) / ELL / PARTS) % (2 * l), step = (n >> 1) / ELL / PARTS / (2 * l);
                const int idx = (a.step_mode ? a.step_idx : (int)step) * ELL + (int)b;
                mb_expect_tx(&full[slot], TILE * 16);
                bulk_g2s(ring + (size_t)slot * TILE, a.brk + (size_t)idx * per_idx + (size_t)(dg * 2 + comp) * H + (size_t)part * TILE, TILE * 16, &full[slot]);
            }
        }
    } else {
        asm volatile("setmaxnreg.inc.sync.aligned.u32 112;");   // 16 * 112 + 4 * 24 <= 20 * 96: the re-split only redistributes the CTA's own launch allocation
        cplx *xa = reinterpret_cast<cplx *>(smem_raw + unit_l * SMEM_UNIT), *xc = xa + XB_LEN;
        // warp w -> lanes 32*(w%4).., columns 96*(w/4)..; per thread: tacc.b [0,32), tacc.a [32,64), acc.b [64,80), acc.a [80,96)
        const uint32_t tm = *tm_base_s + ((uint32_t)(32 * (warp & 3)) << 16) + 96u * (uint32_t)(warp >> 2);
        constexpr uint32_t TM_ACCB = 64, TM_ACCA = 80;
        const size_t unit = (size_t)blockIdx.x * U + unit_l;
        const bool live = unit < a.units;
        const size_t gate = live ? unit : 0;
        uint32_t tile_n = 0;
        auto tile_wait = [&]() -> const cplx * {
            const int slot = tile_n % RING;
            mb_wait(&full[slot], (tile_n / RING) & 1);
            return ring + (size_t)slot * TILE + t;
        };
        auto tile_done = [&]() { mb_arrive(&empty[tile_n % RING]); tile_n++; };

        if (live) {
            uint32_t vb[16], va[16];
            if (!a.step_mode) {                    // test vector: bootstrapping.jl:11-23
                const uint32_t tb = a.tilde[gate * a.lwe_words];
                const uint32_t e8 = 1u << 29;
#pragma unroll
                for (int m = 0; m < 16; m++) {
                    const uint32_t i1 = (uint32_t)(t + 64 * m) + 1;      // 1-based coefficient index
                    vb[m] = tb <= (uint32_t)N ? (i1 <= tb ? e8 : 0u - e8) : (i1 <= tb - (uint32_t)N ? 0u - e8 : e8);
                    va[m] = 0u;
                }
            } else {
                const uint32_t *src = a.acc_io + unit * 2 * N;
#pragma unroll
                for (int m = 0; m < 16; m++) { vb[m] = src[t + 64 * m]; va[m] = src[N + t + 64 * m]; }
            }
            fast::tm_st16(tm + TM_ACCB, vb);
            fast::tm_st16(tm + TM_ACCA, va);
            tm_wait_st();
        }
        const int bit = 32 - l * logB;
        uint32_t cadd = bit > 0 ? 1u << (bit - 1) : 0u;                 // divbits rounding (arithmetic.jl:23-27)
        for (int j = 0; j < l; j++) cadd += 1u << (bit + j * logB + logB - 1);   // + B/2 at every digit position (gsw.jl:86-96)
        const uint32_t mask = (1u << logB) - 1;
        const double dbias = 4503599627370496.0 + (double)(1 << (logB - 1));
        const uint32_t *at_src = a.step_mode ? a.tilde + gate * ELL : a.tilde + gate * a.lwe_words + 1;
        const int brv6t = (int)(__brev((unsigned)t) >> 26);

        for (int step = 0; step < nsteps; step++) {
            uint32_t atv[ELL];
            bool any = false;
#pragma unroll
            for (int b = 0; b < ELL; b++) { atv[b] = live ? at_src[(a.step_mode ? 0 : step * ELL) + b] : 0u; any |= atv[b] > 0; }
            if (!any) {                                                   // :48 / whole-block no-op / dead unit: keep the ring moving
                for (uint32_t i = 0; i < tiles_per_step; i++) { tile_wait(); tile_done(); }
                continue;
            }
            // slot n = 8t + e evaluates at exp(-i*pi*(4*brv9(n)+1)/N), brv9(n) = 64*brv3(e) + brv6(t)
            cplx m1v[ELL];
#pragma unroll
            for (int b = 0; b < ELL; b++) m1v[b] = __ldg(&a.tb.emono[((4 * brv6t + 1) * atv[b]) & 2047]);

            for (int dg = 0; dg < 2 * l; dg++) {
                const int sh = bit + (l - 1 - (dg < l ? dg : dg - l)) * logB;
                cplx x[8];
                {
                    uint32_t v[16];
                    fast::tm_ld16(tm + (dg < l ? TM_ACCB : TM_ACCA), v);
                    tm_wait_ld();
                    fast::tm_pin16(v);
#pragma unroll
                    for (int m = 0; m < 8; m++) {
                        const uint32_t f0 = ((v[m] + cadd) >> sh) & mask, f1 = ((v[m + 8] + cadd) >> sh) & mask;
                        x[m] = make_double2(__hiloint2double(0x43300000, (int)f0) - dbias, dbias - __hiloint2double(0x43300000, (int)f1));
                    }
                }
                fft_fwd(x, xa, xc, tw2, tw3, t, unit_l, []() {});
#pragma unroll
                for (int part = 0; part < PARTS; part++) {
                    cplx kcb[EP], kca[EP];
                    if (ELL == 1) {
                        const cplx *kb = tile_wait();
                        const int slot_b = tile_n % RING;
                        tile_n++;                                   // hold the .b tile while the .a tile is awaited
                        const cplx *ka = tile_wait();
#pragma unroll
                        for (int e = 0; e < EP; e++) { kcb[e] = kb[e * UT]; kca[e] = ka[e * UT]; }
                        mb_arrive(&empty[slot_b]);
                        tile_done();
                    } else {
                        // block (LMSS): fold the monomials of the block's key bits into the keys,
                        //   sum_bit mono_bit * (sum_dg D_dg * K_bit,dg) = sum_dg D_dg * (sum_bit mono_bit * K_bit,dg)
#pragma unroll
                        for (int e = 0; e < EP; e++) kcb[e] = kca[e] = make_double2(0.0, 0.0);
#pragma unroll
                        for (int b = 0; b < ELL; b++) {
                            const cplx *kb = tile_wait();
                            const int slot_b = tile_n % RING;
                            tile_n++;
                            const cplx *ka = tile_wait();
                            if (atv[b] != 0) {
#pragma unroll
                                for (int eh = 0; eh < EP; eh++) {
                                    const int e = EP * part + eh;
                                    const int b3 = ((e & 1) << 2) | (e & 2) | ((e & 4) >> 2);
                                    cplx mo = cmul_f(m1v[b], c_e16[((atv[b] * b3) & 7) * 2]);
                                    mo.x -= 1.0 / H;
                                    kcb[eh] = cmac_f(kcb[eh], mo, kb[eh * UT]);
                                    kca[eh] = cmac_f(kca[eh], mo, ka[eh * UT]);
                                }
                            }
                            mb_arrive(&empty[slot_b]);
                            tile_done();
                        }
                    }
#pragma unroll
                    for (int c2 = 0; c2 < EP / 4; c2++) {
                        const int c = (EP / 4) * part + c2;
                        cplx zb[4], za[4];
                        if (dg == 0) {
#pragma unroll
                            for (int i = 0; i < 4; i++) { zb[i] = cmul_f(x[4 * c + i], kcb[4 * c2 + i]); za[i] = cmul_f(x[4 * c + i], kca[4 * c2 + i]); }
                        } else {
                            uint32_t rb[16], ra[16];
                            fast::tm_ld16(tm + 16 * c, rb);
                            fast::tm_ld16(tm + 32 + 16 * c, ra);
                            tm_wait_ld();
                            fast::tm_pin16(rb); fast::tm_pin16(ra);
#pragma unroll
                            for (int i = 0; i < 4; i++) {
                                zb[i] = make_double2(__hiloint2double((int)rb[4 * i + 1], (int)rb[4 * i]), __hiloint2double((int)rb[4 * i + 3], (int)rb[4 * i + 2]));
                                za[i] = make_double2(__hiloint2double((int)ra[4 * i + 1], (int)ra[4 * i]), __hiloint2double((int)ra[4 * i + 3], (int)ra[4 * i + 2]));
                                zb[i] = cmac_f(zb[i], x[4 * c + i], kcb[4 * c2 + i]); za[i] = cmac_f(za[i], x[4 * c + i], kca[4 * c2 + i]);
                            }
                        }
                        fast::tm_st_c4(tm + 16 * c, zb);
                        fast::tm_st_c4(tm + 32 + 16 * c, za);
                    }
                }
                tm_wait_st();
            }
            // both outputs: (x (X^a - 1)/H for ELL == 1) -> inverse transform -> round -> acc +=
#pragma unroll 1
            for (int pz = 0; pz < 2; pz++) {
                cplx y[8];
                {
                    uint32_t v[2][16];
                    fast::tm_ld16(tm + 32 * pz, v[0]); fast::tm_ld16(tm + 32 * pz + 16, v[1]);
                    tm_wait_ld();
#pragma unroll
                    for (int c = 0; c < 2; c++) {
                        fast::tm_pin16(v[c]);
#pragma unroll
                        for (int i = 0; i < 4; i++)
                            y[4 * c + i] = make_double2(__hiloint2double((int)v[c][4 * i + 1], (int)v[c][4 * i]), __hiloint2double((int)v[c][4 * i + 3], (int)v[c][4 * i + 2]));
                    }
                }
                if (ELL == 1) {
#pragma unroll
                    for (int e = 0; e < 8; e++) {
                        const int b3 = ((e & 1) << 2) | (e & 2) | ((e & 4) >> 2);
                        cplx mo = cmul_f(m1v[0], c_e16[((atv[0] * b3) & 7) * 2]);      // 8th roots = even 16th roots
                        mo.x -= 1.0 / H;
                        y[e] = cmul_f(mo, y[e]);
                    }
                }
                fft_inv(y, xa, xc, tw2, tw3, t, unit_l);
                {
                    uint32_t v[16];
                    fast::tm_ld16(tm + (pz == 0 ? TM_ACCB : TM_ACCA), v);
                    tm_wait_ld();
                    fast::tm_pin16(v);
#pragma unroll
                    for (int m = 0; m < 8; m++) { v[m] += d2torus32(y[m].x); v[m + 8] += d2torus32(-y[m].y); }
                    fast::tm_st16(tm + (pz == 0 ? TM_ACCB : TM_ACCA), v);
                    tm_wait_st();
                }
            }
        }
        if (live) {
            uint32_t *out = a.acc_io + unit * 2 * N;
            uint32_t vb[16], va[16];
            fast::tm_ld16(tm + TM_ACCB, vb);
            fast::tm_ld16(tm + TM_ACCA, va);
            tm_wait_ld();
            fast::tm_pin16(vb); fast::tm_pin16(va);
#pragma unroll
            for (int m = 0; m < 16; m++) { out[t + 64 * m] = vb[m]; out[N + t + 64 * m] = va[m]; }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" :: "r"(*tm_base_s));
}


// reference slot order [poly][8t + e] -> thread order [poly][e][t]
__global__ void k_permute_brk(const cplx *__restrict__ in, cplx *__restrict__ out, size_t polys) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= polys * H) return;
    const size_t p = i / H;
    const int r = (int)(i % H), e = r / UT, t = r % UT;
    out[i] = in[p * H + 8 * t + e];
}

}  // namespace fast32

struct FastKeys32 {
    cplx *brk = nullptr, *t2 = nullptr, *t3 = nullptr, *emono = nullptr;
    bool built = false;
};

static inline bool fast32_supported(const mktfhe_params &p) {
    return p.N == 1024 && (p.scheme == MKTFHE_CGGI || (p.scheme == MKTFHE_LMSS && p.ell == 3)) && p.k == 1;
}

static inline void fast32_free(FastKeys32 &f) {
    if (f.brk) cudaFree(f.brk);
    if (f.t2) cudaFree(f.t2);
    if (f.t3) cudaFree(f.t3);
    if (f.emono) cudaFree(f.emono);
    f = FastKeys32();
}

// Twiddles sqrt(rho(s, i)) = exp(-i*pi*theta(s,i)/2): theta(0,0) = 1/2, theta(s+1, 2i+b) = theta(s,i)/2 + b  (H = 512: 9 stages)
static inline int fast32_build(FastKeys32 &f, const mktfhe_params &p, const cplx *brk_ref, cudaStream_t stream, std::string &err) {
    using namespace fast32;
    fast32_free(f);
    std::vector<__float128> theta(1, (__float128)0.5);
    std::vector<cplx> tw(512, make_double2(0.0, 0.0)), t2(32), t3(256), emono(2048), tw1(8, make_double2(0.0, 0.0));
    const __float128 pi = acosq((__float128)-1);
    for (int s = 0; s < 9; s++) {
        std::vector<__float128> nxt(theta.size() * 2);
        for (size_t i = 0; i < theta.size(); i++) {
            const __float128 ang = pi * theta[i] / 2;
            tw[((size_t)1 << s) + i] = make_double2((double)cosq(ang), (double)-sinq(ang));
            nxt[2 * i] = theta[i] / 2; nxt[2 * i + 1] = theta[i] / 2 + 1;
        }
        theta.swap(nxt);
    }
    for (int blk = 0; blk < 8; blk++) {
        t2[blk] = tw[8 + blk]; t2[8 + blk] = tw[16 + 2 * blk]; t2[16 + blk] = tw[32 + 4 * blk]; t2[24 + blk] = tw[32 + 4 * blk + 2];
    }
    for (int t = 0; t < 64; t++) {
        t3[t] = tw[64 + t]; t3[64 + t] = tw[128 + 2 * t]; t3[128 + t] = tw[256 + 4 * t]; t3[192 + t] = tw[256 + 4 * t + 2];
    }
    for (int m = 0; m < 2048; m++) {
        const __float128 ang = pi * m / 1024;
        emono[m] = make_double2((double)(cosq(ang) / H), (double)(-sinq(ang) / H));
    }
    for (int i = 1; i < 8; i++) tw1[i] = tw[i];
    FCK(cudaMalloc(&f.t2, sizeof(cplx) * 32));
    FCK(cudaMalloc(&f.t3, sizeof(cplx) * 256));
    FCK(cudaMalloc(&f.emono, sizeof(cplx) * 2048));
    FCK(cudaMemcpy(f.t2, t2.data(), sizeof(cplx) * 32, cudaMemcpyHostToDevice));
    FCK(cudaMemcpy(f.t3, t3.data(), sizeof(cplx) * 256, cudaMemcpyHostToDevice));
    FCK(cudaMemcpy(f.emono, emono.data(), sizeof(cplx) * 2048, cudaMemcpyHostToDevice));
    FCK(cudaMemcpyToSymbol(c_tw1, tw1.data(), sizeof(cplx) * 8));
    {   // the 16th-root table is shared with the N = 2048 kernels
        std::vector<cplx> e16(16);
        for (int j = 0; j < 16; j++) { const __float128 ang = pi * j / 8; e16[j] = make_double2((double)cosq(ang), (double)-sinq(ang)); }
        FCK(cudaMemcpyToSymbol(fast::c_e16, e16.data(), sizeof(cplx) * 16));
    }
    if (brk_ref) {
        const size_t polys = (size_t)p.n * 4 * p.l_gsw;
        FCK(cudaMalloc(&f.brk, polys * H * sizeof(cplx)));
        k_permute_brk<<<(unsigned)((polys * H + 255) / 256), 256, 0, stream>>>(brk_ref, f.brk, polys);
        FCK(cudaGetLastError());
    }
    FCK(cudaFuncSetAttribute(k_rgsw_tm<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES));
    FCK(cudaFuncSetAttribute(k_rgsw_tm<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES));
    FCK(cudaFuncSetAttribute(k_rgsw_tma<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES_TMA32));
    FCK(cudaFuncSetAttribute(k_rgsw_tma<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES_TMA32));
    f.built = true;
    return 0;
}

static inline int fast32_launch(FastKeys32 &f, const mktfhe_params &p, fast32::Args a, cudaStream_t stream, int *launches, std::string &err) {
    using namespace fast32;
    if (!f.built) { err = "FAST (N = 1024) keys not built"; return -3; }
    a.brk = f.brk; a.tb = Tables{f.t2, f.t3, f.emono};
    a.n = p.n; a.d = p.d; a.l = p.l_gsw; a.logB = p.logB_gsw; a.lwe_words = (int)mktfhe_lwe_words(&p);
    const unsigned grid = (unsigned)((a.units + U - 1) / U);
    // Keys by per-thread loads (k_rgsw_tm) or through a TMA ring shared by the CTA's eight gates (k_rgsw_tma).  Measured on B200,
    // 4096 gates: LMSS 78.7 ms -> 31.6 ms with the ring (its 48 dependent key loads per digit leave the critical path), CGGI
    // 47.1 ms -> 55.1 ms (one key tile per digit: the ring's mbarrier round trips cost more than the loads they replace).
    // Default here: ring for LMSS, per-thread loads for CGGI; MKTFHE_FAST32_KERNEL = tmem | tma forces one for both (A/B runs).
    // (CGGI gate batches do not come here by default any more: capi.cu sends them to fastw32::k_cggi_w, kernels_fast32_w.cuh.)
    static const int force = []() { const char *e = getenv("MKTFHE_FAST32_KERNEL"); return !e ? 0 : std::string(e) == "tma" ? 1 : std::string(e) == "tmem" ? 2 : 0; }();
    const bool blk = p.scheme == MKTFHE_LMSS && !(a.step_mode == 1);
    const bool tma = force == 1 || (force == 0 && blk);
    if (tma) {
        if (blk) k_rgsw_tma<3><<<grid, CTA_TMA32, SMEM_BYTES_TMA32, stream>>>(a);
        else k_rgsw_tma<1><<<grid, CTA_TMA32, SMEM_BYTES_TMA32, stream>>>(a);
    } else if (blk) k_rgsw_tm<3><<<grid, CTA, SMEM_BYTES, stream>>>(a);
    else k_rgsw_tm<1><<<grid, CTA, SMEM_BYTES, stream>>>(a);
    if (launches) (*launches)++;
    FCK(cudaGetLastError());
    return 0;
}

// ======================================================================================================
// CCS blind rotation, FAST mode: /root/reference/src/tfhe/bootstrapping.jl:234-328.
// One unit = one gate = 64 threads x 8 points (the transform above), 4 units per CTA.  Per (party, key bit) step the
// hybrid product is evaluated component by component, everything of a component staying in registers:
//   u_c  = sum_j D_j(acc_c) * d[j]                                   (:277-284)
//   v_c  = -/+ sum_j D_j(acc_c) * (crs[j] | b_{c-1}[j])              (:286-294)   -> inverse -> 32-bit coefficients
//   w   += sum_j D_j(v_c) * f[j]                                     (:313-320)   (two running accumulators)
// and finally acc_c += ifft(monomial * (u_c [+ w])) for the live components (:322-324).  The accumulator polynomials
// and the u_c live in global memory (L2-resident: (k+1) x 4 KiB + (k+1) x 8 KiB per gate); keys are read through L1,
// which the four gates of a CTA share because they walk the (party, bit) steps together.
// FAST order of the floating-point sums differs from the reference (the w terms are accumulated before u_0 / u_na
// are added); integer stages are the reference's.
namespace fast32 {

constexpr int CU = 4, CCTA = UT * CU;

struct CcsArgs {
    const uint32_t *tilde;          // [B][lwe_words]
    const cplx *const *brk;         // [k] FAST layout [n][lu][3][e][t]
    const cplx *const *pubb;        // [k] FAST layout [lu][e][t], scaled by 1/H
    const cplx *crs;                // FAST layout [lu][e][t], scaled by 1/H
    Tables tb;
    uint32_t *acc;                  // [B][(k+1)][N]
    cplx *tacc;                     // scratch [B][(k+1)][H], thread layout
    int n, k, lu, logB, lwe_words;
    size_t units;
};

constexpr size_t CCS_SMEM_UNIT = (size_t)2 * XB_LEN * 16;
constexpr size_t CCS_SMEM_BYTES = CU * CCS_SMEM_UNIT + (size_t)(32 + 256) * 16;

__global__ void __launch_bounds__(CCTA, 1) k_ccs_fast(const CcsArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int tid = threadIdx.x, unit_l = tid / UT, t = tid % UT;
    cplx *tw2 = reinterpret_cast<cplx *>(smem_raw + CU * CCS_SMEM_UNIT), *tw3 = tw2 + 32;
    for (int i = tid; i < 256; i += CCTA) { if (i < 32) tw2[i] = a.tb.t2[i]; tw3[i] = a.tb.t3[i]; }
    cplx *xa = reinterpret_cast<cplx *>(smem_raw + unit_l * CCS_SMEM_UNIT), *xc = xa + XB_LEN;
    __syncthreads();

    const size_t unit = (size_t)blockIdx.x * CU + unit_l;
    const bool live = unit < a.units;
    const size_t g = live ? unit : 0;
    const int k = a.k, lu = a.lu, logB = a.logB;
    uint32_t *ACC = a.acc + g * (size_t)(k + 1) * N;
    cplx *TACC = a.tacc + g * (size_t)(k + 1) * H;
    const uint32_t *tilde = a.tilde + g * a.lwe_words;

    if (live) {                                     // test vector (bootstrapping.jl:11-23), a components zero
        const uint32_t tb = tilde[0], e8 = 1u << 29;
#pragma unroll
        for (int m = 0; m < 16; m++) {
            const uint32_t i1 = (uint32_t)(t + 64 * m) + 1;
            ACC[t + 64 * m] = tb <= (uint32_t)N ? (i1 <= tb ? e8 : 0u - e8) : (i1 <= tb - (uint32_t)N ? 0u - e8 : e8);
        }
        for (int c = 1; c <= k; c++)
#pragma unroll
            for (int m = 0; m < 16; m++) ACC[(size_t)c * N + t + 64 * m] = 0u;
    }
    const int bit = 32 - lu * logB;
    uint32_t cadd = bit > 0 ? 1u << (bit - 1) : 0u;
    for (int j = 0; j < lu; j++) cadd += 1u << (bit + j * logB + logB - 1);
    const uint32_t mask = (1u << logB) - 1;
    const double dbias = 4503599627370496.0 + (double)(1 << (logB - 1));
    const int brv6t = (int)(__brev((unsigned)t) >> 26);
    const size_t tile = (size_t)H;                  // one key polynomial

    // digit j of 16 cached coefficients (8 low, 8 at +H) -> 8 complex points
    auto digits = [&](const uint32_t (&cf)[16], int j, cplx (&x)[8]) {
        const int sh = bit + (lu - 1 - j) * logB;
#pragma unroll
        for (int m = 0; m < 8; m++) {
            const uint32_t f0 = ((cf[m] + cadd) >> sh) & mask, f1 = ((cf[m + 8] + cadd) >> sh) & mask;
            x[m] = make_double2(__hiloint2double(0x43300000, (int)f0) - dbias, dbias - __hiloint2double(0x43300000, (int)f1));
        }
    };

    for (int idx = 0; idx < k; idx++) {
        const int na = idx + 1;                                     // live a-components a_1 .. a_na
        for (int i = 0; i < a.n; i++) {
            __syncthreads();                                        // the gates of a CTA walk the steps together (L1 key reuse)
            const uint32_t at = live ? tilde[1 + (size_t)idx * a.n + i] : 0u;
            if (at == 0) continue;                                  // :262
            const cplx *uni = a.brk[idx] + (size_t)i * 3 * lu * tile + t;      // [j][d, f.b, f.a][e][t]
            cplx wb[8], wa[8];
#pragma unroll
            for (int e = 0; e < 8; e++) wb[e] = wa[e] = make_double2(0.0, 0.0);

            for (int c = 0; c <= na; c++) {
                uint32_t cf[16];
#pragma unroll
                for (int m = 0; m < 8; m++) { cf[m] = ACC[(size_t)c * N + t + 64 * m]; cf[m + 8] = ACC[(size_t)c * N + t + 64 * m + H]; }
                cplx u[8], v[8];
#pragma unroll
                for (int e = 0; e < 8; e++) u[e] = v[e] = make_double2(0.0, 0.0);
                const cplx *kv = (c == 0 ? a.crs : a.pubb[c - 1]) + t;
                for (int j = 0; j < lu; j++) {
                    cplx x[8], kd[8], kw[8];
                    digits(cf, j, x);
                    fft_fwd(x, xa, xc, tw2, tw3, t, unit_l, [&]() {
#pragma unroll
                        for (int e = 0; e < 8; e++) { kd[e] = __ldg(uni + (size_t)(j * 3) * tile + e * UT); kw[e] = __ldg(kv + (size_t)j * tile + e * UT); }
                    });
#pragma unroll
                    for (int e = 0; e < 8; e++) { u[e] = cmac_f(u[e], x[e], kd[e]); v[e] = cmac_f(v[e], x[e], kw[e]); }
                }
                if (c == 0) {                                       // v0 = - sum (mulsubto!, :288-290)
#pragma unroll
                    for (int e = 0; e < 8; e++) v[e] = make_double2(-v[e].x, -v[e].y);
                }
#pragma unroll
                for (int e = 0; e < 8; e++) TACC[(size_t)c * H + e * UT + t] = u[e];
                fft_inv(v, xa, xc, tw2, tw3, t, unit_l);            // crs / pubb carry the 1/H
#pragma unroll
                for (int m = 0; m < 8; m++) { cf[m] = d2torus32(v[m].x); cf[m + 8] = d2torus32(-v[m].y); }
                for (int j = 0; j < lu; j++) {
                    cplx x[8], kb[8], ka[8];
                    digits(cf, j, x);
                    fft_fwd(x, xa, xc, tw2, tw3, t, unit_l, [&]() {
#pragma unroll
                        for (int e = 0; e < 8; e++) { kb[e] = __ldg(uni + (size_t)(j * 3 + 1) * tile + e * UT); ka[e] = __ldg(uni + (size_t)(j * 3 + 2) * tile + e * UT); }
                    });
#pragma unroll
                    for (int e = 0; e < 8; e++) { wb[e] = cmac_f(wb[e], x[e], kb[e]); wa[e] = cmac_f(wa[e], x[e], ka[e]); }
                }
            }
            // acc_c += ifft(monomial * tacc_c); tacc_0 = u_0 + w_b, tacc_na = u_na + w_a  (:322-324)
            const cplx m1 = __ldg(&a.tb.emono[((4 * brv6t + 1) * at) & 2047]);
            for (int c = 0; c <= na; c++) {
                cplx y[8];
#pragma unroll
                for (int e = 0; e < 8; e++) {
                    y[e] = TACC[(size_t)c * H + e * UT + t];
                    if (c == 0) { y[e].x += wb[e].x; y[e].y += wb[e].y; }
                    if (c == na) { y[e].x += wa[e].x; y[e].y += wa[e].y; }
                    const int b3 = ((e & 1) << 2) | (e & 2) | ((e & 4) >> 2);
                    cplx mo = cmul_f(m1, c_e16[((at * b3) & 7) * 2]);
                    mo.x -= 1.0 / H;
                    y[e] = cmul_f(mo, y[e]);
                }
                fft_inv(y, xa, xc, tw2, tw3, t, unit_l);
#pragma unroll
                for (int m = 0; m < 8; m++) {
                    ACC[(size_t)c * N + t + 64 * m] += d2torus32(y[m].x);
                    ACC[(size_t)c * N + t + 64 * m + H] += d2torus32(-y[m].y);
                }
            }
        }
    }
}

// ---- gadget product hook: the building block of the FAST CCS hybrid product on caller-supplied inputs ---------------------------
// out[c] = native(ifft( Sum_j fft(D_j(poly)) (.) key[j][c] ))  for c < ncomp, with the SAME device functions k_ccs_fast is made of
// (field-extraction digits, fft_fwd, multiply-accumulate, fft_inv, d2torus32).  The hybrid product re-decomposes a computed
// polynomial (v) inside one step, so whole-step coefficients cannot be compared between two roundings; this hook lets a test feed the
// ORACLE's intermediate polynomial into one product and compare within a per-product tolerance
// (bootstrapping.jl:277-294 u_c / v_c from an accumulator component, :313-320 w from v).  keys: [l][ncomp][H] in the reference slot order.
struct Gp32Args {
    const uint32_t *polys;      // [B][N]
    const cplx *keys;           // [l][ncomp][H]
    uint32_t *out;              // [B][ncomp][N]
    Tables tb;
    int l, logB, ncomp;
    size_t units;
};
__global__ void __launch_bounds__(CCTA, 1) k_gadget_product32(const Gp32Args a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int tid = threadIdx.x, unit_l = tid / UT, t = tid % UT;
    cplx *tw2 = reinterpret_cast<cplx *>(smem_raw + CU * CCS_SMEM_UNIT), *tw3 = tw2 + 32;
    for (int i = tid; i < 256; i += CCTA) { if (i < 32) tw2[i] = a.tb.t2[i]; tw3[i] = a.tb.t3[i]; }
    cplx *xa = reinterpret_cast<cplx *>(smem_raw + unit_l * CCS_SMEM_UNIT), *xc = xa + XB_LEN;
    __syncthreads();
    const size_t unit = (size_t)blockIdx.x * CU + unit_l;
    if (unit >= a.units) return;                    // the transforms below synchronise per unit only
    const int l = a.l, logB = a.logB, bit = 32 - l * logB;
    uint32_t cadd = bit > 0 ? 1u << (bit - 1) : 0u;
    for (int j = 0; j < l; j++) cadd += 1u << (bit + j * logB + logB - 1);
    const uint32_t mask = (1u << logB) - 1;
    const double dbias = 4503599627370496.0 + (double)(1 << (logB - 1));
    uint32_t cf[16];
    const uint32_t *src = a.polys + unit * N;
#pragma unroll
    for (int m = 0; m < 8; m++) { cf[m] = src[t + 64 * m]; cf[m + 8] = src[t + 64 * m + H]; }
    for (int c = 0; c < a.ncomp; c++) {
        cplx acc[8];
#pragma unroll
        for (int e = 0; e < 8; e++) acc[e] = make_double2(0.0, 0.0);
        for (int j = 0; j < l; j++) {
            const int sh = bit + (l - 1 - j) * logB;
            cplx x[8];
#pragma unroll
            for (int m = 0; m < 8; m++) {
                const uint32_t f0 = ((cf[m] + cadd) >> sh) & mask, f1 = ((cf[m + 8] + cadd) >> sh) & mask;
                x[m] = make_double2(__hiloint2double(0x43300000, (int)f0) - dbias, dbias - __hiloint2double(0x43300000, (int)f1));
            }
            fft_fwd(x, xa, xc, tw2, tw3, t, unit_l, []() {});
            const cplx *k = a.keys + ((size_t)j * a.ncomp + c) * H + 8 * t;           // slot 8t + e
#pragma unroll
            for (int e = 0; e < 8; e++) acc[e] = cmac_f(acc[e], x[e], k[e]);
        }
#pragma unroll
        for (int e = 0; e < 8; e++) acc[e] = make_double2(acc[e].x * (1.0 / H), acc[e].y * (1.0 / H));
        fft_inv(acc, xa, xc, tw2, tw3, t, unit_l);
        uint32_t *dst = a.out + (unit * a.ncomp + c) * N;
#pragma unroll
        for (int m = 0; m < 8; m++) { dst[t + 64 * m] = d2torus32(acc[m].x); dst[t + 64 * m + H] = d2torus32(-acc[m].y); }
    }
}

// reference slot order -> thread order, with an optional scale (1/H for the keys that feed an inverse transform)
__global__ void k_permute_scale(const cplx *__restrict__ in, cplx *__restrict__ out, size_t polys, double scale) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= polys * H) return;
    const size_t p = i / H;
    const int r = (int)(i % H), e = r / UT, t = r % UT;
    const cplx z = in[p * H + 8 * t + e];
    out[i] = make_double2(z.x * scale, z.y * scale);
}

}  // namespace fast32

struct FastCcsKeys {
    std::vector<cplx *> brk, pubb;
    cplx **d_brk = nullptr, **d_pubb = nullptr, *crs = nullptr;
    bool built = false;
};

static inline bool fastccs_supported(const mktfhe_params &p) { return p.scheme == MKTFHE_CCS && p.N == 1024; }

static inline void fastccs_free(FastCcsKeys &f) {
    for (auto &q : f.brk) if (q) cudaFree(q);
    for (auto &q : f.pubb) if (q) cudaFree(q);
    if (f.d_brk) cudaFree(f.d_brk);
    if (f.d_pubb) cudaFree(f.d_pubb);
    if (f.crs) cudaFree(f.crs);
    f = FastCcsKeys();
}

static inline int fastccs_build(FastCcsKeys &f, const mktfhe_params &p, const std::vector<cplx *> &brk_ref, const std::vector<cplx *> &pubb_ref,
                                const cplx *crs_ref, cudaStream_t stream, std::string &err) {
    using namespace fast32;
    fastccs_free(f);
    const size_t brk_polys = (size_t)p.n * 3 * p.l_uni, small = (size_t)p.l_uni;
    auto perm = [&](const cplx *src, cplx **dst, size_t polys, double scale) -> int {
        FCK(cudaMalloc(dst, polys * H * sizeof(cplx)));
        k_permute_scale<<<(unsigned)((polys * H + 255) / 256), 256, 0, stream>>>(src, *dst, polys, scale);
        FCK(cudaGetLastError());
        return 0;
    };
    f.brk.assign(brk_ref.size(), nullptr); f.pubb.assign(brk_ref.size(), nullptr);
    int rc;
    for (size_t i = 0; i < brk_ref.size(); i++) {
        if ((rc = perm(brk_ref[i], &f.brk[i], brk_polys, 1.0))) return rc;
        if ((rc = perm(pubb_ref[i], &f.pubb[i], small, 1.0 / H))) return rc;
    }
    if ((rc = perm(crs_ref, &f.crs, small, 1.0 / H))) return rc;
    FCK(cudaMalloc(&f.d_brk, sizeof(cplx *) * f.brk.size()));
    FCK(cudaMalloc(&f.d_pubb, sizeof(cplx *) * f.pubb.size()));
    FCK(cudaMemcpy(f.d_brk, f.brk.data(), sizeof(cplx *) * f.brk.size(), cudaMemcpyHostToDevice));
    FCK(cudaMemcpy(f.d_pubb, f.pubb.data(), sizeof(cplx *) * f.pubb.size(), cudaMemcpyHostToDevice));
    FCK(cudaFuncSetAttribute(k_ccs_fast, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)CCS_SMEM_BYTES));
    f.built = true;
    return 0;
}

static inline int fast32_gadget_product(FastKeys32 &tabs, const uint32_t *polys, const cplx *keys, uint32_t *out, int l, int logB, int ncomp,
                                        size_t batch, cudaStream_t stream, int *launches, std::string &err) {
    using namespace fast32;
    if (!tabs.t2) { err = "FAST (N = 1024) tables not built"; return -3; }
    Gp32Args a{};
    a.polys = polys; a.keys = keys; a.out = out; a.tb = Tables{tabs.t2, tabs.t3, tabs.emono};
    a.l = l; a.logB = logB; a.ncomp = ncomp; a.units = batch;
    FCK(cudaFuncSetAttribute(k_gadget_product32, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)CCS_SMEM_BYTES));
    k_gadget_product32<<<(unsigned)((batch + CU - 1) / CU), CCTA, CCS_SMEM_BYTES, stream>>>(a);
    if (launches) (*launches)++;
    FCK(cudaGetLastError());
    return 0;
}

static inline int fastccs_launch(FastCcsKeys &f, FastKeys32 &tabs, const mktfhe_params &p, const uint32_t *tilde, uint32_t *acc, cplx *tacc,
                                 size_t gates, cudaStream_t stream, int *launches, std::string &err) {
    using namespace fast32;
    if (!f.built || !tabs.t2) { err = "FAST CCS keys not built"; return -3; }
    fast32::CcsArgs a{};
    a.tilde = tilde; a.brk = f.d_brk; a.pubb = f.d_pubb; a.crs = f.crs; a.tb = Tables{tabs.t2, tabs.t3, tabs.emono};
    a.acc = acc; a.tacc = tacc; a.n = p.n; a.k = p.k; a.lu = p.l_uni; a.logB = p.logB_uni; a.lwe_words = (int)mktfhe_lwe_words(&p);
    a.units = gates;
    k_ccs_fast<<<(unsigned)((gates + CU - 1) / CU), CCTA, CCS_SMEM_BYTES, stream>>>(a);
    if (launches) (*launches)++;
    FCK(cudaGetLastError());
    return 0;
}
