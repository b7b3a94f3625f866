// kernels_fast32_w.cuh -- FAST-mode CGGI blind rotation, "one half-warp per transform" variant.
//
// What it computes: /root/reference/src/tfhe/bootstrapping.jl:32-76 (blindrotate! of CGGI), exactly like fast32::k_rgsw_tm<1>
// (kernels_fast32.cuh) -- same integer stages, same twist-free product-tree transform, same slot order -- with the mapping of
// fastw::k_phase1_w (kernels_fast_w.cuh) scaled to N = 1024:
//
//   * A 512-point transform belongs to ONE HALF-WARP: 16 threads x 32 points, passes of 5 + 4 stages in registers and a single
//     transposition through shared memory in between (__syncwarp, no block barrier).  The 64-thread x 8-point mapping needs
//     three passes of three stages, i.e. two exchanges and four named barriers per transform, and was bound by the L1 / shared
//     data pipe (80 % busy next to an FP64 pipe at 56 %); this mapping moves half as many bytes through it.
//   * A warp runs the transforms of TWO gates side by side (lanes 0..15 and 16..31).  The two gates of a TMEM lane quadrant are
//     shared by two warps q and q + 4 (same quadrant, same scheduler): warp A owns the .b halves of both RLWE accumulators
//     (decomposes and transforms the l gadget digits of acc.b, inverse-transforms the .b sum, updates acc.b), warp B the .a halves.
//     Every spectrum is multiplied into BOTH RGSW sums; four mbarrier tokens per quadrant fix the order of the additions
//     (.b: A0, B0, A1, B1, ...; .a: B0, A0, B1, A1, ...), so results are deterministic.  No block barrier in the step loop.
//   * All gates use the same key, so the eight gates of a CTA share one stream of key tiles (8 KiB polynomials in thread order
//     [e < 32][t < 16]) through a cp.async.bulk + mbarrier ring; both half-warps of a warp read the same tile addresses
//     (shared-memory broadcast: one wavefront pair per 16-byte load instead of four).
//   * A rotation of zero (bootstrapping.jl:48 skips the step) needs no branch: (X^0 - 1)/H is exactly zero in the monomial's
//     closed form, so the step adds exactly zero to that gate's accumulator while its neighbour in the warp does real work.
//
// Index math of the transform is modelled and checked against the oracle in tools/models/fft16x32_model.py.
#pragma once
#include "kernels_fast32.cuh"
#include "kernels_fast_w.cuh"

namespace fastw32 {

using fast32::H; using fast32::N;
using fast::bf; using fast::bf_mi; using fast::bi; using fast::bi_mi;
using fast::tm_ld16; using fast::tm_st16; using fast::tm_wait_ld; using fast::tm_wait_st; using fast::tm_pin16; using fast::tm_st_c4;
using fast::mb_init; using fast::mb_expect_tx; using fast::bulk_g2s;
using fast32::d2torus32;
using fastw::smem_u32; using fastw::mbs_wait; using fastw::mbs_arrive; using fastw::Token; using fastw::token_pass; using fastw::unpack_c;
using fastw::tm_fence_before; using fastw::tm_fence_after; using fastw::c_e32;

constexpr int WQ = 4;                        // TMEM lane quadrants = pairs of gates per CTA
constexpr int GATES_CTA = 2 * WQ;
constexpr int NCW = 2 * WQ;                  // consumer warps
constexpr int CTA_W = NCW * 32 + 128;        // + producer warpgroup (setmaxnreg works on warpgroups)
constexpr int XBH = H + 16;                  // exchange buffer of a half-warp: position q at q + (q >> 5)
#ifndef W32_RING
#define W32_RING 6
#endif
constexpr int RINGW = W32_RING;              // key tiles (8 KiB polynomials) in flight; 4, 6, 11 measured equal (41.6 ms), 8 one percent slower
#ifndef W32_CREGS
#define W32_CREGS 232
#endif
constexpr int W_LAUNCH_REGS = 168, W_CONSUMER_REGS = W32_CREGS, W_PRODUCER_REGS = 24;
static_assert(NCW * W_CONSUMER_REGS + 4 * W_PRODUCER_REGS <= (CTA_W / 32) * W_LAUNCH_REGS, "setmaxnreg over-subscription");
static_assert(RINGW % 2 == 0, "ring depth must be even: a slot must always serve the same kind of warp (see kernels_fast_w.cuh)");
constexpr size_t SMEM_BYTES_W = ((size_t)NCW * 2 * XBH + 256 + (size_t)RINGW * H) * 16 + (2 * RINGW + 4 * WQ) * 8 + 16;
static_assert(SMEM_BYTES_W <= 232448, "shared memory budget");

__constant__ double2 c_tw1h[32];     // TW[1..31] of the 9-level tree: stages 1..5 (index 2^s + node), entry 0 unused

// TMEM columns of a lane: RGSW sums (32 complex each), RLWE accumulator halves (coefficient t + 16m at 2m, t + 16m + H at 2m + 1)
constexpr uint32_t TMW_TACC_B = 0, TMW_TACC_A = 128, TMW_ACC_B = 256, TMW_ACC_A = 320;

struct Tables {
    const cplx *t2h;       // [16][16] pass-2 twiddles per thread: row 0 TW[32+2t]; 1..2 TW[64+4t+2g]; 3..6 TW[128+8t+2g]; 7..14 TW[256+16t+2g]
    const cplx *emono;     // [2048] exp(-i*pi*m/1024) / H
};
struct Args {
    const uint32_t *tilde;        // [B][lwe_words] / step mode: [units] rotations
    const cplx *brk;              // thread order [idx][dg][comp][e < 32][t < 16]
    Tables tb;
    uint32_t *acc_io;             // out [B][2][N] (step mode: in/out)
    int step_mode, step_idx;
    int n, l, logB, lwe_words;
    size_t units;
};

// ---- transform: x (layout A: element m = point t + 16m) <-> x (layout C: element e = slot 32t + e) ------------------------
__device__ __forceinline__ void pass1_fwd(cplx (&x)[32]) {
#pragma unroll
    for (int s = 0; s < 5; s++) {
#pragma unroll
        for (int m = 0; m < 32; m++) {
            const int half = 16 >> s;
            if (m & half) continue;
            const int node = m >> (5 - s);
            if (node & 1) bf_mi(x[m], x[m + half], c_tw1h[(1 << s) + (node & ~1)]);
            else bf(x[m], x[m + half], c_tw1h[(1 << s) + node]);
        }
    }
}
__device__ __forceinline__ void pass1_inv(cplx (&x)[32]) {
#pragma unroll
    for (int s = 4; s >= 0; s--) {
#pragma unroll
        for (int m = 0; m < 32; m++) {
            const int half = 16 >> s;
            if (m & half) continue;
            const int node = m >> (5 - s);
            if (node & 1) bi_mi(x[m], x[m + half], c_tw1h[(1 << s) + (node & ~1)]);
            else bi(x[m], x[m + half], c_tw1h[(1 << s) + node]);
        }
    }
}
// stages 6..9 on the thread's 32 contiguous slots (two 16-point blocks); tw = shared table [16][16] + lane
__device__ __forceinline__ void pass2_fwd(cplx (&x)[32], const cplx *__restrict__ tw) {
#pragma unroll
    for (int s = 6; s < 10; s++) {
#pragma unroll
        for (int e = 0; e < 32; e++) {
            const int half = 1 << (9 - s);
            if (e & half) continue;
            const int node = e >> (10 - s);
            const cplx w = tw[(((1 << (s - 6)) - 1) + (node >> 1)) * 16];
            if (node & 1) bf_mi(x[e], x[e + half], w); else bf(x[e], x[e + half], w);
        }
    }
}
__device__ __forceinline__ void pass2_inv(cplx (&x)[32], const cplx *__restrict__ tw) {
#pragma unroll
    for (int s = 9; s >= 6; s--) {
#pragma unroll
        for (int e = 0; e < 32; e++) {
            const int half = 1 << (9 - s);
            if (e & half) continue;
            const int node = e >> (10 - s);
            const cplx w = tw[(((1 << (s - 6)) - 1) + (node >> 1)) * 16];
            if (node & 1) bi_mi(x[e], x[e + half], w); else bi(x[e], x[e + half], w);
        }
    }
}
// xb = this half-warp's exchange buffer, t = lane & 15
__device__ __forceinline__ void fft_fwd(cplx (&x)[32], cplx *xb, const cplx *tw, int t) {
    pass1_fwd(x);
    __syncwarp();                                       // earlier readers of xb are done
#pragma unroll
    for (int m = 0; m < 32; m++) xb[t + 16 * m + (m >> 1)] = x[m];
    __syncwarp();
#pragma unroll
    for (int e = 0; e < 32; e++) x[e] = xb[33 * t + e];
    pass2_fwd(x, tw);
}
__device__ __forceinline__ void fft_inv(cplx (&x)[32], cplx *xb, const cplx *tw, int t) {
    pass2_inv(x, tw);
    __syncwarp();
#pragma unroll
    for (int e = 0; e < 32; e++) xb[33 * t + e] = x[e];
    __syncwarp();
#pragma unroll
    for (int m = 0; m < 32; m++) x[m] = xb[t + 16 * m + (m >> 1)];
    pass1_inv(x);
}

struct RingPos {
    uint32_t slot, par;
    __device__ __forceinline__ void advance(uint32_t by) { slot += by; if (slot >= RINGW) { slot -= RINGW; par ^= 1; } }
};

__global__ void __launch_bounds__(CTA_W, 1) k_cggi_w(const Args a) {
    constexpr int SUSPEND = 0;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, t = lane & 15, hw = lane >> 4;
    cplx *xb_all = reinterpret_cast<cplx *>(smem_raw);
    cplx *tw2s = xb_all + (size_t)NCW * 2 * XBH;
    cplx *ring = tw2s + 256;
    uint64_t *full = reinterpret_cast<uint64_t *>(ring + (size_t)RINGW * H), *empty = full + RINGW;
    uint64_t *tokens = empty + RINGW;                               // [WQ][4]: b A->B, b B->A, a B->A, a A->B
    uint32_t *tm_base_s = reinterpret_cast<uint32_t *>(tokens + 4 * WQ);
    for (int i = tid; i < 256; i += CTA_W) tw2s[i] = a.tb.t2h[i];
    if (tid == 0) {
        for (int s = 0; s < RINGW; s++) { mb_init(&full[s], 1); mb_init(&empty[s], WQ); }
        for (int s = 0; s < 4 * WQ; s++) mb_init(&tokens[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" :: "r"((uint32_t)__cvta_generic_to_shared(tm_base_s)));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    tm_fence_before();
    __syncthreads();
    tm_fence_after();

    const int l = a.l;
    const size_t per_idx = (size_t)4 * l * H;
    const int nsteps = a.step_mode ? 1 : a.n;
    const uint32_t ntiles = (uint32_t)nsteps * 2 * l * 2;

    if (warp >= NCW) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 24;");
        // ---- producer: 8 KiB polynomials in thread order [e][t], in consumption order: per step and j < l the four tiles
        //      A_j.b = (digit j, .b), B_j.a = (digit l + j, .a), A_j.a = (digit j, .a), B_j.b = (digit l + j, .b)
        if (tid == NCW * 32) {
            uint32_t slot = 0, par = 1, step = 0, j = 0, r = 0;         // par: parity of the PREVIOUS phase of empty[slot]
            for (uint32_t n = 0; n < ntiles; n++) {
                if (n >= RINGW) fast::mb_wait_suspend(&empty[slot], par);
                const uint32_t dg = (r & 1) ? (uint32_t)l + j : j, comp = (r == 1 || r == 2) ? 1u : 0u;
                const int idx = a.step_mode ? a.step_idx : (int)step;
                mb_expect_tx(&full[slot], H * 16);
                bulk_g2s(ring + (size_t)slot * H, a.brk + (size_t)idx * per_idx + (size_t)(dg * 2 + comp) * H, H * 16, &full[slot]);
                if (++slot == RINGW) { slot = 0; par ^= 1; }
                if (++r == 4) { r = 0; if (++j == (uint32_t)l) { j = 0; step++; } }
            }
        }
    } else {
        asm volatile("setmaxnreg.inc.sync.aligned.u32 " W_STR(W32_CREGS) ";");
        // ---- consumers: quadrant q = two gates (half-warps), role w (0 = A: the .b halves, 1 = B: the .a halves)
        const int q = warp & 3, w = warp >> 2;
        const uint32_t tk = smem_u32(tokens + 4 * q), full_s = smem_u32(full), empty_s = smem_u32(empty);
        Token<SUSPEND> wait_own{tk + 8u * (w == 0 ? 1 : 3), 0u}, wait_oth{tk + 8u * (w == 0 ? 2 : 0), 0u};
        const uint32_t pass_own = tk + 8u * (w == 0 ? 0 : 2), pass_oth = tk + 8u * (w == 0 ? 3 : 1);
        const uint32_t tm = *tm_base_s + ((uint32_t)(32 * q) << 16);
        cplx *xb = xb_all + ((size_t)warp * 2 + hw) * XBH;
        const cplx *tw = tw2s + t;
        const size_t unit = ((size_t)blockIdx.x * WQ + q) * 2 + hw;         // this half-warp's gate
        const bool live = unit < a.units;
        const size_t gate = live ? unit : 0;
        const uint32_t tm_acc = tm + (w == 0 ? TMW_ACC_B : TMW_ACC_A), tm_tacc = tm + (w == 0 ? TMW_TACC_B : TMW_TACC_A);

        const int logB = a.logB, bit = 32 - l * logB;
        // divbits rounding + balanced-digit carry chain in one add (arithmetic.jl:23-27, gsw.jl:86-96).  The accumulator is KEPT
        // with this constant added: digits are then plain bit fields of the stored words; it comes off when the accumulator leaves.
        uint32_t cadd = bit > 0 ? 1u << (bit - 1) : 0u;
        for (int j = 0; j < l; j++) cadd += 1u << (bit + j * logB + logB - 1);
        {   // each warp initialises its halves of the two RLWE accumulators (+ cadd) and of the sums
            uint32_t z[16];
#pragma unroll
            for (int i = 0; i < 16; i++) z[i] = 0u;
#pragma unroll
            for (int c = 0; c < 8; c++) tm_st16(tm_tacc + 16 * c, z);
            if (!a.step_mode) {                     // test vector (bootstrapping.jl:11-23) in .b, zero in .a
                const uint32_t tb = a.tilde[gate * a.lwe_words], e8 = 1u << 29;
#pragma unroll
                for (int c = 0; c < 4; c++) {
                    uint32_t v[16];
#pragma unroll
                    for (int i = 0; i < 16; i++) {
                        const uint32_t k = (uint32_t)(t + 16 * (8 * c + (i >> 1)) + (i & 1) * H), i1 = k + 1;      // 1-based coefficient index
                        const uint32_t tv = tb <= (uint32_t)N ? (i1 <= tb ? e8 : 0u - e8) : (i1 <= tb - (uint32_t)N ? 0u - e8 : e8);
                        v[i] = (w == 0 ? tv : 0u) + cadd;
                    }
                    tm_st16(tm_acc + 16 * c, v);
                }
            } else {
                const uint32_t *src = a.acc_io + gate * 2 * N + (size_t)w * N;
#pragma unroll
                for (int c = 0; c < 4; c++) {
                    uint32_t v[16];
#pragma unroll
                    for (int i = 0; i < 16; i++) v[i] = src[t + 16 * (8 * c + (i >> 1)) + (i & 1) * H] + cadd;
                    tm_st16(tm_acc + 16 * c, v);
                }
            }
            tm_wait_st();
        }
        const uint32_t mask = (1u << logB) - 1;
        const double dbias = 4503599627370496.0 + (double)(1 << (logB - 1));
        const uint32_t *at_src = a.step_mode ? a.tilde + gate : a.tilde + gate * a.lwe_words + 1;
        const int brv4t = (int)(__brev((unsigned)t) >> 28);
        RingPos rp{(uint32_t)w, 0u};                 // this warp consumes every second tile of the producer's sequence, starting at tile w

        for (int step = 0; step < nsteps; step++) {
            const uint32_t at = live ? at_src[a.step_mode ? 0 : step] : 0u;      // 0: the step adds exactly zero (see the header)
            const cplx m1 = __ldg(&a.tb.emono[((4 * brv4t + 1) * at) & 2047]);

            for (int j = 0; j < l; j++) {
                const int sh = bit + (l - 1 - j) * logB;             // digit j (0 = most significant) of my halves of the accumulators
                cplx x[32];
                {
                    uint32_t v[4][16];
#pragma unroll
                    for (int c = 0; c < 4; c++) tm_ld16(tm_acc + 16 * c, v[c]);
                    tm_wait_ld();
#pragma unroll
                    for (int c = 0; c < 4; c++) {
                        tm_pin16(v[c]);
#pragma unroll
                        for (int i = 0; i < 8; i++) {
                            const uint32_t f0 = (v[c][2 * i] >> sh) & mask, f1 = (v[c][2 * i + 1] >> sh) & mask;
                            // signed(d_k) - im*signed(d_{k+H}); 2^52 + field is exact in the double's mantissa
                            x[8 * c + i] = make_double2(__hiloint2double(0x43300000, (int)f0) - dbias, dbias - __hiloint2double(0x43300000, (int)f1));
                        }
                    }
                }
                fft_fwd(x, xb, tw, t);
                // RGSW sums += spectrum x key: my own sum first, the other warp's sum second (the other warp goes the other way round)
#pragma unroll
                for (int ps = 0; ps < 2; ps++) {
                    const uint32_t tmz = ps == 0 ? tm_tacc : tm + (w == 0 ? TMW_TACC_A : TMW_TACC_B);
                    mbs_wait<SUSPEND>(full_s + 8u * rp.slot, rp.par);
                    const cplx *kp = ring + (size_t)rp.slot * H + t;
                    if (ps == 0) { if (j > 0) wait_own.wait(); } else wait_oth.wait();
                    tm_fence_after();
                    uint32_t v[2][16];
                    tm_ld16(tmz, v[0]);
#pragma unroll
                    for (int c = 0; c < 8; c++) {
                        cplx kc[4], z[4];
#pragma unroll
                        for (int i = 0; i < 4; i++) kc[i] = kp[(4 * c + i) * 16];
                        tm_wait_ld();
                        tm_pin16(v[c & 1]);
                        if (c < 7) tm_ld16(tmz + 16 * (c + 1), v[(c + 1) & 1]);
#pragma unroll
                        for (int i = 0; i < 4; i++) z[i] = cmac_f(unpack_c(v[c & 1], i), x[4 * c + i], kc[i]);
                        tm_st_c4(tmz + 16 * c, z);
                    }
                    tm_wait_st();
                    tm_fence_before();
                    token_pass(ps == 0 ? pass_own : pass_oth, lane);   // also orders every lane's key reads before the release below
                    if (lane == 0) mbs_arrive(empty_s + 8u * rp.slot);  // this key tile is done
                    rp.advance(2);
                }
            }
            wait_own.wait();                                        // the other warp's last addition into my sums
            tm_fence_after();
            // this warp's outputs: (x (X^a - 1)/H) -> inverse transform -> round -> acc +=
            {
                cplx y[32];
#pragma unroll
                for (int hb = 0; hb < 2; hb++) {
                    uint32_t v[4][16];
#pragma unroll
                    for (int c = 0; c < 4; c++) tm_ld16(tm_tacc + 64 * hb + 16 * c, v[c]);
                    tm_wait_ld();
#pragma unroll
                    for (int c = 0; c < 4; c++) {
                        tm_pin16(v[c]);
#pragma unroll
                        for (int i = 0; i < 4; i++) y[16 * hb + 4 * c + i] = unpack_c(v[c], i);
                    }
                }
                {   // clear the sums for the next step (every product accumulates, also the first)
                    uint32_t z[16];
#pragma unroll
                    for (int i = 0; i < 16; i++) z[i] = 0u;
#pragma unroll
                    for (int c = 0; c < 8; c++) tm_st16(tm_tacc + 16 * c, z);
                }
                // slot 32t + e evaluates at exp(-i*pi*(4*brv9(n)+1)/N), brv9(n) = 16*brv5(e) + brv4(t)
#pragma unroll
                for (int e = 0; e < 32; e++) {
                    const int b5 = ((e & 1) << 4) | ((e & 2) << 2) | (e & 4) | ((e & 8) >> 2) | ((e & 16) >> 4);
                    cplx mo = cmul_f(m1, c_e32[(at * b5) & 31]);
                    mo.x -= 1.0 / H;
                    const double pr = mo.y * y[e].y, pi = mo.y * y[e].x;
                    y[e] = make_double2(fma(mo.x, y[e].x, -pr), fma(mo.x, y[e].y, pi));
                }
                fft_inv(y, xb, tw, t);
                uint32_t v[2][16];
                tm_ld16(tm_acc, v[0]);
#pragma unroll
                for (int c = 0; c < 4; c++) {
                    tm_wait_ld();
                    tm_pin16(v[c & 1]);
                    if (c < 3) tm_ld16(tm_acc + 16 * (c + 1), v[(c + 1) & 1]);      // next chunk in flight during the rounding
#pragma unroll
                    for (int i = 0; i < 8; i++) {
                        const int m = 8 * c + i;
                        v[c & 1][2 * i] += d2torus32(y[m].x);
                        v[c & 1][2 * i + 1] += d2torus32(-y[m].y);
                    }
                    tm_st16(tm_acc + 16 * c, v[c & 1]);
                }
                tm_wait_st();
            }
        }

        {   // the accumulator halves leave the kernel without the decomposition constant
            uint32_t *dst = a.acc_io + gate * 2 * N + (size_t)w * N;
#pragma unroll
            for (int c = 0; c < 4; c++) {
                uint32_t v[16];
                tm_ld16(tm_acc + 16 * c, v);
                tm_wait_ld();
                tm_pin16(v);
                if (live) {
#pragma unroll
                    for (int i = 0; i < 16; i++) dst[t + 16 * (8 * c + (i >> 1)) + (i & 1) * H] = v[i] - cadd;
                }
            }
        }
    }
    tm_fence_before();
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" :: "r"(*tm_base_s));
}

// reference slot order [poly][32t + e] -> thread order [poly][e][t]
__global__ void k_permute_brk_h(const cplx *__restrict__ in, cplx *__restrict__ out, size_t polys) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= polys * H) return;
    const size_t p = i / H;
    const int r = (int)(i % H), e = r / 16, t = r % 16;
    out[i] = in[p * H + 32 * t + e];
}

}  // namespace fastw32

struct FastKeys32W {
    cplx *brk = nullptr, *t2h = nullptr;
    bool built = false;
};
static inline void fast32w_free(FastKeys32W &f) {
    if (f.brk) cudaFree(f.brk);
    if (f.t2h) cudaFree(f.t2h);
    f = FastKeys32W();
}
// Twiddles of the 9-level product tree as in fast32_build; brk_ref in the reference slot order
static inline int fast32w_build(FastKeys32W &f, const mktfhe_params &p, const cplx *brk_ref, cudaStream_t stream, std::string &err) {
    using namespace fastw32;
    fast32w_free(f);
    std::vector<__float128> theta(1, (__float128)0.5);
    std::vector<cplx> tw(512, make_double2(0.0, 0.0)), t2h(256, make_double2(0.0, 0.0)), tw1(32, make_double2(0.0, 0.0));
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
    for (int i = 1; i < 32; i++) tw1[i] = tw[i];
    for (int t = 0; t < 16; t++) {
        t2h[t] = tw[32 + 2 * t];
        for (int g = 0; g < 2; g++) t2h[(1 + g) * 16 + t] = tw[64 + 4 * t + 2 * g];
        for (int g = 0; g < 4; g++) t2h[(3 + g) * 16 + t] = tw[128 + 8 * t + 2 * g];
        for (int g = 0; g < 8; g++) t2h[(7 + g) * 16 + t] = tw[256 + 16 * t + 2 * g];
    }
    FCK(cudaMalloc(&f.t2h, sizeof(cplx) * 256));
    FCK(cudaMemcpy(f.t2h, t2h.data(), sizeof(cplx) * 256, cudaMemcpyHostToDevice));
    FCK(cudaMemcpyToSymbol(c_tw1h, tw1.data(), sizeof(cplx) * 32));
    {   // exp(-i*pi*j/16): shared with the N = 2048 kernel, which may not have been built for this context
        std::vector<cplx> e32(32);
        for (int j = 0; j < 32; j++) { const __float128 ang = pi * j / 16; e32[j] = make_double2((double)cosq(ang), (double)-sinq(ang)); }
        FCK(cudaMemcpyToSymbol(fastw::c_e32, e32.data(), sizeof(cplx) * 32));
    }
    const size_t polys = (size_t)p.n * 4 * p.l_gsw;
    FCK(cudaMalloc(&f.brk, polys * H * sizeof(cplx)));
    k_permute_brk_h<<<(unsigned)((polys * H + 255) / 256), 256, 0, stream>>>(brk_ref, f.brk, polys);
    FCK(cudaGetLastError());
    FCK(cudaFuncSetAttribute(k_cggi_w, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES_W));
    f.built = true;
    return 0;
}
static inline int fast32w_launch(FastKeys32W &f, const cplx *emono, const mktfhe_params &p, fastw32::Args a, cudaStream_t stream, int *launches,
                                 std::string &err) {
    using namespace fastw32;
    if (!f.built || !emono) { err = "FAST (N = 1024, half-warp transform) keys not built"; return -3; }
    a.brk = f.brk; a.tb = Tables{f.t2h, emono};
    a.n = p.n; a.l = p.l_gsw; a.logB = p.logB_gsw; a.lwe_words = (int)mktfhe_lwe_words(&p);
    const unsigned grid = (unsigned)((a.units + GATES_CTA - 1) / GATES_CTA);
    k_cggi_w<<<grid, CTA_W, SMEM_BYTES_W, stream>>>(a);
    if (launches) (*launches)++;
    FCK(cudaGetLastError());
    return 0;
}
