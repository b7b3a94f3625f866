// kernels_fast_w.cuh -- FAST-mode KMS phase 1, "one warp per transform" variant (production default for ELL = 1).
//
// What it computes: /root/reference/src/tfhe/bootstrapping.jl:389-443 (phase_1 of KMS), exactly like
// fast::k_phase1_tma (kernels_fast.cuh) -- same integer stages, same twist-free product-tree transform, same slot
// order -- with a different mapping onto the SM:
//
//   * A 1024-point transform belongs to ONE warp: 32 threads x 32 points, passes of 5 + 5 stages in registers and a
//     single warp-local transposition through shared memory in between (__syncwarp, no block barrier).  Against the
//     64-thread x 16-point mapping (4 + 4 + 2 stages, two exchanges, four named barriers per transform) this halves the
//     shared-memory exchange traffic -- the L1/shared data pipe was 59 % busy next to an FP64 pipe at 58 % -- removes
//     every barrier from the transforms and doubles the independent butterflies per stage (32 chains).
//   * A unit = one (gate, party, RLEV row) still owns one TMEM lane quadrant (64 KiB: RLWE accumulator + RGSW
//     accumulators), shared by TWO warps q and q + 4 (same quadrant, same scheduler).  Warp A owns the .b half of the RLWE
//     accumulator: it decomposes and transforms the l gadget digits of acc.b, inverse-transforms the .b sum and updates
//     acc.b.  Warp B does the same for acc.a.  Every spectrum must be multiplied into BOTH sums, so per digit A adds into
//     the .b sum first and the .a sum second while B goes .a first, .b second: the two warps always work on opposite
//     sums, and four mbarrier tokens per unit (b: A->B, B->A; a: B->A, A->B) fix the order of the additions
//     (.b: A0, B0, A1, B1, ...; .a: B0, A0, B1, A1, ...), so results are deterministic.  Nothing else couples the warps:
//     no block barrier inside the step loop (the 64-thread mapping needed 32 per step).
//   * Bootstrapping-key tiles arrive once per CTA through the same cp.async.bulk + mbarrier ring as before, in the order the
//     warps consume them (A_j.b, B_j.a, A_j.a, B_j.b); a tile is consumed by one warp per unit (one elected lane per warp
//     releases the slot).  All mbarrier waits suspend in hardware (try_wait with a time hint) instead of spinning.
//
// Index math of the transform is modelled and checked against the oracle in tools/models/fft32_model.py.
#pragma once
#include "kernels_fast.cuh"

namespace fastw {

using fast::H; using fast::N;
using fast::bf; using fast::bf_mi; using fast::bi; using fast::bi_mi;
using fast::tm_ld16; using fast::tm_st16; using fast::tm_wait_ld; using fast::tm_wait_st; using fast::tm_pin16; using fast::tm_st_c4;
using fast::mb_init; using fast::mb_expect_tx; using fast::mb_arrive; using fast::mb_wait; using fast::bulk_g2s; using fast::d2torus;

constexpr int WU = 4;                       // units per CTA = TMEM lane quadrants
constexpr int NCW = 2 * WU;                 // consumer warps
constexpr int CTA_W = NCW * 32 + 128;       // + producer warpgroup (setmaxnreg works on warpgroups)
constexpr int XBW = H + 32;                 // exchange buffer: element n at n + (n >> 5)

// ---- build switches (production values are the defaults; set with MKTFHE_NVCC_EXTRA, see mktfhe_b200/build.py) --------------------
//   W_INV_DIT, W_RING, W_FIRST_STORE, W_CREGS, W_DECOMP_I2F   measured alternatives, each described where it is used
//   KO_*                                                        knock-out TIMING builds: parts of the work removed, results wrong by
//                                                               construction (profiles/README_r2.md lists what each one really removes)
// W_INV_DIT: the inverse transform as a plain decimation-in-time network + untwist (tools/models/ifft_dit_model.py) instead of the
// forward product tree walked backwards: multiply-then-add butterflies (6 FMA-pipe instructions against 8), twiddles 1 and i in
// the first two stages, one complex multiplication per point for the untwist: 31.9k FP64 instructions per transform instead of
// 41.0k.  Its two tables (24 KiB) take the place of the fifth key tile of the ring.  Measured on B200 (KMS2party, 4096 gates):
// results within the same per-step tolerance (2^31.6), output noise 2^26.23 against 2^26.35, run time 211.2 ms against 208.9 ms --
// 5 % fewer FP64 instructions do not shorten the step (the kernel is not bound by the FP64 instruction count alone, see
// profiles/README_r2.md), so the product-tree inverse with the five-tile ring stays the default.
#ifndef W_INV_DIT
#define W_INV_DIT 0
#endif
// Key tiles (16 KiB polynomials) in flight: TWO rings, one per kind of consumer warp (tiles alternate A, B, A, B).  A slot must always
// serve the same kind of warp: in one ring of odd depth a slot alternated between the kinds, a warp saw only every second phase of
// its slot's `full` barrier, and a parity wait cannot tell "my phase" from "two phases earlier" -- a warp that ran ahead (dead units of
// a small batch, a skipped step) took the completion of the tile BEFORE the other kind's tile for its own when bulk copies completed
// out of order, released the slot twice, and the CTA deadlocked (KMS32party with one gate per call: one run in three).
#ifndef W_RING_A
#define W_RING_A 3
#endif
#ifndef W_RING_B
#define W_RING_B (W_INV_DIT ? 1 : 2)
#endif
constexpr int RING_A = W_RING_A, RING_B = W_RING_B, RINGW = RING_A + RING_B;
#ifndef W_FIRST_STORE
#define W_FIRST_STORE 0                     // see the sweep loop
#endif
#ifndef W_CREGS
#define W_CREGS 240                         // consumer registers after setmaxnreg: all the launch allocation allows (8 x 240 + 4 x 24 = 12 x 168);
#endif                                      // 232 left 24 B of spills after the key ring was split in two: 211.9 against 205.0 ms at KMS2party
#define W_STR2(x) #x
#define W_STR(x) W_STR2(x)
constexpr int W_LAUNCH_REGS = 168, W_CONSUMER_REGS = W_CREGS, W_PRODUCER_REGS = 24;
static_assert(NCW * W_CONSUMER_REGS + 4 * W_PRODUCER_REGS <= (CTA_W / 32) * W_LAUNCH_REGS, "setmaxnreg over-subscription");
constexpr int TABW = W_INV_DIT ? 512 + 512 + 514 : 512;     // [16][32] forward pass 2 | [16][32] inverse stages 5..9 | untwist [513] (+1 pad)
constexpr int NTOK = 4;                                     // tokens per unit
constexpr size_t SMEM_BYTES_W = ((size_t)NCW * XBW + TABW + (size_t)RINGW * H) * 16 + (2 * RINGW + NTOK * WU) * 8 + 16;
static_assert(SMEM_BYTES_W <= 232448, "shared memory budget");

__constant__ double2 c_tw1w[32];     // TW[1..31]: stages 1..5 (index 2^s + node), entry 0 unused
__constant__ double2 c_e32[32];      // exp(-i*pi*j/16)
__constant__ double2 c_w32i[16];     // exp(+2*pi*i*j/32): inverse stages 0..4

// TMEM columns of a lane: RGSW accumulators (32 complex each), RLWE accumulator (coefficient t + 32m at 4m, 4m+1 and
// coefficient t + 32m + H at 4m+2, 4m+3 of its half)
constexpr uint32_t TMW_TACC_B = 0, TMW_TACC_A = 128, TMW_ACC_B = 256, TMW_ACC_A = 384;

// ---- transform: x (layout A: element m = point t + 32m) <-> x (layout C: element e = slot 32t + e) -----------------
__device__ __forceinline__ void pass1_fwd(cplx (&x)[32]) {
#pragma unroll
    for (int s = 0; s < 5; s++) {
#pragma unroll
        for (int m = 0; m < 32; m++) {
            const int half = 16 >> s;
            if (m & half) continue;
            const int node = m >> (5 - s);
            if (node & 1) bf_mi(x[m], x[m + half], c_tw1w[(1 << s) + (node & ~1)]);
            else bf(x[m], x[m + half], c_tw1w[(1 << s) + node]);
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
            if (node & 1) bi_mi(x[m], x[m + half], c_tw1w[(1 << s) + (node & ~1)]);
            else bi(x[m], x[m + half], c_tw1w[(1 << s) + node]);
        }
    }
}
// stages 6..10 on the thread's 32 contiguous slots; tw = shared table [16][32] + lane:
//   row 0: TW[32 + t]; row 1: TW[64 + 2t]; rows 2+g: TW[128 + 4t + 2g]; rows 4+g: TW[256 + 8t + 2g]; rows 8+g: TW[512 + 16t + 2g]
// (A variant that swept the butterflies of one twiddle in source order -- all products with w.x, all with w.y, all completions --
// to feed the operand-reuse cache was measured: ptxas reschedules to the same mix, 21 % of the FP64 instructions keep three fresh
// register sources, run time unchanged within noise.)
__device__ __forceinline__ void pass2_fwd(cplx (&x)[32], const cplx *__restrict__ tw) {
    {
#ifdef KO_TW2CONST
        const cplx w = c_tw1w[1];
#else
        const cplx w = tw[0];
#endif
#pragma unroll
        for (int e = 0; e < 16; e++) bf(x[e], x[e + 16], w);
    }
#pragma unroll
    for (int s = 6; s < 10; s++) {
#pragma unroll
        for (int e = 0; e < 32; e++) {
            const int half = 16 >> (s - 5);
            if (e & half) continue;
            const int sub = e >> (10 - s);
#ifdef KO_TW2CONST
            const cplx w = c_tw1w[((1 << (s - 6)) + (sub >> 1))];
#else
            const cplx w = tw[((1 << (s - 6)) + (sub >> 1)) * 32];
#endif
            if (sub & 1) bf_mi(x[e], x[e + half], w); else bf(x[e], x[e + half], w);
        }
    }
}
__device__ __forceinline__ void pass2_inv(cplx (&x)[32], const cplx *__restrict__ tw) {
#pragma unroll
    for (int s = 9; s >= 6; s--) {
#pragma unroll
        for (int e = 0; e < 32; e++) {
            const int half = 16 >> (s - 5);
            if (e & half) continue;
            const int sub = e >> (10 - s);
#ifdef KO_TW2CONST
            const cplx w = c_tw1w[((1 << (s - 6)) + (sub >> 1))];
#else
            const cplx w = tw[((1 << (s - 6)) + (sub >> 1)) * 32];
#endif
            if (sub & 1) bi_mi(x[e], x[e + half], w); else bi(x[e], x[e + half], w);
        }
    }
    {
#ifdef KO_TW2CONST
        const cplx w = c_tw1w[1];
#else
        const cplx w = tw[0];
#endif
#pragma unroll
        for (int e = 0; e < 16; e++) bi(x[e], x[e + 16], w);
    }
}
// KO_NOFFT: no transforms at all; KO_BIDLE: role B (warps 4..7) computes no transforms -- how fast is one warp alone on its scheduler?
#if defined(KO_NOFFT)
#define KO_SKIP_FFT() return
#elif defined(KO_BIDLE)
#define KO_SKIP_FFT() if (threadIdx.x >= 128) return
#else
#define KO_SKIP_FFT()
#endif
__device__ __forceinline__ void fft_fwd(cplx (&x)[32], cplx *xb, const cplx *tw, int t) {
    KO_SKIP_FFT();
    pass1_fwd(x);
    __syncwarp();                                       // earlier readers of xb are done
#ifndef KO_XCHG                                         // KO_*: knock-out timing variants (wrong results), profiles/README_r2.md
#pragma unroll
    for (int m = 0; m < 32; m++) xb[t + 33 * m] = x[m];
    __syncwarp();
#pragma unroll
    for (int e = 0; e < 32; e++) x[e] = xb[33 * t + e];
#endif
    pass2_fwd(x, tw);
}
__device__ __forceinline__ void fft_inv(cplx (&x)[32], cplx *xb, const cplx *tw, int t) {
    KO_SKIP_FFT();
    pass2_inv(x, tw);
    __syncwarp();
#ifndef KO_XCHG
#pragma unroll
    for (int e = 0; e < 32; e++) xb[33 * t + e] = x[e];
    __syncwarp();
#pragma unroll
    for (int m = 0; m < 32; m++) x[m] = xb[t + 33 * m];
#endif
    pass1_inv(x);
}

// ---- inverse as decimation in time on the slot array (position p = slot index = bit-reversed frequency) --------------------------
// (a, b) -> (a + w*b, a - w*b) is fast::bf; the same with twiddle i*w, 1 and i:
__device__ __forceinline__ void bf_pi(cplx &a, cplx &b, const cplx w) {
    const double lx = fma(-w.x, b.y, fma(-w.y, b.x, a.x));      // a.x - Im(w*b)
    const double ly = fma(w.x, b.x, fma(-w.y, b.y, a.y));       // a.y + Re(w*b)
    b.x = fma(2.0, a.x, -lx); b.y = fma(2.0, a.y, -ly);
    a.x = lx; a.y = ly;
}
__device__ __forceinline__ void bf_one(cplx &a, cplx &b) {
    const cplx u = b;
    b = make_double2(a.x - u.x, a.y - u.y); a = make_double2(a.x + u.x, a.y + u.y);
}
__device__ __forceinline__ void bf_i(cplx &a, cplx &b) {
    const cplx u = b;
    b = make_double2(a.x + u.y, a.y - u.x); a = make_double2(a.x - u.y, a.y + u.x);
}
// stages 0..4 on the thread's slots 32t + e: twiddle exp(2*pi*i*(e mod 2^s)/2^(s+1)), compile-time
__device__ __forceinline__ void dit_pass_a(cplx (&x)[32]) {
#pragma unroll
    for (int s = 0; s < 5; s++) {
#pragma unroll
        for (int e = 0; e < 32; e++) {
            const int half = 1 << s;
            if (e & half) continue;
            const int idx = (e & (half - 1)) * (16 >> s);
            if (idx == 0) bf_one(x[e], x[e + half]);
            else if (idx == 8) bf_i(x[e], x[e + half]);
            else bf(x[e], x[e + half], c_w32i[idx]);
        }
    }
}
// stages 5..9 on elements t + 32m: twiddle exp(2*pi*i*(t + 32m')/2^(s+1)), m' = m mod 2^(s-5); tiw = shared table [16][32] + lane with
// rows s5: 0; s6: 1; s7: 2,3; s8: 4..7; s9: 8..15 holding m' < 2^(s-6); the siblings m' + 2^(s-6) carry the extra factor i
__device__ __forceinline__ void dit_pass_b(cplx (&x)[32], const cplx *__restrict__ tiw) {
#pragma unroll
    for (int s = 5; s < 10; s++) {
#pragma unroll
        for (int m = 0; m < 32; m++) {
            const int dm = 1 << (s - 5);
            if (m & dm) continue;
            const int mp = m & (dm - 1);
            const bool sib = s >= 6 && mp >= dm / 2;
            const int row = (s == 5 ? 0 : dm / 2) + (sib ? mp - dm / 2 : mp);
            const cplx w = tiw[row * 32];
            if (sib) bf_pi(x[m], x[m + dm], w); else bf(x[m], x[m + dm], w);
        }
    }
}
// slots (layout C) -> H * zeta^k * c_k in layout A (element m = coefficient pair k = t + 32m); the untwist is left to the caller
__device__ __forceinline__ void fft_inv_dit(cplx (&x)[32], cplx *xb, const cplx *tiw, int t) {
    dit_pass_a(x);
    __syncwarp();
#pragma unroll
    for (int e = 0; e < 32; e++) xb[33 * t + e] = x[e];
    __syncwarp();
#pragma unroll
    for (int m = 0; m < 32; m++) x[m] = xb[t + 33 * m];
    dit_pass_b(x, tiw);
}
// untwist factor exp(i*pi*k/2048) of element m, k = t + 32m, from the half table T[0..512]: T[k] = (T[1024-k].y, T[1024-k].x)
__device__ __forceinline__ cplx untwist(const cplx *__restrict__ T, int t, int m) {
    if (m < 16) return T[t + 32 * m];
    const cplx u = T[1024 - 32 * m - t];
    return make_double2(u.y, u.x);
}

__device__ __forceinline__ void nb_sync(int id) { asm volatile("bar.sync %0, 64;" :: "r"(id) : "memory"); }
__device__ __forceinline__ void nb_arrive(int id) { asm volatile("bar.arrive %0, 64;" :: "r"(id) : "memory"); }
__device__ __forceinline__ void tm_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tm_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void mb_wait_sleep(uint64_t *bar, uint32_t parity) { fast::mb_wait_suspend(bar, parity); }
// the same on 32-bit shared-window addresses computed once per kernel: the generic -> shared conversion at every call site cost
// an S2UR + ULEA with a scoreboard wait each (1.5 % of the consumer warps' time in the v3 profile)
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
template <int SUSPEND> __device__ __forceinline__ void mbs_wait(uint32_t bar, uint32_t parity) {
    if (!SUSPEND) { fast::mbs_spin(bar, parity); return; }
    uint32_t spins = 0, ok = 0;
    do {
        asm volatile("{\n.reg .pred p;\n"
                     "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n"
                     "selp.u32 %0, 1, 0, p;\n}" : "=r"(ok) : "r"(bar), "r"(parity), "r"(20000u) : "memory");
        if (!ok && ++spins > (1u << 22)) __trap();
    } while (!ok);
}
__device__ __forceinline__ void mbs_arrive(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(bar) : "memory"); }
// token = mbarrier with one arrival per phase; the waiter keeps the phase parity
template <int SUSPEND> struct Token {
    uint32_t bar;          // shared-window address
    uint32_t par;
#ifdef KO_TOKENS
    __device__ __forceinline__ void wait() {}
#else
    __device__ __forceinline__ void wait() { mbs_wait<SUSPEND>(bar, par); par ^= 1; }
#endif
};
__device__ __forceinline__ void token_pass(uint32_t bar, int t) {     // all lanes' TMEM stores are complete (tcgen05.wait::st)
    __syncwarp();
#ifndef KO_TOKENS
    if (t == 0) mbs_arrive(bar);
#endif
}
// ring position of a consumer warp: tile -> (slot, phase parity), advanced without divisions
struct RingPos {                             // role w walks the slots [0, RING_A) (A) or [RING_A, RINGW) (B) of its own ring
    uint32_t slot, par;
    __device__ __forceinline__ void advance(int w) { if (++slot == (w ? (uint32_t)RINGW : (uint32_t)RING_A)) { slot = w ? (uint32_t)RING_A : 0u; par ^= 1; } }
};

__device__ __forceinline__ cplx unpack_c(const uint32_t (&v)[16], int i) {
    return make_double2(__hiloint2double((int)v[4 * i + 1], (int)v[4 * i]), __hiloint2double((int)v[4 * i + 3], (int)v[4 * i + 2]));
}

// Producer waits are suspended in hardware (measured: 209 ms against 215 ms with a spinning producer at KMS2party, 4096 gates);
// consumer waits spin (suspended consumer waits measured the same within noise).
__global__ void __launch_bounds__(CTA_W, 1) k_phase1_w(const fast::Args a) {
    constexpr int SUSPEND = 0;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int tid = threadIdx.x, warp = tid >> 5, t = tid & 31;
    cplx *xb_all = reinterpret_cast<cplx *>(smem_raw);
    cplx *tw2s = xb_all + (size_t)NCW * XBW;
    cplx *ring = tw2s + TABW;
    const cplx *tiws = tw2s + 512, *twist = tw2s + 1024;          // W_INV_DIT only
    uint64_t *full = reinterpret_cast<uint64_t *>(ring + (size_t)RINGW * H), *empty = full + RINGW;
    uint64_t *tokens = empty + RINGW;                               // [WU][4]: b A->B, b B->A, a B->A, a A->B
    uint32_t *tm_base_s = reinterpret_cast<uint32_t *>(tokens + NTOK * WU);
    for (int i = tid; i < TABW; i += CTA_W) tw2s[i] = a.tb.t2w[i];
    if (tid == 0) {
        for (int s = 0; s < RINGW; s++) { mb_init(&full[s], 1); mb_init(&empty[s], WU); }
        for (int s = 0; s < NTOK * WU; s++) mb_init(&tokens[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" :: "r"((uint32_t)__cvta_generic_to_shared(tm_base_s)));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    tm_fence_before();
    __syncthreads();
    tm_fence_after();

    // ---- CTA -> (party, units): party 0 has one RLEV row per gate, the others l_lev (bootstrapping.jl:400)
    int party, rows;
    size_t u0, gates;
    if (!a.step_mode) {
        gates = a.units / a.R;
        const size_t ctas0 = (gates + WU - 1) / WU, ctasp = (gates * a.l_lev + WU - 1) / WU;
        if (blockIdx.x < ctas0) { party = 0; rows = 1; u0 = (size_t)blockIdx.x * WU; }
        else { party = 1 + (int)((blockIdx.x - ctas0) / ctasp); rows = a.l_lev; u0 = ((blockIdx.x - ctas0) % ctasp) * WU; }
    } else { gates = a.units; party = a.step_party; rows = 1; u0 = (size_t)blockIdx.x * WU; }
    const int l = a.l;
    const size_t per_idx = (size_t)4 * l * H;
    const int nsteps = a.step_mode ? 1 : a.n;
    const cplx *brk = a.brk[party];
    const uint32_t ntiles = (uint32_t)nsteps * 2 * l * 2;

    if (warp >= NCW) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 24;");
        // ---- producer: 16 KiB polynomials in thread order [e][t], in consumption order: per step and j < l the four tiles
        //      A_j.b = (digit j, .b), B_j.a = (digit l + j, .a), A_j.a = (digit j, .a), B_j.b = (digit l + j, .b)
        if (tid == NCW * 32) {
            uint32_t step = 0, j = 0, r = 0;
            // per ring: next slot, parity of the PREVIOUS phase of that slot's empty barrier; the first lap needs no wait
            uint32_t sa = 0, pa = 1, sb = RING_A, pb = 1;
            for (uint32_t n = 0; n < ntiles; n++) {
                const bool kb = r & 1;                                 // tile for the B warps
                const uint32_t slot = kb ? sb : sa;
                if (n >= 2 * (kb ? RING_B : RING_A)) mb_wait_sleep(&empty[slot], kb ? pb : pa);
                const uint32_t dg = (r & 1) ? (uint32_t)l + j : j, comp = (r == 1 || r == 2) ? 1u : 0u;
                const int idx = a.step_mode ? a.step_idx : (int)step;
                mb_expect_tx(&full[slot], H * 16);
                bulk_g2s(ring + (size_t)slot * H, brk + (size_t)idx * per_idx + (size_t)(dg * 2 + comp) * H, H * 16, &full[slot]);
                if (kb) { if (++sb == RINGW) { sb = RING_A; pb ^= 1; } }
                else { if (++sa == RING_A) { sa = 0; pa ^= 1; } }
                if (++r == 4) { r = 0; if (++j == (uint32_t)l) { j = 0; step++; } }
            }
        }
    } else {
        asm volatile("setmaxnreg.inc.sync.aligned.u32 " W_STR(W_CREGS) ";");
        // ---- consumers: unit q = TMEM lane quadrant, role w (0 = A: the .b half, 1 = B: the .a half)
        const int q = warp & 3, w = warp >> 2;
        const uint32_t tk = smem_u32(tokens + 4 * q);
        const uint32_t full_s = smem_u32(full), empty_s = smem_u32(empty);
        // sum I add into FIRST (own) and SECOND (other): tokens I wait on / pass on
        //   own sum:   wait other's second-pass token of the previous digit, pass my first-pass token
        //   other sum: wait other's first-pass token of this digit, pass my second-pass token
        Token<SUSPEND> wait_own{tk + 8u * (w == 0 ? 1 : 3), 0u}, wait_oth{tk + 8u * (w == 0 ? 2 : 0), 0u};
        const uint32_t pass_own = tk + 8u * (w == 0 ? 0 : 2), pass_oth = tk + 8u * (w == 0 ? 3 : 1);
        const uint32_t tm = *tm_base_s + ((uint32_t)(32 * q) << 16);
        cplx *xb = xb_all + (size_t)warp * XBW;
        const cplx *tw = tw2s + t;
        const size_t up = u0 + q;                                  // unit index inside the party
        const bool live = up < gates * (size_t)rows;
        const int gate = live ? (int)(up / rows) : 0, row = live ? (int)(up % rows) : 0;
        const size_t unit_out = a.step_mode ? up : (size_t)gate * a.R + (party == 0 ? 0 : 1 + (size_t)(party - 1) * a.l_lev + row);
        const uint32_t tm_acc = tm + (w == 0 ? TMW_ACC_B : TMW_ACC_A), tm_tacc = tm + (w == 0 ? TMW_TACC_B : TMW_TACC_A);

        uint64_t cadd = (uint64_t)1 << (64 - l * a.logB - 1);
        for (int j = 0; j < l; j++) cadd += (uint64_t)1 << (64 - l * a.logB + j * a.logB + a.logB - 1);
        if (live) {                                                // each warp initialises its half of the RLWE accumulator (+ cadd) and of the sums
            if (!W_FIRST_STORE) {
                uint32_t z[16];
#pragma unroll
                for (int i = 0; i < 16; i++) z[i] = 0u;
#pragma unroll
                for (int c = 0; c < 8; c++) tm_st16(tm_tacc + 16 * c, z);
            }
            if (!a.step_mode) {
                uint32_t z[16];
#pragma unroll
                for (int i = 0; i < 16; i += 2) { z[i] = (uint32_t)cadd; z[i + 1] = (uint32_t)(cadd >> 32); }
                const uint64_t gv = cadd + ((t == 0 && w == 0) ? (uint64_t)1 << (64 - (row + 1) * a.logB_lev) : 0);   // bootstrapping.jl:402-408
#pragma unroll
                for (int c = 0; c < 8; c++) {
                    z[0] = c == 0 ? (uint32_t)gv : (uint32_t)cadd;
                    z[1] = c == 0 ? (uint32_t)(gv >> 32) : (uint32_t)(cadd >> 32);
                    tm_st16(tm_acc + 16 * c, z);
                }
            } else {
                const uint64_t *src = a.acc_io + up * 2 * N + (size_t)w * N;
#pragma unroll
                for (int c = 0; c < 8; c++) {
                    uint32_t v[16];
#pragma unroll
                    for (int i = 0; i < 4; i++) {
                        const uint64_t c0 = src[t + 32 * (4 * c + i)] + cadd, c1 = src[t + 32 * (4 * c + i) + H] + cadd;
                        v[4 * i] = (uint32_t)c0; v[4 * i + 1] = (uint32_t)(c0 >> 32); v[4 * i + 2] = (uint32_t)c1; v[4 * i + 3] = (uint32_t)(c1 >> 32);
                    }
                    tm_st16(tm_acc + 16 * c, v);
                }
            }
            tm_wait_st();
        }
        const int logB = a.logB;
        const int bit = 64 - l * logB;
        // divbits rounding + balanced-digit carry chain in one 64-bit add (see kernels_fast.cuh).  The accumulator is KEPT with
        // this constant added (acc + cadd mod 2^64): digits are then plain bit fields of the stored words; the constant comes
        // off again when the accumulator leaves the kernel.
        const uint32_t mask = (1u << logB) - 1;
        const double dbias = 4503599627370496.0 + (double)(1 << (logB - 1));
        const int hB = 1 << (logB - 1);
        const uint32_t *at_src = a.step_mode ? a.tilde + up : a.tilde + (size_t)gate * a.lwe_words + 1 + (size_t)party * a.n;
        const int brv5t = (int)(__brev((unsigned)t) >> 27);
        // this warp consumes every second tile of the producer's sequence, starting at tile w, from its own ring
        RingPos rp{w == 0 ? 0u : (uint32_t)RING_A, 0u};

        for (int step = 0; step < nsteps; step++) {
            const uint32_t at = live ? at_src[a.step_mode ? 0 : step] : 0u;
            if (at == 0) {                                          // :413 / dead unit: keep the ring moving, compute nothing
                for (int j = 0; j < 2 * l; j++) {
                    mbs_wait<SUSPEND>(full_s + 8u * rp.slot, rp.par);
                    __syncwarp();
                    if (t == 0) mbs_arrive(empty_s + 8u * rp.slot);
                    rp.advance(w);
                }
                continue;
            }
            const cplx m1 = __ldg(&a.tb.emono[((4 * brv5t + 1) * at) & 4095]);

            for (int j = 0; j < l; j++) {
                const int sh = bit + (l - 1 - j) * logB;             // digit j (0 = most significant) of my half of the accumulator
                cplx x[32];
#ifdef KO_NODECOMP
#pragma unroll
                for (int i = 0; i < 32; i++) x[i] = make_double2(1e-3 * t + i, 1.0 + j);
#else
#pragma unroll
                for (int hb = 0; hb < 2; hb++) {                     // gadget digit of 64 coefficients -> 32 complex points
                    uint32_t v[4][16];
#pragma unroll
                    for (int c = 0; c < 4; c++) tm_ld16(tm_acc + 64 * hb + 16 * c, v[c]);
                    tm_wait_ld();
#pragma unroll
                    for (int c = 0; c < 4; c++) {
                        tm_pin16(v[c]);
#pragma unroll
                        for (int i = 0; i < 4; i++) {
                            const uint64_t v0 = ((uint64_t)v[c][4 * i + 1] << 32) | v[c][4 * i];
                            const uint64_t v1 = ((uint64_t)v[c][4 * i + 3] << 32) | v[c][4 * i + 2];
                            const uint32_t f0 = (uint32_t)(v0 >> sh) & mask, f1 = (uint32_t)(v1 >> sh) & mask;
                            // signed(d_j) - im*signed(d_{j+H}); 2^52 + field is exact in the double's mantissa
#ifdef W_DECOMP_I2F
                            // conversion unit instead of the FP64 pipe (one I2F against one DADD per coefficient)
                            x[16 * hb + 4 * c + i] = make_double2(__int2double_rn((int)f0 - hB), __int2double_rn(hB - (int)f1));
#else
                            x[16 * hb + 4 * c + i] = make_double2(__hiloint2double(0x43300000, (int)f0) - dbias, dbias - __hiloint2double(0x43300000, (int)f1));
#endif
                        }
                    }
                }
#endif
                fft_fwd(x, xb, tw, t);
                // RGSW sums += spectrum x key: my own sum first, the other warp's sum second (the other warp goes the other way
                // round); the accumulator chunk after the one being updated is already in flight
#pragma unroll
                for (int ps = 0; ps < 2; ps++) {
                    const uint32_t tmz = ps == 0 ? tm_tacc : tm + (w == 0 ? TMW_TACC_A : TMW_TACC_B);
                    mbs_wait<SUSPEND>(full_s + 8u * rp.slot, rp.par);
                    const cplx *kp = ring + (size_t)rp.slot * H + t;
                    if (ps == 0) { if (j > 0) wait_own.wait(); } else wait_oth.wait();
                    tm_fence_after();
                    // W_FIRST_STORE: the first addition of a step into a sum is its owner's (ps == 0, j == 0) and follows the owner's own
                    // read of the previous step's sum in program order, so it could store x * key instead of accumulating and the sums
                    // would never be cleared.  Measured slower (218.3 against 211.2 ms: the second code path costs registers), off.
                    const bool first = W_FIRST_STORE && ps == 0 && j == 0;
                    uint32_t v[2][16];
#if defined(KO_NOMAC)
                    constexpr int NCH = 0;
#else
                    constexpr int NCH = 8;
#endif
#ifdef KO_MACTM
#pragma unroll
                    for (int i = 0; i < 16; i++) v[0][i] = v[1][i] = 0x3ff00000u ^ (uint32_t)(i * t);
#else
#ifndef KO_MACNOLD
                    if (!first && NCH) tm_ld16(tmz, v[0]);
#endif
#endif
#pragma unroll
                    for (int c = 0; c < NCH; c++) {
                        cplx kc[4], z[4];
#pragma unroll
#if defined(KO_KEYLDS)
                        for (int i = 0; i < 4; i++) kc[i] = c_tw1w[(4 * c + i) & 31];
#elif defined(KO_KEYHALF)                       // half of the key reads (every value used twice)
                        for (int i = 0; i < 4; i++) kc[i] = kp[(4 * c + (i & 1)) * 32];
#else
                        for (int i = 0; i < 4; i++) kc[i] = kp[(4 * c + i) * 32];
#endif
                        if (first) {
#pragma unroll
                            for (int i = 0; i < 4; i++) z[i] = cmul_f(x[4 * c + i], kc[i]);
                        } else {
#if defined(KO_MACNOLD)                         // no TMEM loads in the sweep (stores stay)
                            if (c == 0) {
#pragma unroll
                                for (int i = 0; i < 16; i++) v[0][i] = v[1][i] = 0x3ff00000u ^ (uint32_t)(i * t);
                            }
#elif !defined(KO_MACTM)
                            tm_wait_ld();
                            tm_pin16(v[c & 1]);
                            if (c < 7) tm_ld16(tmz + 16 * (c + 1), v[(c + 1) & 1]);
#endif
#pragma unroll
#ifdef KO_MACNOFMA                              // loads and stores stay, no arithmetic
                            for (int i = 0; i < 4; i++) z[i] = unpack_c(v[c & 1], i);
#else
                            for (int i = 0; i < 4; i++) z[i] = cmac_f(unpack_c(v[c & 1], i), x[4 * c + i], kc[i]);
#endif
                        }
#ifdef KO_MACTM
#pragma unroll
                        for (int i = 0; i < 4; i++) { x[4 * c + i].x += z[i].x * 1e-300; x[4 * c + i].y += z[i].y * 1e-300; }   // keep the arithmetic alive
#elif defined(KO_MACNOST)                       // no TMEM stores in the sweep (loads and arithmetic stay: a store that never happens keeps them alive)
#pragma unroll
                        for (int i = 0; i < 4; i++) if (z[i].x == 1.2345e-300) a.lev_out[i] = z[i];
#else
                        tm_st_c4(tmz + 16 * c, z);
#endif
                    }
                    tm_wait_st();
                    tm_fence_before();
                    token_pass(ps == 0 ? pass_own : pass_oth, t);      // also orders every lane's key reads before the release below
                    if (t == 0) mbs_arrive(empty_s + 8u * rp.slot);     // this key tile is done
                    rp.advance(w);
                }
            }
            wait_own.wait();                                        // the other warp's last addition into my sum
            tm_fence_after();
            // this warp's output: (x (X^a - 1)/H) -> inverse transform -> round -> acc +=
#ifndef KO_NOEND
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
                if (!W_FIRST_STORE) {   // clear the sums for the next step (every product accumulates, also the first)
                    uint32_t z[16];
#pragma unroll
                    for (int i = 0; i < 16; i++) z[i] = 0u;
#pragma unroll
                    for (int c = 0; c < 8; c++) tm_st16(tm_tacc + 16 * c, z);
                }
                // slot 32t + e evaluates at exp(-i*pi*(4*brv10(n)+1)/N), brv10(n) = 32*brv5(e) + brv5(t)
#pragma unroll
                for (int e = 0; e < 32; e++) {
                    const int b5 = ((e & 1) << 4) | ((e & 2) << 2) | (e & 4) | ((e & 8) >> 2) | ((e & 16) >> 4);
                    cplx mo = cmul_f(m1, c_e32[(at * b5) & 31]);
                    mo.x -= 1.0 / H;
                    const double pr = mo.y * y[e].y, pi = mo.y * y[e].x;               // first source mo.y twice, then mo.x twice
                    y[e] = make_double2(fma(mo.x, y[e].x, -pr), fma(mo.x, y[e].y, pi));
                }
#if W_INV_DIT
                fft_inv_dit(y, xb, tiws + t, t);
#else
                fft_inv(y, xb, tw, t);
#endif
                uint32_t v[2][16];
                tm_ld16(tm_acc, v[0]);
#pragma unroll
                for (int c = 0; c < 8; c++) {
                    tm_wait_ld();
                    tm_pin16(v[c & 1]);
                    if (c < 7) tm_ld16(tm_acc + 16 * (c + 1), v[(c + 1) & 1]);      // next chunk in flight during the rounding
#pragma unroll
                    for (int i = 0; i < 4; i++) {
                        const int m = 4 * c + i;
#if W_INV_DIT
                        const cplx ym = cmul_f(y[m], untwist(twist, t, m));
#else
                        const cplx ym = y[m];
#endif
                        const uint64_t w0 = (((uint64_t)v[c & 1][4 * i + 1] << 32) | v[c & 1][4 * i]) + d2torus(ym.x);
                        const uint64_t w1 = (((uint64_t)v[c & 1][4 * i + 3] << 32) | v[c & 1][4 * i + 2]) + d2torus(-ym.y);
                        v[c & 1][4 * i] = (uint32_t)w0; v[c & 1][4 * i + 1] = (uint32_t)(w0 >> 32);
                        v[c & 1][4 * i + 2] = (uint32_t)w1; v[c & 1][4 * i + 3] = (uint32_t)(w1 >> 32);
                    }
                    tm_st16(tm_acc + 16 * c, v[c & 1]);
                }
                tm_wait_st();
            }
#endif
        }

        if (live) {
            if (!a.step_mode) {            // fftto!(tacc, acc): bootstrapping.jl:441 -- full-width coefficients, warp A .b, warp B .a
                cplx *out = a.lev_out + unit_out * 2 * H + (size_t)w * H;
                cplx x[32];
#pragma unroll
                for (int hb = 0; hb < 2; hb++) {
                    uint32_t v[4][16];
#pragma unroll
                    for (int c = 0; c < 4; c++) tm_ld16(tm_acc + 64 * hb + 16 * c, v[c]);
                    tm_wait_ld();
#pragma unroll
                    for (int c = 0; c < 4; c++) {
                        tm_pin16(v[c]);
#pragma unroll
                        for (int i = 0; i < 4; i++) {
                            const uint64_t v0 = (((uint64_t)v[c][4 * i + 1] << 32) | v[c][4 * i]) - cadd, v1 = (((uint64_t)v[c][4 * i + 3] << 32) | v[c][4 * i + 2]) - cadd;
                            x[16 * hb + 4 * c + i] = make_double2(__ll2double_rn((long long)v0), __ll2double_rn((long long)((uint64_t)0 - v1)));
                        }
                    }
                }
                fft_fwd(x, xb, tw, t);
                if (a.lev_fast) {          // for k_phase2: its thread order [n & 15][n >> 4] and the 1/H of its inverse transforms (exact)
#pragma unroll
                    for (int e = 0; e < 32; e++) out[(e & 15) * 64 + 2 * t + (e >> 4)] = make_double2(x[e].x * (1.0 / H), x[e].y * (1.0 / H));
                } else {
#pragma unroll
                    for (int e = 0; e < 32; e++) out[32 * t + e] = x[e];
                }
            } else {
                uint64_t *dst = a.acc_io + up * 2 * N + (size_t)w * N;
#pragma unroll
                for (int c = 0; c < 8; c++) {
                    uint32_t v[16];
                    tm_ld16(tm_acc + 16 * c, v);
                    tm_wait_ld();
                    tm_pin16(v);
#pragma unroll
                    for (int i = 0; i < 4; i++) {
                        dst[t + 32 * (4 * c + i)] = (((uint64_t)v[4 * i + 1] << 32) | v[4 * i]) - cadd;
                        dst[t + 32 * (4 * c + i) + H] = (((uint64_t)v[4 * i + 3] << 32) | v[4 * i + 2]) - cadd;
                    }
                }
            }
        }
    }
    tm_fence_before();
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" :: "r"(*tm_base_s));
}

// reference slot order [poly][32t + e] -> thread order [poly][e][t]
__global__ void k_permute_brk_w(const cplx *__restrict__ in, cplx *__restrict__ out, size_t polys) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= polys * H) return;
    const size_t p = i / H;
    const int r = (int)(i % H), e = r / 32, t = r % 32;
    out[i] = in[p * H + 32 * t + e];
}

}  // namespace fastw
