// keyswitch.cuh -- integer stages: gate linear part, modulus switch, sample extraction + key switch.
// All arithmetic is UInt32 / UInt64 wraparound and must be bit-exact against the reference.
//
//   k_gate_prep      gate.jl:1-52 (linear part) + bootstrapping.jl:8-9 (modswitch by divbits)
//   k_keyswitch      keyswitch! CGGI bootstrapping.jl:81-109, LMSS :170-229, CCS :333-364,
//                    KMS :564-594, KMS_block :664-695
#pragma once
#include "common.cuh"

// lin = gate linear combination (or copy when op < 0); tilde = divbits(lin, 32 - log2(N) - 1).
// Circuit levels pass per-gate opcodes (`ops`) and gather the operands from a wire table (`idx1`, `idx2`: row
// indices into in1 / in2); all three are null for a plain batch.
__global__ void k_gate_prep(const uint32_t *in1, const uint32_t *in2, uint32_t *lin, uint32_t *tilde,
                            int op, int lwe_words, size_t total, int shift,
                            const int32_t *ops = nullptr, const int32_t *idx1 = nullptr, const int32_t *idx2 = nullptr) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const size_t g = i / lwe_words, w = i % lwe_words;
    const bool is_b = w == 0;
    if (ops) op = ops[g];
    const size_t i1 = idx1 ? (size_t)idx1[g] * lwe_words + w : i;
    uint32_t v;
    if (op < 0) v = in1[i1];
    else {
        const size_t i2 = idx2 ? (size_t)idx2[g] * lwe_words + w : i;
        const uint32_t s = in1[i1] + in2[i2];
        uint32_t cst;
        switch (op) {                                    // Julia: `T(c) << 29 - x - y` = (c << 29) - x - y
        case 0:  cst = 1u << 29; v = 0u - s; break;      // NAND  gate.jl:2-4
        case 1:  cst = 7u << 29; v = s; break;           // AND   :11-13
        case 2:  cst = 1u << 29; v = s; break;           // OR    :20-22
        case 3:  cst = 1u << 30; v = 2u * s; break;      // XOR   :29-31
        case 4:  cst = 3u << 30; v = 0u - 2u * s; break; // XNOR  :38-40
        default: cst = 7u << 29; v = 0u - s; break;      // NOR   :47-49
        }
        if (is_b) v += cst;
    }
    if (lin) lin[i] = v;
    if (tilde) tilde[i] = divbits<uint32_t>(v, shift);
}

// Wire-table plumbing of a circuit level: rows of `src` (one per gate) go to wires[dst[g]]; with negate, rows are
// gathered from wires[from[g]] and negated on the way (NOT!, gate.jl:55-58: no bootstrap).
__global__ void k_wire_scatter(const uint32_t *src, uint32_t *wires, const int32_t *from, const int32_t *dst,
                               int lwe_words, size_t total, int negate) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const size_t g = i / lwe_words, w = i % lwe_words;
    const uint32_t v = negate ? 0u - wires[(size_t)from[g] * lwe_words + w] : src[i];
    wires[(size_t)dst[g] * lwe_words + w] = v;
}

struct KsArgs {
    const void *acc;              // [B][(k+1)][N] torus
    const uint32_t *const *ksk;   // [k]: [N][Dk][f][n+1]
    uint32_t *out;                // [B][1 + n*k]
    int N, n, k, f, logD, Dk, bits64, block;
    int rowp;                     // padded row stride in words (multiple of 4: rows are 16-byte aligned for bulk copies)
};

// One CTA per gate; parties in sequence; thread tid owns LWE columns tid, tid+256, tid+512 (column 0 = b).
// Digits of all N extracted coefficients are staged in shared memory first (balanced digits need the
// carry chain), then the ksk rows they select are streamed with coalesced loads.
__global__ void __launch_bounds__(MK_THREADS) k_keyswitch(const KsArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    int8_t *dig = reinterpret_cast<int8_t *>(smem_raw);                 // [N][f]
    uint32_t *copy = reinterpret_cast<uint32_t *>(smem_raw + (size_t)a.N * a.f);   // [n] (block schemes)
    const int tid = threadIdx.x, gate = blockIdx.x;
    const int N = a.N, n = a.n, f = a.f, row = n + 1;
    constexpr int MAXC = 3;                                             // ceil((n+1)/256), n <= 767
    const uint32_t *acc32 = reinterpret_cast<const uint32_t *>(a.acc) + (size_t)gate * (a.k + 1) * N;
    const uint64_t *acc64 = reinterpret_cast<const uint64_t *>(a.acc) + (size_t)gate * (a.k + 1) * N;
    auto A = [&](int comp, int c) -> uint32_t {                         // `T(x >> bitdiff)`: truncation (:569,575,583)
        return a.bits64 ? (uint32_t)(acc64[(size_t)comp * N + c] >> 32) : acc32[(size_t)comp * N + c];
    };
    uint32_t *out = a.out + (size_t)gate * (1 + (size_t)n * a.k);
    uint32_t bsum = A(0, 0);                                            // res.b = acc.b[0]
    const uint32_t mask = (1u << a.logD) - 1, halfD = 1u << (a.logD - 1);

    for (int p = 0; p < a.k; p++) {
        __syncthreads();
        for (int c = tid; c < N; c += MK_THREADS) {
            // sample extraction of coefficient 0: a'_1 = a[0], a'_j = -a[N-j+1]  (:91-98, :575-583)
            const uint32_t val = c == 0 ? A(1 + p, 0) : 0u - A(1 + p, N - c);
            if (a.block && c < n) { copy[c] = val; continue; }          // :678-681 / :177-190
            uint32_t ai = divbits<uint32_t>(val, 32 - f * a.logD);
            if (!a.block) {                                             // unbalanceddecompto! gsw.jl:34-40
                for (int i = f - 1; i >= 0; i--) { dig[c * f + i] = (int8_t)(ai & mask); ai >>= a.logD; }
            } else {                                                    // decompto! gsw.jl:42-52
                for (int i = f - 1; i >= 1; i--) {
                    const uint32_t d = ai & mask;
                    ai >>= a.logD;
                    ai += d >> (a.logD - 1);
                    dig[c * f + i] = (int8_t)((int32_t)d - (int32_t)((d & halfD) << 1));
                }
                const uint32_t d = ai & mask;
                dig[c * f] = (int8_t)((int32_t)d - (int32_t)((d & halfD) << 1));
            }
        }
        __syncthreads();
        uint32_t sum[MAXC] = {0u, 0u, 0u};
        const uint32_t *ksk = a.ksk[p];
        for (int c = a.block ? n : 0; c < N; c++) {
            const uint32_t *base = ksk + (size_t)c * a.Dk * f * a.rowp;
#pragma unroll 4
            for (int lv = 0; lv < f; lv++) {
                const int d = dig[c * f + lv];
                if (d == 0) continue;
                const uint32_t *r = base + ((size_t)((d > 0 ? d : -d) - 1) * f + lv) * a.rowp;
#pragma unroll
                for (int q = 0; q < MAXC; q++) {
                    const int col = tid + q * MK_THREADS;
                    if (col < row) { const uint32_t x = __ldg(&r[col]); sum[q] = d > 0 ? sum[q] + x : sum[q] - x; }
                }
            }
        }
#pragma unroll
        for (int q = 0; q < MAXC; q++) {
            const int col = tid + q * MK_THREADS;
            if (col == 0) bsum += sum[q];
            else if (col < row) out[1 + (size_t)p * n + (col - 1)] = sum[q] + ((a.block && col - 1 < n) ? copy[col - 1] : 0u);
        }
    }
    if (tid == 0) out[0] = bsum;
}

static inline size_t keyswitch_smem_bytes(int N, int f, int n) { return (size_t)N * f + (size_t)n * 4 + 16; }

// ---------------------------------------------------------------------------------------------------
// Tiled key switch (production): one CTA = G = 16 gates x one party; warp w owns gates 2w, 2w+1 and lane i owns LWE
// columns i, i+32, ... of their partial sums (registers).  All gates walk (c, level) together: the Dk candidate ksk
// rows of a (c, level) are fetched from L2 ONCE per CTA by TMA bulk copies (cp.async.bulk + mbarrier, 4-stage ring in
// shared memory, issued by one thread) and every gate adds the row its digit selects with conflict-free
// shared-memory reads -- a warp-uniform choice, so there is no predicated-off work.  Digits of the tile are staged as
// 2-bit fields (f*logD = 16 bits per coefficient).  Integer adds commute: per-party partial sums of b are combined
// with atomicAdd into a zeroed output.
constexpr int KS_G = 32, KS_GW = 4, KS_J = 6, KS_STAGES = 6;     // lane owns words 4*lane + 128*j: 6 * 128 = 768 >= n + 1 for every set in params.jl

__device__ __forceinline__ void mbar_init(uint64_t *bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"((uint32_t)__cvta_generic_to_shared(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"((uint32_t)__cvta_generic_to_shared(bar)), "r"(bytes) : "memory");
}
// tight try_wait loop; bounded (trap after 2^28 failed polls) in -DMKTFHE_DEBUG_SPIN builds, see kernels_fast.cuh
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
#ifdef MKTFHE_DEBUG_SPIN
    asm volatile("{\n.reg .pred p, q;\n.reg .u32 cnt;\nmov.u32 cnt, 0;\nWAIT_%=:\n"
                 "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
                 "@p bra DONE_%=;\n"
                 "add.u32 cnt, cnt, 1;\nsetp.gt.u32 q, cnt, 268435456;\n@q trap;\n"
                 "bra WAIT_%=;\nDONE_%=:\n}" :: "r"((uint32_t)__cvta_generic_to_shared(bar)), "r"(parity) : "memory");
#else
    asm volatile("{\n.reg .pred p;\nWAIT_%=:\n"
                 "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
                 "@p bra DONE_%=;\nbra WAIT_%=;\nDONE_%=:\n}" :: "r"((uint32_t)__cvta_generic_to_shared(bar)), "r"(parity) : "memory");
#endif
}
__device__ __forceinline__ void tma_bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"((uint32_t)__cvta_generic_to_shared(dst)), "l"(src), "r"(bytes), "r"((uint32_t)__cvta_generic_to_shared(bar)) : "memory");
}

template <bool BLOCK>
__global__ void __launch_bounds__(MK_THREADS) k_keyswitch_tiled(const KsArgs a, int batch, int split) {
    constexpr int G = KS_G, DK = BLOCK ? 2 : 3, S = KS_STAGES;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    // the coefficient range [c_all, N) is split over blockIdx.z; each CTA stages only its own coefficients' digits
    const int c_all = BLOCK ? a.n : 0, c_per = (a.N - c_all + split - 1) / split;
    const int c_begin = c_all + (int)blockIdx.z * c_per, c_end = min(a.N, c_begin + c_per);
    uint16_t *dig = reinterpret_cast<uint16_t *>(smem_raw);                                  // [c_per][G]
    uint32_t *stage = reinterpret_cast<uint32_t *>(smem_raw + ((size_t)c_per * G * sizeof(uint16_t) + 15) / 16 * 16);   // [S][DK][rowp]
    __shared__ __align__(8) uint64_t full[S];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, p = blockIdx.y, g0 = blockIdx.x * G;
    constexpr int f = 8;                         // the launcher takes this kernel only for f = 8, logD = 2 (every set in params.jl)
    const int N = a.N, n = a.n, row = n + 1, rowp = a.rowp;
    const int ng = min(G, batch - g0);
    auto A = [&](int g, int comp, int c) -> uint32_t {
        const size_t off = ((size_t)(g0 + g) * (a.k + 1) + comp) * N + c;
        return a.bits64 ? (uint32_t)(reinterpret_cast<const uint64_t *>(a.acc)[off] >> 32)
                        : reinterpret_cast<const uint32_t *>(a.acc)[off];
    };
    auto extract = [&](int g, int c) -> uint32_t { return c == 0 ? A(g, 1 + p, 0) : 0u - A(g, 1 + p, N - c); };

    if (tid == 0) {
        for (int s = 0; s < S; s++) mbar_init(&full[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    // stage digits: field lv (0 = most significant) of coefficient c sits at bits [2*(f-1-lv), +2)
    for (int i = tid; i < (c_end - c_begin) * G; i += MK_THREADS) {
        const int c = c_begin + i / G, g = i % G;
        uint16_t packed = 0;
        if (g < ng) {
            uint32_t ai = divbits<uint32_t>(extract(g, c), 32 - f * a.logD);
            if (!BLOCK) packed = (uint16_t)ai;                          // unbalanced digits (gsw.jl:34-40) are the bit fields
            else {                                                      // balanced: gsw.jl:42-52, two's-complement 2-bit fields
                uint32_t acc_bits = 0;
                for (int lv = f - 1; lv >= 1; lv--) {
                    const uint32_t d = ai & 3u;
                    ai >>= 2; ai += d >> 1;
                    acc_bits |= d << (2 * (f - 1 - lv));
                }
                acc_bits |= (ai & 3u) << (2 * (f - 1));
                packed = (uint16_t)acc_bits;
            }
        }
        dig[i] = packed;
    }
    __syncthreads();

    const uint32_t *ksk = a.ksk[p];
    const size_t dig_stride = (size_t)f * rowp;
    const uint32_t row_bytes = (uint32_t)rowp * 4;
    const int steps = c_end > c_begin ? (c_end - c_begin) * f : 0;
    auto issue = [&](int it) {              // one thread: DK bulk copies of one padded row each into stage it % S
        const int c = c_begin + it / f, lv = it % f, s = it % S;
        const uint32_t *src = ksk + (size_t)c * DK * dig_stride + (size_t)lv * rowp;
        mbar_expect_tx(&full[s], DK * row_bytes);
#pragma unroll
        for (int v = 0; v < DK; v++) tma_bulk_g2s(stage + ((size_t)s * DK + v) * rowp, src + (size_t)v * dig_stride, row_bytes, &full[s]);
    };
    if (tid == 0) for (int it = 0; it < S - 1 && it < steps; it++) issue(it);

    // lane owns words 4*lane + 128*j .. +3 of the LWE row (16-byte shared-memory reads); jn = live j range of this lane
    uint4 sum[KS_GW][KS_J];
#pragma unroll
    for (int gw = 0; gw < KS_GW; gw++)
#pragma unroll
        for (int j = 0; j < KS_J; j++) sum[gw][j] = make_uint4(0u, 0u, 0u, 0u);
    const int jn = 4 * lane < rowp ? (rowp - 4 * lane + 127) / 128 : 0;

    if (tid == 0 && S - 1 < steps) issue(S - 1);       // (the loop below refills two stages per barrier)
    for (int it = 0; it < steps; it++) {
        if ((it & 1) == 0) {
            __syncthreads();                // every warp is done with steps it - 2, it - 1, whose stages are refilled next
            if (tid == 0 && it >= 2) {
                if (it + S - 2 < steps) issue(it + S - 2);
                if (it + S - 1 < steps) issue(it + S - 1);
            }
        }
        mbar_wait(&full[it % S], (uint32_t)((it / S) & 1));
        const uint32_t *cur = stage + (size_t)(it % S) * DK * rowp;
        const int ci = it / f, sh = 2 * (f - 1 - it % f);
        const uint2 dpack = *reinterpret_cast<const uint2 *>(&dig[ci * G + warp * KS_GW]);      // the warp's four gates
#pragma unroll
        for (int gw = 0; gw < KS_GW; gw++) {
            const uint32_t dw = gw < 2 ? dpack.x : dpack.y;
            const uint32_t d = (((gw & 1) ? dw >> 16 : dw & 0xFFFFu) >> sh) & 3u;              // uniform over the warp
            if (d == 0) continue;
            // unbalanced: +row d-1.  balanced: d = 1: +row 0; d = 3 (-1): -row 0; d = 2 (-2): -row 1
            const uint4 *r = reinterpret_cast<const uint4 *>(cur + (BLOCK ? (d == 2 ? 1 : 0) : (int)d - 1) * rowp) + lane;
            if (!BLOCK || d == 1) {
#pragma unroll
                for (int j = 0; j < KS_J; j++) if (j < jn) {
                    const uint4 v = r[32 * j];
                    sum[gw][j].x += v.x; sum[gw][j].y += v.y; sum[gw][j].z += v.z; sum[gw][j].w += v.w;
                }
            } else {
#pragma unroll
                for (int j = 0; j < KS_J; j++) if (j < jn) {
                    const uint4 v = r[32 * j];
                    sum[gw][j].x -= v.x; sum[gw][j].y -= v.y; sum[gw][j].z -= v.z; sum[gw][j].w -= v.w;
                }
            }
        }
    }
#pragma unroll
    for (int gw = 0; gw < KS_GW; gw++) {
        const int g = warp * KS_GW + gw;
        if (g >= ng) continue;
        uint32_t *out = a.out + (size_t)(g0 + g) * (1 + (size_t)n * a.k);
        const bool first = blockIdx.z == 0;
#pragma unroll
        for (int j = 0; j < KS_J; j++) {
            const uint32_t w4[4] = {sum[gw][j].x, sum[gw][j].y, sum[gw][j].z, sum[gw][j].w};
#pragma unroll
            for (int q = 0; q < 4; q++) {
                const int col = 4 * lane + 128 * j + q;
                if (col == 0) atomicAdd(out, w4[q] + (p == 0 && first ? A(g, 0, 0) : 0u));  // res.b = acc.b[0] + sum of parts
                else if (col < row) atomicAdd(out + 1 + (size_t)p * n + (col - 1), w4[q] + ((BLOCK && first && col - 1 < n) ? extract(g, col - 1) : 0u));
            }
        }
    }
}
// digits of one coefficient slice + S stages of DK rows
static inline size_t keyswitch_tiled_smem(int c_per, int rowp) {
    return ((size_t)c_per * KS_G * sizeof(uint16_t) + 15) / 16 * 16 + (size_t)KS_STAGES * 3 * rowp * 4;
}
