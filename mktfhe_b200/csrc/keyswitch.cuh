// keyswitch.cuh -- integer stages: gate linear part, modulus switch, sample extraction + key switch.
// All arithmetic is UInt32 / UInt64 wraparound and must be bit-exact against the reference.
//
//   k_gate_prep      gate.jl:1-52 (linear part) + bootstrapping.jl:8-9 (modswitch by divbits)
//   k_keyswitch      keyswitch! CGGI bootstrapping.jl:81-109, LMSS :170-229, CCS :333-364,
//                    KMS :564-594, KMS_block :664-695
#pragma once
#include "common.cuh"

// lin = gate linear combination (or copy when op < 0); tilde = divbits(lin, 32 - log2(N) - 1).
__global__ void k_gate_prep(const uint32_t *in1, const uint32_t *in2, uint32_t *lin, uint32_t *tilde,
                            int op, int lwe_words, size_t total, int shift) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const bool is_b = (i % lwe_words) == 0;
    uint32_t v;
    if (op < 0) v = in1[i];
    else {
        const uint32_t s = in1[i] + in2[i];
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

struct KsArgs {
    const void *acc;              // [B][(k+1)][N] torus
    const uint32_t *const *ksk;   // [k]: [N][Dk][f][n+1]
    uint32_t *out;                // [B][1 + n*k]
    int N, n, k, f, logD, Dk, bits64, block;
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
            const uint32_t *base = ksk + (size_t)c * a.Dk * f * row;
#pragma unroll 4
            for (int lv = 0; lv < f; lv++) {
                const int d = dig[c * f + lv];
                if (d == 0) continue;
                const uint32_t *r = base + ((size_t)((d > 0 ? d : -d) - 1) * f + lv) * row;
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
// Tiled key switch (production): one CTA = G gates x one party.  All G gates walk (c, level) together, so
// each selected ksk row is fetched from L2 once per tile and served to the other gates from L1; the digits
// of the tile are staged in shared memory as 2-bit fields (f*logD = 16 bits per coefficient).
// Integer adds commute, so the per-party partial sums of b are combined with atomicAdd into a zeroed output.
template <bool BLOCK, int G>
__global__ void __launch_bounds__(MK_THREADS) k_keyswitch_tiled(const KsArgs a, int batch) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint16_t *dig = reinterpret_cast<uint16_t *>(smem_raw);             // [N][G]
    const int tid = threadIdx.x, p = blockIdx.y, g0 = blockIdx.x * G;
    const int N = a.N, n = a.n, f = a.f, row = n + 1;
    constexpr int MAXC = 3;
    const int ng = min(G, batch - g0);
    auto A = [&](int g, int comp, int c) -> uint32_t {
        const size_t off = ((size_t)(g0 + g) * (a.k + 1) + comp) * N + c;
        return a.bits64 ? (uint32_t)(reinterpret_cast<const uint64_t *>(a.acc)[off] >> 32)
                        : reinterpret_cast<const uint32_t *>(a.acc)[off];
    };
    auto extract = [&](int g, int c) -> uint32_t { return c == 0 ? A(g, 1 + p, 0) : 0u - A(g, 1 + p, N - c); };

    // stage digits: field lv (0 = most significant) of coefficient c sits at bits [2*(f-1-lv), +2)
    for (int i = tid; i < N * G; i += MK_THREADS) {
        const int c = i / G, g = i % G;
        uint16_t packed = 0;
        if (g < ng && !(BLOCK && c < n)) {
            uint32_t ai = divbits<uint32_t>(extract(g, c), 32 - f * a.logD);
            if (!BLOCK) packed = (uint16_t)ai;                          // unbalanced digits are the bit fields themselves
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
        dig[c * G + g] = packed;
    }
    __syncthreads();

    uint32_t sum[G][MAXC];
#pragma unroll
    for (int g = 0; g < G; g++)
#pragma unroll
        for (int q = 0; q < MAXC; q++) sum[g][q] = 0u;
    const uint32_t *ksk = a.ksk[p];
    const size_t lvl_stride = row, dig_stride = (size_t)f * row;
    const bool has[MAXC] = {tid < row, tid + MK_THREADS < row, tid + 2 * MK_THREADS < row};

    // Per (c, level): fetch the Dk candidate rows first (independent coalesced loads, one L2 latency, pipelined across
    // iterations), then let every gate of the tile pick its row with a warp-uniform branch: no memory operation sits
    // on the gates' dependency chain.
    constexpr int DK = BLOCK ? 2 : 3;
    for (int c = BLOCK ? n : 0; c < N; c++) {
        const uint32_t *base = ksk + (size_t)c * DK * dig_stride + tid;
        uint32_t w[G];
#pragma unroll
        for (int g = 0; g < G; g++) w[g] = dig[c * G + g];
#pragma unroll 4
        for (int lv = 0; lv < f; lv++) {
            const int sh = 2 * (f - 1 - lv);
            const uint32_t *lbase = base + lv * lvl_stride;
            uint32_t x[DK][MAXC];
#pragma unroll
            for (int v = 0; v < DK; v++)
#pragma unroll
                for (int q = 0; q < MAXC; q++) x[v][q] = has[q] ? __ldg(lbase + (size_t)v * dig_stride + q * MK_THREADS) : 0u;
#pragma unroll
            for (int g = 0; g < G; g++) {
                const uint32_t d = (w[g] >> sh) & 3u;                   // uniform over the CTA
                if (!BLOCK) {
                    if (d == 1) {
#pragma unroll
                        for (int q = 0; q < MAXC; q++) sum[g][q] += x[0][q];
                    } else if (d == 2) {
#pragma unroll
                        for (int q = 0; q < MAXC; q++) sum[g][q] += x[1][q];
                    } else if (d == 3) {
#pragma unroll
                        for (int q = 0; q < MAXC; q++) sum[g][q] += x[2][q];
                    }
                } else {                                                // d = 1: +row 1; d = 3 (-1): -row 1; d = 2 (-2): -row 2
                    if (d == 1) {
#pragma unroll
                        for (int q = 0; q < MAXC; q++) sum[g][q] += x[0][q];
                    } else if (d == 3) {
#pragma unroll
                        for (int q = 0; q < MAXC; q++) sum[g][q] -= x[0][q];
                    } else if (d == 2) {
#pragma unroll
                        for (int q = 0; q < MAXC; q++) sum[g][q] -= x[1][q];
                    }
                }
            }
        }
    }
#pragma unroll
    for (int g = 0; g < G; g++) {
        if (g >= ng) break;
        uint32_t *out = a.out + (size_t)(g0 + g) * (1 + (size_t)n * a.k);
#pragma unroll
        for (int q = 0; q < MAXC; q++) {
            const int col = tid + q * MK_THREADS;
            if (col == 0) atomicAdd(out, sum[g][q] + (p == 0 ? A(g, 0, 0) : 0u));      // res.b = acc.b[0] + sum of parts
            else if (col < row) out[1 + (size_t)p * n + (col - 1)] = sum[g][q] + ((BLOCK && col - 1 < n) ? extract(g, col - 1) : 0u);
        }
    }
}
