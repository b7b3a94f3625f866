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
