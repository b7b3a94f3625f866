// keygen.cuh -- evaluation-key generation ON THE DEVICE.
//
// What it computes: the same key material as csrc/host_keygen.cpp, i.e. the reference's
//   party_keygen / setup   /root/reference/src/tfhe/keygen.jl:3-155 (BootKey_bin / _block / _CCS / _KMS / _KMS_block)
//   rgsw_encrypt           /root/reference/src/ciphertext/gsw.jl:174-184
//   unienc_encrypt, gen_b  /root/reference/src/ciphertext/unienc.jl:36-90
//   lev_encrypt            /root/reference/src/ciphertext/lev.jl:31-45
// from the SAME seeded ChaCha20 streams (stream = (kind, party, index), counter-addressed, so every value has a fixed position and
// all of it is generated in parallel), with exact integer negacyclic products (the keys are binary / ternary) and the STRICT
// transform (the reference's Float64 FFT, bit-exact with the oracle) for the uploaded FFT form.  The result is byte-identical to
// the host library's for the same seed wherever the normal deviates round the same way (the device's log / sin / cos may differ
// from glibc's in the last bit: a rounded noise value then differs with probability ~1e-13 per sample); tests compare SHA-256.
// Nothing crosses PCIe: a 32-party KMS key set is 10.6 GB that the host path generates in ~15 s and uploads in ~3 s.
//
// Stream positions (32-bit words from the start of a stream; wT = words per torus element):
//   S_BRK, RGSW (party, i):  row r = basket*l + j at r*(N*wT + 2N): a (N*wT words), then N normals (2N words)
//   unienc (S_RLK, or S_BRK of CCS): r (N words, ternary = word % 3 - 1), l rows of N normals (d), then l RLWE rows (f)
//   S_PUBB: l rows of N normals;  S_CRS: l*N torus values;  S_KSK (party, c): row q at q*n + 4*ceil(q/2): n mask words, and one
//   normal per row taken alternately as the cosine / sine branch of one Box-Muller pair (host: cached spare).
#pragma once
#include "common.cuh"
#include "kernels_strict.cuh"

namespace kg {

struct Key { uint32_t k[8]; };

enum : uint64_t { S_CRS = 1, S_LWEKEY, S_RINGKEY, S_GSWKEY, S_BRK, S_RLK, S_PUBB, S_KSK, S_ENC };
__host__ __device__ inline uint64_t stream_id(uint64_t kind, int party, uint64_t idx) {
    return (kind << 56) | ((uint64_t)(party & 0xFFFF) << 40) | (idx & 0xFFFFFFFFFFull);
}

__device__ __forceinline__ uint32_t rotl32(uint32_t v, int c) { return (v << c) | (v >> (32 - c)); }
// ChaCha20 block (DJB variant, 64-bit counter in words 12-13, stream id in words 14-15), as csrc/host_keygen.cpp
__device__ inline void chacha_block(const Key &key, uint64_t stream, uint64_t counter, uint32_t (&out)[16]) {
    uint32_t st[16] = {0x61707865u, 0x3320646eu, 0x79622d32u, 0x6b206574u, key.k[0], key.k[1], key.k[2], key.k[3], key.k[4], key.k[5],
                       key.k[6], key.k[7], (uint32_t)counter, (uint32_t)(counter >> 32), (uint32_t)stream, (uint32_t)(stream >> 32)};
    uint32_t x[16];
#pragma unroll
    for (int i = 0; i < 16; i++) x[i] = st[i];
#define KG_QR(a, b, c, d)                                                                    \
    x[a] += x[b]; x[d] = rotl32(x[d] ^ x[a], 16); x[c] += x[d]; x[b] = rotl32(x[b] ^ x[c], 12); \
    x[a] += x[b]; x[d] = rotl32(x[d] ^ x[a], 8);  x[c] += x[d]; x[b] = rotl32(x[b] ^ x[c], 7);
#pragma unroll 1
    for (int r = 0; r < 10; r++) {
        KG_QR(0, 4, 8, 12) KG_QR(1, 5, 9, 13) KG_QR(2, 6, 10, 14) KG_QR(3, 7, 11, 15)
        KG_QR(0, 5, 10, 15) KG_QR(1, 6, 11, 12) KG_QR(2, 7, 8, 13) KG_QR(3, 4, 9, 14)
    }
#undef KG_QR
#pragma unroll
    for (int i = 0; i < 16; i++) out[i] = x[i] + st[i];
}

// words [off, off + count) of a stream into dst (shared or global), cooperatively by the CTA
__device__ inline void fill_words(uint32_t *dst, const Key &key, uint64_t stream, uint64_t off, uint32_t count) {
    const uint64_t b0 = off >> 4, b1 = (off + count + 15) >> 4;
    for (uint64_t b = b0 + threadIdx.x; b < b1; b += blockDim.x) {
        uint32_t w[16];
        chacha_block(key, stream, b, w);
#pragma unroll
        for (int i = 0; i < 16; i++) {
            const uint64_t pos = (b << 4) + i;
            if (pos >= off && pos < off + count) dst[pos - off] = w[i];
        }
    }
}

// one Box-Muller pair from four stream words, as ChaCha20::normal() (host): (r cos, r sin)
__device__ __forceinline__ void normal_pair(const uint32_t *w, double &n0, double &n1) {
    const uint64_t a = (uint64_t)w[0] | ((uint64_t)w[1] << 32), b = (uint64_t)w[2] | ((uint64_t)w[3] << 32);
    const double u1 = (double)((a >> 11) + 1) * (1.0 / 9007199254740993.0), u2 = (double)((b >> 11) + 1) * (1.0 / 9007199254740993.0);
    const double r = sqrt(-2.0 * log(u1)), th = 6.283185307179586476925 * u2;
    double sn, cs;
    sincos(th, &sn, &cs);
    n0 = r * cs; n1 = r * sn;
}
template <class T> __device__ __forceinline__ T noise_of(double sigma, double nrm) {
    return (T)(typename Torus<T>::S)rint(sigma * nrm);                           // unsigned(round(signed(T), gaussian)) : lwe.jl:12,89
}

// ---- secrets (keygen_params: scheme.jl:118-119,183-187,221-222,267-268,314-319; sampler.jl:1-21; key.jl:52-88) ------------------
// one CTA; lwekey [n], ringkey [N], gswkey [N] as int8
__global__ void k_kg_secrets(Key key, int party, int n, int N, int block, int d, int ell, int8_t *lwekey, int8_t *ringkey, int8_t *gswkey) {
    extern __shared__ uint32_t kw[];
    const int nl = block ? d : n;
    fill_words(kw, key, stream_id(S_LWEKEY, party, 0), 0, nl);
    __syncthreads();
    if (!block) for (int i = threadIdx.x; i < n; i += blockDim.x) lwekey[i] = (int8_t)(kw[i] & 1);
    else {
        for (int i = threadIdx.x; i < n; i += blockDim.x) lwekey[i] = 0;
        __syncthreads();
        for (int i = threadIdx.x; i < d; i += blockDim.x) {
            const uint32_t idx = kw[i] % (uint32_t)(ell + 1);                      // below(ell + 1): no rejection for ell + 1 = 4
            if (idx) lwekey[i * ell + idx - 1] = 1;
        }
    }
    __syncthreads();
    fill_words(kw, key, stream_id(S_RINGKEY, party, 0), 0, N);
    __syncthreads();
    for (int i = threadIdx.x; i < N; i += blockDim.x) ringkey[i] = (block && i < n) ? lwekey[i] : (int8_t)(kw[i] & 1);
    __syncthreads();
    fill_words(kw, key, stream_id(S_GSWKEY, party, 0), 0, N);
    __syncthreads();
    for (int i = threadIdx.x; i < N; i += blockDim.x) gswkey[i] = (int8_t)(kw[i] & 1);
}

// ternary vectors r = below(3) - 1 (key.jl:41-50) at the start of `count` streams (kind, party, idx0 + s): out [count][N]
__global__ void k_kg_ternary(Key key, uint64_t kind, int party, uint64_t idx0, int N, int8_t *out) {
    extern __shared__ uint32_t kw[];
    fill_words(kw, key, stream_id(kind, party, idx0 + blockIdx.x), 0, N);
    __syncthreads();
    for (int i = threadIdx.x; i < N; i += blockDim.x) out[(size_t)blockIdx.x * N + i] = (int8_t)((int)(kw[i] % 3u) - 1);
}

// uniform torus polynomials straight from a stream (CRS: scheme.jl:409-410, polynomial.jl:48-49): out[i] for i < count
template <class T> __global__ void k_kg_uniform(Key key, uint64_t stream, T *out, size_t count) {
    constexpr int WT = sizeof(T) / 4;
    const size_t blocks = (count * WT + 15) / 16;
    for (size_t b = (size_t)blockIdx.x * blockDim.x + threadIdx.x; b < blocks; b += (size_t)gridDim.x * blockDim.x) {
        uint32_t w[16];
        chacha_block(key, stream, b, w);
#pragma unroll
        for (int i = 0; i < 16 / WT; i++) {
            const size_t e = b * (16 / WT) + i;
            if (e < count) out[e] = WT == 2 ? (T)((uint64_t)w[2 * i] | ((uint64_t)w[2 * i + 1] << 32)) : (T)w[i];
        }
    }
}

// ---- one ring row:  out0 = sign * (A (*) s) + e (+ message),  optionally out1 = A --------------------------------------------
// A is either drawn from the stream (an RLWE mask, also written to out1) or given (a CRS polynomial); s has entries in {-1, 0, 1}.
template <class T> struct RowDesc {
    uint64_t stream, a_off, e_off;     // stream id, word offsets of the mask (gen_a) and of the N normals
    const T *a_given;                  // !gen_a
    const int8_t *s;                   // [N]
    const int8_t *msg_vec;             // msg_mode 3: out0[c] += msg * msg_vec[c]
    T *out0, *out1;
    T msg;
    double sigma;
    int gen_a, negate, msg_mode;       // msg_mode: 0 none, 1 out0[0] += msg, 2 out1[0] += msg, 3 see msg_vec
};

template <class T, int N>
__global__ void __launch_bounds__(256) k_kg_rows(Key key, const RowDesc<T> *rows) {
    constexpr int WT = sizeof(T) / 4;
    extern __shared__ __align__(16) unsigned char kraw[];
    T *A = reinterpret_cast<T *>(kraw);                                   // [N]  (little endian: the mask words ARE the torus values)
    uint32_t *ew = reinterpret_cast<uint32_t *>(A + N);                   // [2N] noise words
    T *E = reinterpret_cast<T *>(ew + 2 * N);                             // [N]
    int16_t *pos = reinterpret_cast<int16_t *>(E + N);                    // [N] nonzero positions of s, sign in bit 15
    __shared__ int npos;
    const RowDesc<T> r = rows[blockIdx.x];
    if (threadIdx.x == 0) npos = 0;
    if (r.gen_a) fill_words(reinterpret_cast<uint32_t *>(A), key, r.stream, r.a_off, N * WT);
    else for (int c = threadIdx.x; c < N; c += blockDim.x) A[c] = r.a_given[c];
    fill_words(ew, key, r.stream, r.e_off, 2 * N);
    __syncthreads();
    for (int q = threadIdx.x; q < N / 2; q += blockDim.x) {
        double n0, n1;
        normal_pair(ew + 4 * q, n0, n1);
        E[2 * q] = noise_of<T>(r.sigma, n0); E[2 * q + 1] = noise_of<T>(r.sigma, n1);
    }
    // compact the nonzero entries of s (order is irrelevant: integer sums are exact)
    for (int j = threadIdx.x; j < N; j += blockDim.x) {
        const int8_t v = r.s[j];
        if (v) pos[atomicAdd(&npos, 1)] = (int16_t)(j | (v < 0 ? 0x8000 : 0));
    }
    __syncthreads();
    const int np = npos;
    constexpr int PER = N / 256;
    T sum[PER];
#pragma unroll
    for (int i = 0; i < PER; i++) sum[i] = 0;
    for (int p = 0; p < np; p++) {
        const int pj = pos[p], j = pj & 0x7FFF;
        const bool neg = pj & 0x8000;
#pragma unroll
        for (int i = 0; i < PER; i++) {
            const int c = threadIdx.x + 256 * i, d = c - j;
            const T v = A[d & (N - 1)];
            // X^N = -1: the term wraps with a sign flip when c < j
            if ((d < 0) != neg) sum[i] -= v; else sum[i] += v;
        }
    }
#pragma unroll
    for (int i = 0; i < PER; i++) {
        const int c = threadIdx.x + 256 * i;
        T v = (r.negate ? (T)((T)0 - sum[i]) : sum[i]) + E[c];
        if (r.msg_mode == 1 && c == 0) v += r.msg;
        if (r.msg_mode == 3) v += r.msg * (T)(typename Torus<T>::S)r.msg_vec[c];
        r.out0[c] = v;
        if (r.out1) { T av = A[c]; if (r.msg_mode == 2 && c == 0) av += r.msg; r.out1[c] = av; }
    }
}
template <class T, int N> constexpr size_t rows_smem() { return (size_t)N * sizeof(T) * 2 + (size_t)2 * N * 4 + (size_t)N * 2; }

// ---- key-switching key: ksk[c][digit-1][level] = LWE(ringkey[c] * digit * g_level) under lwekey (keygen.jl:16-24,43-52,75-79,110-114,147-151)
// one CTA per ring coefficient c; rows of rowp words (16-byte aligned stride, as uploaded keys are stored)
__global__ void __launch_bounds__(256) k_kg_ksk(Key key, int party, int n, int N, int Dk, int f, int logD, int block, double alpha,
                                                 const int8_t *lwekey, const int8_t *ringkey, uint32_t *ksk, int rowp) {
    extern __shared__ uint32_t kw[];
    const int c = blockIdx.x, nrows = Dk * f;
    uint32_t *base = ksk + (size_t)c * nrows * rowp;
    if (block && c < n) {                                                   // never read (undef in the reference)
        for (int i = threadIdx.x; i < nrows * rowp; i += blockDim.x) base[i] = 0u;
        return;
    }
    const uint32_t total = (uint32_t)nrows * n + 4u * ((nrows + 1) / 2);
    fill_words(kw, key, stream_id(S_KSK, party, (uint64_t)c), 0, total);
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int q = warp; q < nrows; q += blockDim.x / 32) {
        const uint32_t off = (uint32_t)q * n + 4u * ((q + 1) / 2);         // words consumed before row q
        const uint32_t *a = kw + off;
        uint32_t dot = 0;
        for (int i = lane; i < n; i += 32) {
            const uint32_t v = a[i];
            base[(size_t)q * rowp + 1 + i] = v;
            if (lwekey[i]) dot += v;
        }
#pragma unroll
        for (int o = 16; o; o >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, o);
        if (lane == 0) {
            // rows 2p and 2p + 1 share one Box-Muller pair: its words follow the mask of row 2p
            const uint32_t *pw = kw + (uint32_t)(q & ~1) * n + 4u * (((q & ~1) + 1) / 2) + n;
            double n0, n1;
            normal_pair(pw, n0, n1);
            const int dg = q / f + 1, lv = q % f;
            const uint32_t m = (uint32_t)(int32_t)ringkey[c] * (uint32_t)dg * ((uint32_t)1 << (32 - (lv + 1) * logD));
            base[(size_t)q * rowp] = (0u - dot) + noise_of<uint32_t>(alpha, (q & 1) ? n1 : n0) + m;
        }
        for (int i = n + 1 + lane; i < rowp; i += 32) base[(size_t)q * rowp + i] = 0u;
    }
}

}  // namespace kg
