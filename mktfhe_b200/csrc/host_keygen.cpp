// host_keygen.cpp -- host-side key generation / encryption / decryption (see include/mktfhe_host.h).
//
// Mirrors, on the CPU and with a seeded ChaCha20 stream, what the reference's Julia host code builds:
//   samplers        /root/reference/src/ring/sampler.jl:1-34
//   keys            /root/reference/src/ciphertext/key.jl:1-88
//   LWE/RLWE/LEV/RGSW/UniEnc encryption  src/ciphertext/{lwe,lev,gsw,unienc}.jl
//   BootKey_*       /root/reference/src/tfhe/keygen.jl:3-155
//   encrypt/decrypt /root/reference/src/tfhe/scheme.jl:352-410
// Not on the GPU hot path; this is the part of the drop-in surface that stays on the host.
#include "../../include/mktfhe_host.h"

#include <cmath>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <mutex>
#include <map>
#include <quadmath.h>
#ifdef _OPENMP
#include <omp.h>
#endif

namespace {

// ------------------------------------------------------------------ ChaCha20 stream (DJB variant, 64-bit counter)
struct ChaCha20 {
    uint32_t st[16];
    uint32_t buf[16];
    int pos = 16;
    bool have_spare = false;
    double spare = 0.0;

    static uint64_t splitmix(uint64_t &x) {
        uint64_t z = (x += 0x9E3779B97F4A7C15ull);
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
        return z ^ (z >> 31);
    }
    // The 256-bit ChaCha key.  Production callers pass 32 bytes from the OS CSPRNG (one independent key per party, per
    // CRS and per encryption, like the reference's unseeded ChaCha20Stream() calls: sampler.jl:2-34, lwe.jl:13).  A 64-bit
    // integer seed expands to a key through splitmix64: reproducible streams for tests, benchmarks and golden vectors only.
    struct Key {
        uint32_t k[8];
        static Key from_seed(uint64_t seed) {
            Key r; uint64_t x = seed;
            for (int i = 0; i < 4; i++) { uint64_t kx = splitmix(x); r.k[2 * i] = (uint32_t)kx; r.k[2 * i + 1] = (uint32_t)(kx >> 32); }
            return r;
        }
        static Key from_bytes(const uint8_t *b) { Key r; memcpy(r.k, b, 32); return r; }
    };
    ChaCha20(const Key &key, uint64_t stream) {
        st[0] = 0x61707865; st[1] = 0x3320646e; st[2] = 0x79622d32; st[3] = 0x6b206574;
        for (int i = 0; i < 8; i++) st[4 + i] = key.k[i];
        st[12] = 0; st[13] = 0;
        st[14] = (uint32_t)stream; st[15] = (uint32_t)(stream >> 32);
    }
    static inline uint32_t rotl(uint32_t v, int c) { return (v << c) | (v >> (32 - c)); }
    void refill() {
        uint32_t x[16];
        memcpy(x, st, sizeof(x));
#define QR(a, b, c, d) \
    x[a] += x[b]; x[d] = rotl(x[d] ^ x[a], 16); x[c] += x[d]; x[b] = rotl(x[b] ^ x[c], 12); \
    x[a] += x[b]; x[d] = rotl(x[d] ^ x[a], 8);  x[c] += x[d]; x[b] = rotl(x[b] ^ x[c], 7);
        for (int r = 0; r < 10; r++) {
            QR(0, 4, 8, 12) QR(1, 5, 9, 13) QR(2, 6, 10, 14) QR(3, 7, 11, 15)
            QR(0, 5, 10, 15) QR(1, 6, 11, 12) QR(2, 7, 8, 13) QR(3, 4, 9, 14)
        }
#undef QR
        for (int i = 0; i < 16; i++) buf[i] = x[i] + st[i];
        if (++st[12] == 0) ++st[13];
        pos = 0;
    }
    inline uint32_t u32() { if (pos == 16) refill(); return buf[pos++]; }
    inline uint64_t u64() { uint64_t lo = u32(); return lo | ((uint64_t)u32() << 32); }
    inline double unif() { return (double)((u64() >> 11) + 1) * (1.0 / 9007199254740993.0); }   // (0, 1)
    // standard normal, Box-Muller (the reference uses Julia's randn; distribution parity only)
    double normal() {
        if (have_spare) { have_spare = false; return spare; }
        const double u1 = unif(), u2 = unif();
        const double r = std::sqrt(-2.0 * std::log(u1)), th = 6.283185307179586476925 * u2;
        spare = r * std::sin(th); have_spare = true;
        return r * std::cos(th);
    }
    // uniform integer in [0, bound)
    uint32_t below(uint32_t bound) {
        const uint32_t lim = (uint32_t)(0x100000000ull / bound) * bound;
        uint32_t v;
        do v = u32(); while (lim != 0 && v >= lim);
        return v % bound;
    }
    template <class T> T torus();
};
template <> inline uint32_t ChaCha20::torus<uint32_t>() { return u32(); }
template <> inline uint64_t ChaCha20::torus<uint64_t>() { return u64(); }

enum StreamKind : uint64_t { S_CRS = 1, S_LWEKEY, S_RINGKEY, S_GSWKEY, S_BRK, S_RLK, S_PUBB, S_KSK, S_ENC };
inline uint64_t stream_id(StreamKind kind, int party, uint64_t idx) {
    return ((uint64_t)kind << 56) | ((uint64_t)(party & 0xFFFF) << 40) | (idx & 0xFFFFFFFFFFull);
}

// ------------------------------------------------------------------ Float64 transform (fft.jl:18-63, 105-155)
struct cplx { double re, im; };
struct Tables {
    int N, H;
    std::vector<cplx> psi, psiinv, roots, rootsinv;
};
void bit_reverse(std::vector<cplx> &v) {
    const int n = (int)v.size();
    for (int i = 1, j = 0; i < n; i++) {
        int bit = n >> 1;
        for (; j >= bit; bit >>= 1) j -= bit;
        j += bit;
        if (i < j) std::swap(v[i], v[j]);
    }
}
const Tables &tables_for(int N) {
    static std::mutex mu;
    static std::map<int, Tables *> cache;
    std::lock_guard<std::mutex> g(mu);
    auto it = cache.find(N);
    if (it != cache.end()) return *it->second;
    Tables *t = new Tables;
    const int H = N / 2;
    t->N = N; t->H = H;
    t->psi.resize(H); t->psiinv.resize(H); t->roots.resize(H); t->rootsinv.resize(H);
    for (int j = 0; j < H; j++) {
        // BigFloat exp() rounded to Float64 in the reference; binary128 here (correctly rounded in practice).
        const __float128 th = M_PIq * j / H, ph = M_PIq * j / N;
        t->psi[j] = { (double)cosq(th), (double)-sinq(th) };
        t->psiinv[j] = { (double)cosq(th), (double)sinq(th) };
        t->roots[j] = { (double)cosq(ph), (double)sinq(ph) };
        t->rootsinv[j] = { (double)(cosq(ph) / H), (double)(-sinq(ph) / H) };
    }
    bit_reverse(t->psi);
    bit_reverse(t->psiinv);
    cache[N] = t;
    return *t;
}

template <class T> struct Signed;
template <> struct Signed<uint32_t> { typedef int32_t type; };
template <> struct Signed<uint64_t> { typedef int64_t type; };

// fftto! (fft.jl:57-63): coefficient form -> H complex in the reference's slot order.
template <class T> void fft_poly(const T *p, cplx *out, const Tables &tb) {
    typedef typename Signed<T>::type S;
    const int H = tb.H;
    for (int i = 0; i < H; i++) {
        const double ar = (double)(S)p[i], ai = (double)(S)(T)((T)0 - p[H + i]);
        const cplx w = tb.roots[i];
        out[i] = { ar * w.re - ai * w.im, ar * w.im + ai * w.re };
    }
    for (int m = 1, k = H >> 1; m < H; m <<= 1, k >>= 1)
        for (int i = 0; i < m; i++) {
            const cplx w = tb.psi[m + i];
            for (int j = 2 * i * k; j < 2 * i * k + k; j++) {
                const cplx t = out[j], v = out[j + k];
                const cplx u = { v.re * w.re - v.im * w.im, v.re * w.im + v.im * w.re };
                out[j] = { t.re + u.re, t.im + u.im };
                out[j + k] = { t.re - u.re, t.im - u.im };
            }
        }
}

// ------------------------------------------------------------------ ring helpers
// out = a * s in Z_{2^w}[X]/(X^N + 1), s with coefficients in {-1, 0, 1}: exact.
template <class T> void negacyclic_mul_small(T *out, const T *a, const int8_t *s, int N) {
    std::fill(out, out + N, (T)0);
    for (int j = 0; j < N; j++) {
        if (s[j] == 0) continue;
        if (s[j] > 0) {
            for (int c = j; c < N; c++) out[c] += a[c - j];
            for (int c = 0; c < j; c++) out[c] -= a[c - j + N];
        } else {
            for (int c = j; c < N; c++) out[c] -= a[c - j];
            for (int c = 0; c < j; c++) out[c] += a[c - j + N];
        }
    }
}

// unsigned(round(signed(T), gaussian(sigma)))  (lwe.jl:12, 89)
template <class T> inline T noise(ChaCha20 &rng, double sigma) {
    typedef typename Signed<T>::type S;
    return (T)(S)std::nearbyint(sigma * rng.normal());
}

// RLWEsample (lwe.jl:78-93), k = 1: a uniform, b = -a*s + e.
template <class T> void rlwe_sample(ChaCha20 &rng, const int8_t *key, double sigma, T *b, T *a, int N, T *scratch) {
    for (int c = 0; c < N; c++) a[c] = rng.torus<T>();
    negacyclic_mul_small(scratch, a, key, N);
    for (int c = 0; c < N; c++) b[c] = (T)((T)0 - scratch[c]) + noise<T>(rng, sigma);
}

template <class T> inline T gvec(int j, int logB) { return (T)1 << (sizeof(T) * 8 - (size_t)(j + 1) * logB); }   // gsw.jl:15-16

// rgsw_encrypt(m scalar) -> fft (gsw.jl:174-178, lev.jl:88-102, lwe.jl:95-105; keygen.jl:13,40,106,143)
// out: [basket][j][comp][H]
template <class T> void rgsw_fft(ChaCha20 &rng, T m, const int8_t *key, double sigma, int l, int logB, cplx *out, const Tables &tb) {
    const int N = tb.N, H = tb.H;
    std::vector<T> b(N), a(N), scr(N);
    for (int basket = 0; basket < 2; basket++)
        for (int j = 0; j < l; j++) {
            rlwe_sample(rng, key, sigma, b.data(), a.data(), N, scr.data());
            if (basket == 0) b[0] += gvec<T>(j, logB) * m; else a[0] += gvec<T>(j, logB) * m;
            fft_poly(b.data(), out + (size_t)((basket * l + j) * 2 + 0) * H, tb);
            fft_poly(a.data(), out + (size_t)((basket * l + j) * 2 + 1) * H, tb);
        }
}

// unienc_encrypt -> fft (unienc.jl:36-75).  msg: N coefficients (poly form) or scalar in msg[0] with the rest 0.
// out: [j][d, f.b, f.a][H]
template <class T> void unienc_fft(ChaCha20 &rng, const T *crs /*[l][N]*/, const T *msg, const int8_t *key, double sigma,
                                   int l, int logB, cplx *out, const Tables &tb) {
    const int N = tb.N, H = tb.H;
    std::vector<int8_t> r(N);
    for (int c = 0; c < N; c++) r[c] = (int8_t)((int)rng.below(3) - 1);        // ternary_ringkey (key.jl:41-50)
    std::vector<T> d(N), b(N), a(N), scr(N);
    for (int j = 0; j < l; j++) {
        negacyclic_mul_small(d.data(), crs + (size_t)j * N, r.data(), N);
        const T g = gvec<T>(j, logB);
        for (int c = 0; c < N; c++) d[c] += msg[c] * g + noise<T>(rng, sigma);
        fft_poly(d.data(), out + (size_t)(j * 3 + 0) * H, tb);
    }
    for (int j = 0; j < l; j++) {                                               // f = rlev_encrypt(r, key) (lev.jl:104-105)
        rlwe_sample(rng, key, sigma, b.data(), a.data(), N, scr.data());
        const T g = gvec<T>(j, logB);
        for (int c = 0; c < N; c++) b[c] += g * (T)(typename Signed<T>::type)r[c];
        fft_poly(b.data(), out + (size_t)(j * 3 + 1) * H, tb);
        fft_poly(a.data(), out + (size_t)(j * 3 + 2) * H, tb);
    }
}

void binary_vec(ChaCha20 &rng, int8_t *v, int n) { for (int i = 0; i < n; i++) v[i] = (int8_t)(rng.u32() & 1); }   // sampler.jl:1-2
void block_binary_vec(ChaCha20 &rng, int8_t *v, int d, int ell) {                                                 // sampler.jl:7-21
    std::fill(v, v + (size_t)d * ell, (int8_t)0);
    for (int i = 0; i < d; i++) {
        const uint32_t idx = rng.below((uint32_t)ell + 1);
        if (idx != 0) v[i * ell + idx - 1] = 1;
    }
}

// LWEsample + message (lwe.jl:11-22): out[0] = b, out[1..n] = a.
void lwe_sample(ChaCha20 &rng, const int8_t *key, int n, double sigma, uint32_t m, uint32_t *out) {
    uint32_t dot = 0;
    for (int i = 0; i < n; i++) { out[1 + i] = rng.u32(); if (key[i]) dot += out[1 + i]; }
    out[0] = (0u - dot) + noise<uint32_t>(rng, sigma) + m;
}

template <class T>
int party_keygen_impl(const mktfhe_params *p, const ChaCha20::Key &seed, int party, const T *crs, uint32_t *lwekey_out, T *ringkey_out,
                      double *brk, double *rlk, double *pubb, uint32_t *ksk, int nthreads) {
    const int N = p->N, H = N / 2, n = p->n;
    const Tables &tb = tables_for(N);
    const bool block = p->scheme == MKTFHE_LMSS || p->scheme == MKTFHE_KMS_BLOCK;
    const bool kms = p->scheme == MKTFHE_KMS || p->scheme == MKTFHE_KMS_BLOCK;
    const bool mk = kms || p->scheme == MKTFHE_CCS;
    if (mk && !crs) return -1;
    if (!mk && p->k != 1) return -2;
#ifdef _OPENMP
    if (nthreads <= 0) nthreads = omp_get_max_threads();
#else
    nthreads = 1;
#endif

    // keygen_params (scheme.jl:118-119, 183-187, 221-222, 267-268, 314-319)
    std::vector<int8_t> lwekey(n), ringkey(N), gswkey(N);
    {
        ChaCha20 r1(seed, stream_id(S_LWEKEY, party, 0));
        if (block) block_binary_vec(r1, lwekey.data(), p->d, p->ell); else binary_vec(r1, lwekey.data(), n);
        ChaCha20 r2(seed, stream_id(S_RINGKEY, party, 0));
        binary_vec(r2, ringkey.data(), N);
        if (block) std::copy(lwekey.begin(), lwekey.end(), ringkey.begin());       // partial_ringkey (key.jl:52-88), n < N
        ChaCha20 r3(seed, stream_id(S_GSWKEY, party, 0));
        binary_vec(r3, gswkey.data(), N);
    }
    for (int i = 0; i < n; i++) lwekey_out[i] = (uint32_t)lwekey[i];
    if (ringkey_out) for (int c = 0; c < N; c++) ringkey_out[c] = (T)ringkey[c];
    const int8_t *brk_key = kms ? gswkey.data() : ringkey.data();   // RGSW key: gswkey for KMS*, ringkey for CGGI/LMSS

    // public key b = gen_b(ta, key) (unienc.jl:77-90; keygen.jl:68,100,136)
    if (mk && pubb) {
        ChaCha20 rng(seed, stream_id(S_PUBB, party, 0));
        std::vector<T> prod(N), b(N);
        for (int j = 0; j < p->l_uni; j++) {
            negacyclic_mul_small(prod.data(), crs + (size_t)j * N, ringkey.data(), N);
            for (int c = 0; c < N; c++) b[c] = (T)((T)0 - prod[c]) + noise<T>(rng, p->beta);
            fft_poly(b.data(), (cplx *)pubb + (size_t)j * H, tb);
        }
    }
    // rlk = UniEnc(gswkey poly) under unikey (keygen.jl:103,139)
    if (kms && rlk) {
        ChaCha20 rng(seed, stream_id(S_RLK, party, 0));
        std::vector<T> msg(N);
        for (int c = 0; c < N; c++) msg[c] = (T)gswkey[c];
        unienc_fft<T>(rng, crs, msg.data(), ringkey.data(), p->beta, p->l_uni, p->logB_uni, (cplx *)rlk, tb);
    }
    // brk (keygen.jl:12-14, 39-41, 71-73, 106-108, 143-145)
    if (brk) {
        const size_t per = (p->scheme == MKTFHE_CCS ? (size_t)3 * p->l_uni : (size_t)4 * p->l_gsw) * H;
#pragma omp parallel for schedule(dynamic, 4) num_threads(nthreads)
        for (int i = 0; i < n; i++) {
            ChaCha20 rng(seed, stream_id(S_BRK, party, (uint64_t)i));
            cplx *out = (cplx *)brk + (size_t)i * per;
            if (p->scheme == MKTFHE_CCS) {
                std::vector<T> msg(N, (T)0);
                msg[0] = (T)lwekey[i];
                unienc_fft<T>(rng, crs, msg.data(), ringkey.data(), p->beta, p->l_uni, p->logB_uni, out, tb);
            } else {
                rgsw_fft<T>(rng, (T)lwekey[i], brk_key, p->beta, p->l_gsw, p->logB_gsw, out, tb);
            }
        }
    }
    // ksk[digit, c] = lev_encrypt(key[c] * digit, lwekey, alpha, kskpar) (keygen.jl:16-24, 43-52, 75-79, 110-114, 147-151)
    if (ksk) {
        const int Dk = mktfhe_ksk_rows(p), f = p->f;
        const size_t row = (size_t)n + 1;
#pragma omp parallel for schedule(dynamic, 16) num_threads(nthreads)
        for (int c = 0; c < N; c++) {
            uint32_t *base = ksk + (size_t)c * Dk * f * row;
            if (block && c < n) { memset(base, 0, sizeof(uint32_t) * (size_t)Dk * f * row); continue; }   // never read (undef in the reference)
            ChaCha20 rng(seed, stream_id(S_KSK, party, (uint64_t)c));
            for (int dg = 1; dg <= Dk; dg++)
                for (int lv = 0; lv < f; lv++) {
                    const uint32_t m = (uint32_t)ringkey[c] * (uint32_t)dg * gvec<uint32_t>(lv, p->logD);
                    lwe_sample(rng, lwekey.data(), n, p->alpha, m, base + ((size_t)(dg - 1) * f + lv) * row);
                }
        }
    }
    return 0;
}

template <class T> int crs_impl(const mktfhe_params *p, const ChaCha20::Key &seed, T *coeff, double *fftout) {
    const int N = p->N, H = N / 2;
    const Tables &tb = tables_for(N);
    ChaCha20 rng(seed, stream_id(S_CRS, 0, 0));
    for (size_t i = 0; i < (size_t)p->l_uni * N; i++) coeff[i] = rng.torus<T>();      // randnativepoly (polynomial.jl:48-49)
    if (fftout) for (int j = 0; j < p->l_uni; j++) fft_poly(coeff + (size_t)j * N, (cplx *)fftout + (size_t)j * H, tb);
    return 0;
}

inline uint32_t mu_of(int m) { return ((uint32_t)(2 * (m ? 1 : 0) - 1)) << 29; }   // scheme.jl:356-357: (2m-1) << 29

}  // namespace

extern "C" {

static int crs_any(const mktfhe_params *p, const ChaCha20::Key &seed, void *crs_coeff, double *crs_fft) {
    if (p->scheme != MKTFHE_CCS && p->scheme != MKTFHE_KMS && p->scheme != MKTFHE_KMS_BLOCK) return -1;
    return mktfhe_torus_bits(p) == 64 ? crs_impl<uint64_t>(p, seed, (uint64_t *)crs_coeff, crs_fft)
                                      : crs_impl<uint32_t>(p, seed, (uint32_t *)crs_coeff, crs_fft);
}
int mktfhe_host_crs(const mktfhe_params *p, uint64_t seed, void *crs_coeff, double *crs_fft) {
    return crs_any(p, ChaCha20::Key::from_seed(seed), crs_coeff, crs_fft);
}
int mktfhe_host_crs_key(const mktfhe_params *p, const uint8_t *key32, void *crs_coeff, double *crs_fft) {
    if (!key32) return -3;
    return crs_any(p, ChaCha20::Key::from_bytes(key32), crs_coeff, crs_fft);
}

static int party_keygen_any(const mktfhe_params *p, const ChaCha20::Key &seed, int party, const void *crs_coeff, uint32_t *lwekey,
                            void *ringkey, double *brk, double *rlk, double *pubb, uint32_t *ksk, int nthreads);
int mktfhe_host_party_keygen(const mktfhe_params *p, uint64_t seed, int party, const void *crs_coeff, uint32_t *lwekey,
                             void *ringkey, double *brk, double *rlk, double *pubb, uint32_t *ksk, int nthreads) {
    return party_keygen_any(p, ChaCha20::Key::from_seed(seed), party, crs_coeff, lwekey, ringkey, brk, rlk, pubb, ksk, nthreads);
}
int mktfhe_host_party_keygen_key(const mktfhe_params *p, const uint8_t *key32, int party, const void *crs_coeff, uint32_t *lwekey,
                                 void *ringkey, double *brk, double *rlk, double *pubb, uint32_t *ksk, int nthreads) {
    if (!key32) return -3;
    return party_keygen_any(p, ChaCha20::Key::from_bytes(key32), party, crs_coeff, lwekey, ringkey, brk, rlk, pubb, ksk, nthreads);
}
static int party_keygen_any(const mktfhe_params *p, const ChaCha20::Key &seed, int party, const void *crs_coeff, uint32_t *lwekey,
                            void *ringkey, double *brk, double *rlk, double *pubb, uint32_t *ksk, int nthreads) {
    if (!lwekey) return -3;
    return mktfhe_torus_bits(p) == 64
               ? party_keygen_impl<uint64_t>(p, seed, party, (const uint64_t *)crs_coeff, lwekey, (uint64_t *)ringkey, brk, rlk, pubb, ksk, nthreads)
               : party_keygen_impl<uint32_t>(p, seed, party, (const uint32_t *)crs_coeff, lwekey, (uint32_t *)ringkey, brk, rlk, pubb, ksk, nthreads);
}

static int enc_any(const mktfhe_params *p, const ChaCha20::Key &seed, int m, const uint32_t *lwekey, uint32_t *out, uint64_t nonce = 0);
static int enc_ith_any(const mktfhe_params *p, const ChaCha20::Key &seed, int m, int i, const uint32_t *lwekey_i, uint32_t *out, uint64_t nonce = 0);
static int enc_full_any(const mktfhe_params *p, const ChaCha20::Key &seed, int m, const uint32_t *lwekeys, uint32_t *out, uint64_t nonce = 0);
int mktfhe_host_lwe_encrypt(const mktfhe_params *p, uint64_t seed, int m, const uint32_t *lwekey, uint32_t *out) {
    return enc_any(p, ChaCha20::Key::from_seed(seed), m, lwekey, out);
}
int mktfhe_host_lwe_ith_encrypt(const mktfhe_params *p, uint64_t seed, int m, int i, const uint32_t *lwekey_i, uint32_t *out) {
    return enc_ith_any(p, ChaCha20::Key::from_seed(seed), m, i, lwekey_i, out);
}
int mktfhe_host_lwe_encrypt_full(const mktfhe_params *p, uint64_t seed, int m, const uint32_t *lwekeys, uint32_t *out) {
    return enc_full_any(p, ChaCha20::Key::from_seed(seed), m, lwekeys, out);
}
int mktfhe_host_lwe_encrypt_key(const mktfhe_params *p, const uint8_t *key32, int m, const uint32_t *lwekey, uint32_t *out) {
    return key32 ? enc_any(p, ChaCha20::Key::from_bytes(key32), m, lwekey, out) : -3;
}
int mktfhe_host_lwe_ith_encrypt_key(const mktfhe_params *p, const uint8_t *key32, int m, int i, const uint32_t *lwekey_i, uint32_t *out) {
    return key32 ? enc_ith_any(p, ChaCha20::Key::from_bytes(key32), m, i, lwekey_i, out) : -3;
}
int mktfhe_host_lwe_encrypt_full_key(const mktfhe_params *p, const uint8_t *key32, int m, const uint32_t *lwekeys, uint32_t *out) {
    return key32 ? enc_full_any(p, ChaCha20::Key::from_bytes(key32), m, lwekeys, out) : -3;
}
static int enc_any(const mktfhe_params *p, const ChaCha20::Key &seed, int m, const uint32_t *lwekey, uint32_t *out, uint64_t nonce) {
    ChaCha20 rng(seed, stream_id(S_ENC, 0, 0 | (nonce << 2)));
    uint32_t dot = 0;
    for (int i = 0; i < p->n; i++) { out[1 + i] = rng.u32(); dot += out[1 + i] * lwekey[i]; }
    out[0] = noise<uint32_t>(rng, p->alpha) + ((0u - dot) + mu_of(m));
    return 0;
}

static int enc_ith_any(const mktfhe_params *p, const ChaCha20::Key &seed, int m, int i, const uint32_t *lwekey_i, uint32_t *out, uint64_t nonce) {
    if (i < 0 || i >= p->k) return -1;
    ChaCha20 rng(seed, stream_id(S_ENC, i, 1 | (nonce << 2)));
    memset(out, 0, sizeof(uint32_t) * mktfhe_lwe_words(p));
    uint32_t *a = out + 1 + (size_t)i * p->n, dot = 0;
    for (int j = 0; j < p->n; j++) { a[j] = rng.u32(); dot += a[j] * lwekey_i[j]; }
    out[0] = noise<uint32_t>(rng, p->alpha) + ((0u - dot) + mu_of(m));
    return 0;
}

static int enc_full_any(const mktfhe_params *p, const ChaCha20::Key &seed, int m, const uint32_t *lwekeys, uint32_t *out, uint64_t nonce) {
    ChaCha20 rng(seed, stream_id(S_ENC, 0, 2 | (nonce << 2)));
    uint32_t dot = 0;
    const size_t len = (size_t)p->n * p->k;
    for (size_t j = 0; j < len; j++) { out[1 + j] = rng.u32(); dot += out[1 + j] * lwekeys[j]; }
    out[0] = noise<uint32_t>(rng, p->alpha) + ((0u - dot) + mu_of(m));
    return 0;
}

uint32_t mktfhe_host_lwe_phase(const mktfhe_params *p, const uint32_t *lwekeys, const uint32_t *ct) {
    uint32_t ph = ct[0];
    const size_t len = (size_t)p->n * p->k;
    for (size_t j = 0; j < len; j++) ph += ct[1 + j] * lwekeys[j];
    return ph;
}

int mktfhe_host_lwe_decrypt(const mktfhe_params *p, const uint32_t *lwekeys, const uint32_t *ct) {
    const uint32_t ph = mktfhe_host_lwe_phase(p, lwekeys, ct);
    if (p->scheme == MKTFHE_CGGI || p->scheme == MKTFHE_LMSS) {
        // divbits(phase, 29) == 1   (scheme.jl:388-389, arithmetic.jl:23-27)
        const uint32_t carry = (uint32_t)(ph << 3) >> 31;
        return ((ph >> 29) + carry) == 1u;
    }
    return ph < 0x80000000u;   // scheme.jl:391-407
}

// ---- batched forms (scheme.jl:352-407 over many ciphertexts; OpenMP over ciphertexts)
static int enc_batch_any(const mktfhe_params *p, bool keyed, uint64_t seed0, const uint8_t *key32, int kind, int party, const uint8_t *bits,
                         size_t count, const uint32_t *lwekeys, uint32_t *out, int nthreads) {
    if (!bits || !lwekeys || !out || kind < 0 || kind > 2) return -3;
    const bool mk = p->scheme == MKTFHE_CCS || p->scheme == MKTFHE_KMS || p->scheme == MKTFHE_KMS_BLOCK;
    if ((kind == 0) == mk) return -1;                       // lwe_encrypt is single-key, the other two multi-key
    if (kind == 1 && (party < 0 || party >= p->k)) return -1;
    const size_t lw = mktfhe_lwe_words(p);
    const ChaCha20::Key k0 = keyed ? ChaCha20::Key::from_bytes(key32) : ChaCha20::Key::from_seed(seed0);
#ifdef _OPENMP
    if (nthreads <= 0) nthreads = omp_get_max_threads();
#pragma omp parallel for num_threads(nthreads) schedule(static)
#endif
    for (long long g = 0; g < (long long)count; g++) {
        // seeded form: ciphertext g is exactly the single call with seed0 + g; keyed form: one key, nonce g
        const ChaCha20::Key kg = keyed ? k0 : ChaCha20::Key::from_seed(seed0 + (uint64_t)g);
        const uint64_t nonce = keyed ? (uint64_t)g + 1 : 0;
        uint32_t *o = out + (size_t)g * lw;
        if (kind == 0) enc_any(p, kg, bits[g], lwekeys, o, nonce);
        else if (kind == 1) enc_ith_any(p, kg, bits[g], party, lwekeys, o, nonce);
        else enc_full_any(p, kg, bits[g], lwekeys, o, nonce);
    }
    return 0;
}
int mktfhe_host_encrypt_batch(const mktfhe_params *p, uint64_t seed0, int kind, int party, const uint8_t *bits, size_t count,
                              const uint32_t *lwekeys, uint32_t *out, int nthreads) {
    return enc_batch_any(p, false, seed0, nullptr, kind, party, bits, count, lwekeys, out, nthreads);
}
int mktfhe_host_encrypt_batch_key(const mktfhe_params *p, const uint8_t *key32, int kind, int party, const uint8_t *bits, size_t count,
                                  const uint32_t *lwekeys, uint32_t *out, int nthreads) {
    if (!key32) return -3;
    return enc_batch_any(p, true, 0, key32, kind, party, bits, count, lwekeys, out, nthreads);
}
int mktfhe_host_phase_batch(const mktfhe_params *p, const uint32_t *lwekeys, const uint32_t *cts, size_t count, uint32_t *phases_out, int nthreads) {
    if (!lwekeys || !cts || !phases_out) return -3;
    const size_t lw = mktfhe_lwe_words(p);
#ifdef _OPENMP
    if (nthreads <= 0) nthreads = omp_get_max_threads();
#pragma omp parallel for num_threads(nthreads) schedule(static)
#endif
    for (long long g = 0; g < (long long)count; g++) phases_out[g] = mktfhe_host_lwe_phase(p, lwekeys, cts + (size_t)g * lw);
    return 0;
}
int mktfhe_host_decrypt_batch(const mktfhe_params *p, const uint32_t *lwekeys, const uint32_t *cts, size_t count, uint8_t *bits_out, int nthreads) {
    if (!lwekeys || !cts || !bits_out) return -3;
    const size_t lw = mktfhe_lwe_words(p);
#ifdef _OPENMP
    if (nthreads <= 0) nthreads = omp_get_max_threads();
#pragma omp parallel for num_threads(nthreads) schedule(static)
#endif
    for (long long g = 0; g < (long long)count; g++) bits_out[g] = (uint8_t)mktfhe_host_lwe_decrypt(p, lwekeys, cts + (size_t)g * lw);
    return 0;
}

void mktfhe_host_fft_tables(int N, double *psi, double *psiinv, double *roots, double *rootsinv) {
    const Tables &t = tables_for(N);
    memcpy(psi, t.psi.data(), sizeof(cplx) * t.H);
    memcpy(psiinv, t.psiinv.data(), sizeof(cplx) * t.H);
    memcpy(roots, t.roots.data(), sizeof(cplx) * t.H);
    memcpy(rootsinv, t.rootsinv.data(), sizeof(cplx) * t.H);
}

}  // extern "C"
