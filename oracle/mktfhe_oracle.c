/* mktfhe_oracle.c -- CPU oracle (TEST INFRASTRUCTURE ONLY; see mktfhe_oracle.h).
 *
 * Restates /root/reference/src/{ring,ciphertext,tfhe} for the Float64 hot path.
 * Build: see oracle/Makefile (-O2 -ffp-contract=off -fopenmp).
 */
#include "mktfhe_oracle.h"
#include <math.h>
#include <quadmath.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define ORC_MAXL 32

typedef struct { double re, im; } cplx;

/* Julia Complex{Float64} `*` and `+`/`-` (base/complex.jl): no fma. */
static inline cplx cmul(cplx z, cplx w) {
    cplx r;
    r.re = z.re * w.re - z.im * w.im;
    r.im = z.re * w.im + z.im * w.re;
    return r;
}
static inline cplx cadd(cplx a, cplx b) { cplx r = { a.re + b.re, a.im + b.im }; return r; }
static inline cplx csub(cplx a, cplx b) { cplx r = { a.re - b.re, a.im - b.im }; return r; }

typedef struct {
    int N, H;
    cplx *psi, *psiinv, *roots, *rootsinv;
} fft_tab;

/* src/ring/fft.jl:1-15 */
static void bit_reverse(cplx *v, int n) {
    int j = 0;
    for (int i = 1; i < n; i++) {
        int bit = n >> 1;
        while (j >= bit) { j -= bit; bit >>= 1; }
        j += bit;
        if (i < j) { cplx t = v[i]; v[i] = v[j]; v[j] = t; }
    }
}

/* src/ring/fft.jl:26-44.  The reference evaluates exp() in BigFloat and rounds to Float64;
 * binary128 (113-bit) sincos rounded to double gives the same correctly rounded values. */
static void fft_init(fft_tab *F, int N) {
    const int H = N >> 1;
    F->N = N; F->H = H;
    F->psi = malloc(sizeof(cplx) * H); F->psiinv = malloc(sizeof(cplx) * H);
    F->roots = malloc(sizeof(cplx) * H); F->rootsinv = malloc(sizeof(cplx) * H);
    for (int j = 0; j < H; j++) {
        __float128 th = M_PIq * (__float128)j / (__float128)H;
        F->psi[j].re = (double)cosq(th);  F->psi[j].im = (double)(-sinq(th));
        F->psiinv[j].re = (double)cosq(th); F->psiinv[j].im = (double)sinq(th);
        __float128 ph = M_PIq * (__float128)j / (__float128)N;
        F->roots[j].re = (double)cosq(ph); F->roots[j].im = (double)sinq(ph);
        F->rootsinv[j].re = (double)(cosq(ph) / (__float128)H);
        F->rootsinv[j].im = (double)(-sinq(ph) / (__float128)H);
    }
    bit_reverse(F->psi, H);
    bit_reverse(F->psiinv, H);
}
static void fft_free(fft_tab *F) { free(F->psi); free(F->psiinv); free(F->roots); free(F->rootsinv); }

/* src/ring/fft.jl:105-155 Cooley-Tukey, no reordering.  The x8/x4/x2 bodies there are plain
 * unrolling of this loop nest. */
static void fft_inplace(cplx *a, const cplx *psi, int H) {
    int m = 1, k = H >> 1;
    while (m < H) {
        for (int i = 0; i < m; i++) {
            const cplx w = psi[m + i];
            const int j1 = 2 * i * k;
            for (int j = j1; j < j1 + k; j++) {
                cplx t = a[j], u = cmul(a[j + k], w);
                a[j] = cadd(t, u);
                a[j + k] = csub(t, u);
            }
        }
        m <<= 1; k >>= 1;
    }
}

/* src/ring/fft.jl:159-210 Gentleman-Sande (the second, live `ifft!` method). */
static void ifft_inplace(cplx *a, const cplx *psiinv, int H) {
    int m = H >> 1, k = 1;
    while (m > 0) {
        for (int i = 0; i < m; i++) {
            const cplx w = psiinv[m + i];
            const int j1 = 2 * i * k;
            for (int j = j1; j < j1 + k; j++) {
                cplx t = a[j], u = a[j + k];
                a[j] = cadd(t, u);
                a[j + k] = cmul(csub(t, u), w);
            }
        }
        m >>= 1; k <<= 1;
    }
}

#define CAT_(a, b) a##b
#define CAT(a, b) CAT_(a, b)

#define TORUS uint32_t
#define STORUS int32_t
#define BITS 32
#define SUF(x) CAT(x, _32)
#include "orc_ring.inc"
#undef TORUS
#undef STORUS
#undef BITS
#undef SUF

#define TORUS uint64_t
#define STORUS int64_t
#define BITS 64
#define SUF(x) CAT(x, _64)
#include "orc_ring.inc"
#undef TORUS
#undef STORUS
#undef BITS
#undef SUF

struct orc_ctx {
    orc_params p;
    orc_keys keys;
    fft_tab F;
    cplx *monomial;   /* [2N][H] */
    int bits;         /* RLWE torus width */
    int Dk;           /* ksk rows per coefficient */
};

static inline int is_kms(const orc_ctx *c) { return c->p.scheme == ORC_KMS || c->p.scheme == ORC_KMS_BLOCK; }
static inline int is_block(const orc_ctx *c) { return c->p.scheme == ORC_LMSS || c->p.scheme == ORC_KMS_BLOCK; }

/* src/tfhe/scheme.jl:121-146 */
static void monomials_build(const fft_tab *F, cplx *out) {
    const int N = F->N, H = F->H;
    uint32_t *tmp = calloc(N, sizeof(uint32_t));
    memset(out + (size_t)(2 * N - 1) * H, 0, sizeof(cplx) * H);
    tmp[0] = (uint32_t)-1;
    for (int i = 1; i < N; i++) {
        tmp[i] = 1;
        fftto_32(out + (size_t)(i - 1) * H, tmp, F);
        tmp[i] = 0;
    }
    tmp[0] = (uint32_t)-2;
    fftto_32(out + (size_t)(N - 1) * H, tmp, F);
    tmp[0] = (uint32_t)-1;
    for (int i = 1; i < N; i++) {
        tmp[i] = (uint32_t)-1;
        fftto_32(out + (size_t)(N + i - 1) * H, tmp, F);
        tmp[i] = 0;
    }
    free(tmp);
}

orc_ctx *orc_create(const orc_params *p, const orc_keys *keys) {
    orc_ctx *c = calloc(1, sizeof(orc_ctx));
    c->p = *p;
    c->keys = *keys;
    fft_init(&c->F, p->N);
    c->monomial = malloc(sizeof(cplx) * (size_t)2 * p->N * c->F.H);
    monomials_build(&c->F, c->monomial);
    c->bits = (p->scheme == ORC_KMS || p->scheme == ORC_KMS_BLOCK) ? 64 : 32;
    const int D = 1 << p->logD;
    c->Dk = (p->scheme == ORC_LMSS || p->scheme == ORC_KMS_BLOCK) ? D / 2 : D - 1;
    return c;
}

void orc_destroy(orc_ctx *c) {
    if (!c) return;
    fft_free(&c->F);
    free(c->monomial);
    free(c);
}

void orc_fft_tables(int N, double *psi, double *psiinv, double *roots, double *rootsinv) {
    fft_tab F; fft_init(&F, N);
    memcpy(psi, F.psi, sizeof(cplx) * F.H);
    memcpy(psiinv, F.psiinv, sizeof(cplx) * F.H);
    memcpy(roots, F.roots, sizeof(cplx) * F.H);
    memcpy(rootsinv, F.rootsinv, sizeof(cplx) * F.H);
    fft_free(&F);
}

void orc_monomials(int N, double *out) {
    fft_tab F; fft_init(&F, N);
    monomials_build(&F, (cplx *)out);
    fft_free(&F);
}

void orc_fft(int N, int bits, const void *poly, double *out) {
    fft_tab F; fft_init(&F, N);
    if (bits == 32) fftto_32((cplx *)out, poly, &F); else fftto_64((cplx *)out, poly, &F);
    fft_free(&F);
}

void orc_ifft(int N, int bits, double *in_consumed, void *poly) {
    fft_tab F; fft_init(&F, N);
    if (bits == 32) ifftto_32(poly, (cplx *)in_consumed, &F); else ifftto_64(poly, (cplx *)in_consumed, &F);
    fft_free(&F);
}

void orc_decomp(int N, int bits, int l, int logB, const void *poly, void *digits) {
    if (bits == 32) {
        uint32_t *ptr[ORC_MAXL];
        for (int j = 0; j < l; j++) ptr[j] = (uint32_t *)digits + (size_t)j * N;
        decomp_poly_32(ptr, poly, N, l, logB);
    } else {
        uint64_t *ptr[ORC_MAXL];
        for (int j = 0; j < l; j++) ptr[j] = (uint64_t *)digits + (size_t)j * N;
        decomp_poly_64(ptr, poly, N, l, logB);
    }
}

/* ------------------------------------------------------------------ gates, modswitch */

/* src/tfhe/gate.jl:1-58 (linear parts; UInt32 wraparound; Julia shift binds tighter than +/-). */
void orc_gate_linear(const orc_ctx *c, int op, const uint32_t *x, const uint32_t *y, uint32_t *out) {
    const int len = 1 + c->p.n * c->p.k;
    uint32_t cst; int mode;      /* mode 0: -x-y ; 1: x+y ; 2: 2(x+y) ; 3: -2(x+y) */
    switch (op) {
    case ORC_NAND: cst = 1u << 29; mode = 0; break;
    case ORC_AND:  cst = 7u << 29; mode = 1; break;
    case ORC_OR:   cst = 1u << 29; mode = 1; break;
    case ORC_XOR:  cst = 1u << 30; mode = 2; break;
    case ORC_XNOR: cst = 3u << 30; mode = 3; break;
    default:       cst = 7u << 29; mode = 0; break;   /* NOR */
    }
    for (int i = 0; i < len; i++) {
        uint32_t s = x[i] + y[i], v;
        switch (mode) {
        case 0: v = 0u - s; break;
        case 1: v = s; break;
        case 2: v = 2u * s; break;
        default: v = 0u - 2u * s; break;
        }
        out[i] = (i == 0) ? cst + v : v;
    }
}

/* src/tfhe/bootstrapping.jl:8-9 */
void orc_modswitch(const orc_ctx *c, const uint32_t *lwe, uint32_t *tilde) {
    const int len = 1 + c->p.n * c->p.k;
    int logN = 0; while ((1 << logN) < c->p.N) logN++;
    for (int i = 0; i < len; i++) tilde[i] = divbits_32(lwe[i], 32 - logN - 1);
}

/* ------------------------------------------------------------------ single-key blind rotation */

static inline const cplx *brk_rgsw(const orc_ctx *c, int party, int idx) {
    return (const cplx *)c->keys.brk[party] + (size_t)idx * 4 * c->p.l_gsw * c->F.H;
}
static inline const cplx *mono(const orc_ctx *c, uint32_t a) { return c->monomial + (size_t)(a - 1) * c->F.H; }

/* CGGI: bootstrapping.jl:32-76 ; LMSS: :114-165 (k = 1). */
static void blindrotate_sk(const orc_ctx *c, const uint32_t *ta, uint32_t *acc /* [2][N] */) {
    const int N = c->p.N, H = c->F.H, l = c->p.l_gsw, logB = c->p.logB_gsw;
    cplx *tvec = malloc(sizeof(cplx) * 2 * l * H), *tacc = malloc(sizeof(cplx) * 2 * H), *tacc2 = malloc(sizeof(cplx) * 2 * H);
    uint32_t *scr = malloc(sizeof(uint32_t) * (size_t)(l + 1) * N), *tmp = scr + (size_t)l * N;
    if (c->p.scheme == ORC_CGGI) {
        for (int idx = 0; idx < c->p.n; idx++) {
            if (ta[idx] == 0) continue;
            decomp_fft_row_32(tvec, acc, acc + N, l, logB, &c->F, scr);
            rgsw_mac_32(tacc, tvec, brk_rgsw(c, 0, idx), l, H);
            mono_ifft_add_32(acc, acc + N, tacc, mono(c, ta[idx]), &c->F, tmp);
        }
    } else {
        for (int i1 = 0; i1 < c->p.d; i1++) {
            decomp_fft_row_32(tvec, acc, acc + N, l, logB, &c->F, scr);
            memset(tacc2, 0, sizeof(cplx) * 2 * H);
            for (int i2 = 0; i2 < c->p.ell; i2++) {
                const int idx = i1 * c->p.ell + i2;
                if (ta[idx] == 0) continue;
                rgsw_mac_32(tacc, tvec, brk_rgsw(c, 0, idx), l, H);
                const cplx *mo = mono(c, ta[idx]);
                for (int s = 0; s < H; s++) {   /* muladdto!(tacc2, monomial, tacc)  :157 */
                    tacc2[s] = cadd(tacc2[s], cmul(mo[s], tacc[s]));
                    tacc2[H + s] = cadd(tacc2[H + s], cmul(mo[s], tacc[H + s]));
                }
            }
            mono_ifft_add_32(acc, acc + N, tacc2, NULL, &c->F, tmp);
        }
    }
    free(tvec); free(tacc); free(tacc2); free(scr);
}

/* ------------------------------------------------------------------ CCS blind rotation */

/* bootstrapping.jl:234-328.  acc: [(k+1)][N] uint32. */
static void blindrotate_ccs(const orc_ctx *c, const uint32_t *ta, uint32_t *acc) {
    const int N = c->p.N, H = c->F.H, k = c->p.k, n = c->p.n, l = c->p.l_uni, logB = c->p.logB_uni;
    const size_t lH = (size_t)l * H;
    cplx *tbvec = malloc(sizeof(cplx) * lH), *tavec = malloc(sizeof(cplx) * lH * k);
    cplx *tv0 = malloc(sizeof(cplx) * H), *tv = malloc(sizeof(cplx) * (size_t)H * k);
    cplx *tv0vec = malloc(sizeof(cplx) * lH), *tvvec = malloc(sizeof(cplx) * lH * k);
    cplx *tacc = malloc(sizeof(cplx) * (size_t)H * (k + 1));
    uint32_t *v0 = malloc(sizeof(uint32_t) * N), *v = malloc(sizeof(uint32_t) * (size_t)N * k);
    uint32_t *scr = malloc(sizeof(uint32_t) * (size_t)(l + 1) * N), *tmp = scr + (size_t)l * N;
    const cplx *crs = (const cplx *)c->keys.crs;

    for (int idx = 0; idx < k; idx++) {
        for (int i = 0; i < n; i++) {
            const uint32_t at = ta[(size_t)idx * n + i];   /* reshape(tildeavec, n, k)[i, idx] */
            if (at == 0) continue;
            const cplx *uni = (const cplx *)c->keys.brk[idx] + (size_t)i * 3 * lH;  /* [j][d,f.b,f.a][H] */
            const int na = idx + 1;                          /* components a_1 .. a_idx (1-based idx) */
            decomp_fft_32(tbvec, acc, l, logB, &c->F, scr);
            for (int j1 = 0; j1 < na; j1++) decomp_fft_32(tavec + j1 * lH, acc + (size_t)(1 + j1) * N, l, logB, &c->F, scr);
            /* u : :277-284 */
            memset(tacc, 0, sizeof(cplx) * (size_t)H * (k + 1));
            for (int j = 0; j < l; j++) {
                const cplx *dj = uni + (size_t)(j * 3 + 0) * H;
                for (int s = 0; s < H; s++) tacc[s] = cadd(tacc[s], cmul(tbvec[j * H + s], dj[s]));
            }
            for (int j1 = 0; j1 < na; j1++) for (int j2 = 0; j2 < l; j2++) {
                const cplx *dj = uni + (size_t)(j2 * 3 + 0) * H;
                cplx *ta_ = tacc + (size_t)(1 + j1) * H;
                for (int s = 0; s < H; s++) ta_[s] = cadd(ta_[s], cmul(tavec[j1 * lH + j2 * H + s], dj[s]));
            }
            /* v : :286-294 */
            memset(tv, 0, sizeof(cplx) * (size_t)H * k);
            memset(tv0, 0, sizeof(cplx) * H);
            for (int j = 0; j < l; j++)
                for (int s = 0; s < H; s++) tv0[s] = csub(tv0[s], cmul(tbvec[j * H + s], crs[j * H + s]));
            for (int j1 = 0; j1 < na; j1++) for (int j2 = 0; j2 < l; j2++) {
                const cplx *pb = (const cplx *)c->keys.pubb[j1] + (size_t)j2 * H;
                for (int s = 0; s < H; s++) tv[j1 * H + s] = cadd(tv[j1 * H + s], cmul(tavec[j1 * lH + j2 * H + s], pb[s]));
            }
            ifftto_32(v0, tv0, &c->F);
            for (int j = 0; j < na; j++) ifftto_32(v + (size_t)j * N, tv + (size_t)j * H, &c->F);
            decomp_fft_32(tv0vec, v0, l, logB, &c->F, scr);
            for (int j1 = 0; j1 < na; j1++) decomp_fft_32(tvvec + j1 * lH, v + (size_t)j1 * N, l, logB, &c->F, scr);
            /* w : :313-320 */
            cplx *tb_ = tacc, *tai = tacc + (size_t)(1 + idx) * H;
            for (int j = 0; j < l; j++) {
                const cplx *fb = uni + (size_t)(j * 3 + 1) * H, *fa = uni + (size_t)(j * 3 + 2) * H;
                for (int s = 0; s < H; s++) tb_[s] = cadd(tb_[s], cmul(tv0vec[j * H + s], fb[s]));
                for (int s = 0; s < H; s++) tai[s] = cadd(tai[s], cmul(tv0vec[j * H + s], fa[s]));
            }
            for (int j1 = 0; j1 < na; j1++) for (int j2 = 0; j2 < l; j2++) {
                const cplx *fb = uni + (size_t)(j2 * 3 + 1) * H, *fa = uni + (size_t)(j2 * 3 + 2) * H;
                const cplx *x = tvvec + j1 * lH + (size_t)j2 * H;
                for (int s = 0; s < H; s++) tb_[s] = cadd(tb_[s], cmul(x[s], fb[s]));
                for (int s = 0; s < H; s++) tai[s] = cadd(tai[s], cmul(x[s], fa[s]));
            }
            /* :322-324  mul!(monomial, tacc); ifftto!(acc2, tacc); add!(acc, acc2) -- all k+1 components */
            const cplx *mo = mono(c, at);
            for (int q = 0; q <= k; q++) {
                cplx *t = tacc + (size_t)q * H;
                for (int s = 0; s < H; s++) t[s] = cmul(mo[s], t[s]);
                ifftto_32(tmp, t, &c->F);
                uint32_t *a = acc + (size_t)q * N;
                for (int cc = 0; cc < N; cc++) a[cc] += tmp[cc];
            }
        }
    }
    free(tbvec); free(tavec); free(tv0); free(tv); free(tv0vec); free(tvvec); free(tacc); free(v0); free(v); free(scr);
}

/* ------------------------------------------------------------------ KMS phase 1 / phase 2 */

void orc_cmux_step(const orc_ctx *c, int party, int idx, uint32_t atilde, void *acc_row) {
    const int N = c->p.N, H = c->F.H, l = c->p.l_gsw, logB = c->p.logB_gsw;
    if (atilde == 0) return;
    cplx *tvec = malloc(sizeof(cplx) * 2 * l * H), *tacc = malloc(sizeof(cplx) * 2 * H);
    if (c->bits == 32) {
        uint32_t *acc = acc_row, *scr = malloc(sizeof(uint32_t) * (size_t)(l + 1) * N);
        decomp_fft_row_32(tvec, acc, acc + N, l, logB, &c->F, scr);
        rgsw_mac_32(tacc, tvec, brk_rgsw(c, party, idx), l, H);
        mono_ifft_add_32(acc, acc + N, tacc, mono(c, atilde), &c->F, scr + (size_t)l * N);
        free(scr);
    } else {
        uint64_t *acc = acc_row, *scr = malloc(sizeof(uint64_t) * (size_t)(l + 1) * N);
        decomp_fft_row_64(tvec, acc, acc + N, l, logB, &c->F, scr);
        rgsw_mac_64(tacc, tvec, brk_rgsw(c, party, idx), l, H);
        mono_ifft_add_64(acc, acc + N, tacc, mono(c, atilde), &c->F, scr + (size_t)l * N);
        free(scr);
    }
    free(tvec); free(tacc);
}

/* One block iteration (idx1) of the LMSS / KMS_block loop on one RLWE row: bootstrapping.jl:124-163 / :624-655.
 * at: ell rotations of the block's key bits. */
void orc_block_step(const orc_ctx *c, int party, int blk, const uint32_t *at, void *acc_row) {
    const int N = c->p.N, H = c->F.H, l = c->p.l_gsw, logB = c->p.logB_gsw, ell = c->p.ell;
    cplx *tvec = malloc(sizeof(cplx) * 2 * l * H), *t1 = malloc(sizeof(cplx) * 2 * H), *t2 = calloc(2 * H, sizeof(cplx));
    if (c->bits == 32) {
        uint32_t *acc = acc_row, *scr = malloc(sizeof(uint32_t) * (size_t)(l + 1) * N);
        decomp_fft_row_32(tvec, acc, acc + N, l, logB, &c->F, scr);
        for (int i2 = 0; i2 < ell; i2++) {
            if (at[i2] == 0) continue;
            rgsw_mac_32(t1, tvec, brk_rgsw(c, party, blk * ell + i2), l, H);
            const cplx *mo = mono(c, at[i2]);
            for (int s = 0; s < 2 * H; s++) t2[s] = cadd(t2[s], cmul(mo[s % H], t1[s]));
        }
        mono_ifft_add_32(acc, acc + N, t2, NULL, &c->F, scr + (size_t)l * N);
        free(scr);
    } else {
        uint64_t *acc = acc_row, *scr = malloc(sizeof(uint64_t) * (size_t)(l + 1) * N);
        decomp_fft_row_64(tvec, acc, acc + N, l, logB, &c->F, scr);
        for (int i2 = 0; i2 < ell; i2++) {
            if (at[i2] == 0) continue;
            rgsw_mac_64(t1, tvec, brk_rgsw(c, party, blk * ell + i2), l, H);
            const cplx *mo = mono(c, at[i2]);
            for (int s = 0; s < 2 * H; s++) t2[s] = cadd(t2[s], cmul(mo[s % H], t1[s]));
        }
        mono_ifft_add_64(acc, acc + N, t2, NULL, &c->F, scr + (size_t)l * N);
        free(scr);
    }
    free(tvec); free(t1); free(t2);
}

/* KMS: bootstrapping.jl:389-443 ; KMS_BLOCK: :599-659 */
void orc_phase1(const orc_ctx *c, int party, const uint32_t *ta, double *levkey_out) {
    const int N = c->p.N, H = c->F.H, l = c->p.l_gsw, logB = c->p.logB_gsw;
    const int iter = party == 0 ? 1 : c->p.l_lev;
    uint64_t *acc = calloc((size_t)iter * 2 * N, sizeof(uint64_t));
    for (int r = 0; r < iter; r++) acc[(size_t)r * 2 * N] = (uint64_t)1 << (64 - (r + 1) * c->p.logB_lev);  /* levpar.gvec[r] */
    cplx *tvec = malloc(sizeof(cplx) * 2 * l * H);
    cplx *tacc = malloc(sizeof(cplx) * (size_t)iter * 2 * H), *tacc2 = malloc(sizeof(cplx) * (size_t)iter * 2 * H);
    uint64_t *scr = malloc(sizeof(uint64_t) * (size_t)(l + 1) * N), *tmp = scr + (size_t)l * N;

    if (c->p.scheme == ORC_KMS) {
        for (int idx = 0; idx < c->p.n; idx++) {
            if (ta[idx] == 0) continue;
            for (int r = 0; r < iter; r++) {
                uint64_t *row = acc + (size_t)r * 2 * N;
                decomp_fft_row_64(tvec, row, row + N, l, logB, &c->F, scr);
                rgsw_mac_64(tacc + (size_t)r * 2 * H, tvec, brk_rgsw(c, party, idx), l, H);
            }
            for (int r = 0; r < iter; r++) {
                uint64_t *row = acc + (size_t)r * 2 * N;
                mono_ifft_add_64(row, row + N, tacc + (size_t)r * 2 * H, mono(c, ta[idx]), &c->F, tmp);
            }
        }
    } else {
        for (int i1 = 0; i1 < c->p.d; i1++) {
            for (int r = 0; r < iter; r++) {
                uint64_t *row = acc + (size_t)r * 2 * N;
                cplx *t1 = tacc + (size_t)r * 2 * H, *t2 = tacc2 + (size_t)r * 2 * H;
                decomp_fft_row_64(tvec, row, row + N, l, logB, &c->F, scr);
                memset(t2, 0, sizeof(cplx) * 2 * H);
                for (int i2 = 0; i2 < c->p.ell; i2++) {
                    const int idx = i1 * c->p.ell + i2;
                    if (ta[idx] == 0) continue;
                    rgsw_mac_64(t1, tvec, brk_rgsw(c, party, idx), l, H);
                    const cplx *mo = mono(c, ta[idx]);
                    for (int s = 0; s < H; s++) {
                        t2[s] = cadd(t2[s], cmul(mo[s], t1[s]));
                        t2[H + s] = cadd(t2[H + s], cmul(mo[s], t1[H + s]));
                    }
                }
            }
            for (int r = 0; r < iter; r++) {
                uint64_t *row = acc + (size_t)r * 2 * N;
                mono_ifft_add_64(row, row + N, tacc2 + (size_t)r * 2 * H, NULL, &c->F, tmp);
            }
        }
    }
    /* fftto!(tacc, acc, ffter) :441 / :657 */
    cplx *out = (cplx *)levkey_out;
    for (int r = 0; r < iter; r++) {
        fftto_64(out + (size_t)(r * 2 + 0) * H, acc + (size_t)r * 2 * N, &c->F);
        fftto_64(out + (size_t)(r * 2 + 1) * H, acc + (size_t)r * 2 * N + N, &c->F);
    }
    free(acc); free(tvec); free(tacc); free(tacc2); free(scr);
}

/* src/tfhe/bootstrapping.jl:11-23: test vector for b~ into acc.b; acc.a = 0. */
static void testvector_64(uint64_t *acc, int N, int k, uint32_t tb) {
    const uint64_t e = (uint64_t)1 << 61;
    memset(acc, 0, sizeof(uint64_t) * (size_t)(k + 1) * N);
    if (tb <= (uint32_t)N) for (int i = 1; i <= N; i++) acc[i - 1] = (uint32_t)i <= tb ? e : (uint64_t)0 - e;
    else { tb -= N; for (int i = 1; i <= N; i++) acc[i - 1] = (uint32_t)i <= tb ? (uint64_t)0 - e : e; }
}
static void testvector_32(uint32_t *acc, int N, int k, uint32_t tb) {
    const uint32_t e = 1u << 29;
    memset(acc, 0, sizeof(uint32_t) * (size_t)(k + 1) * N);
    if (tb <= (uint32_t)N) for (int i = 1; i <= N; i++) acc[i - 1] = (uint32_t)i <= tb ? e : 0u - e;
    else { tb -= N; for (int i = 1; i <= N; i++) acc[i - 1] = (uint32_t)i <= tb ? 0u - e : e; }
}

/* bootstrapping.jl:448-558.  levkeys: [k][l_lev][2][H]; acc: [(k+1)][N] uint64, holds the test vector on entry. */
static void phase2_64(const orc_ctx *c, const cplx *levkeys, uint64_t *acc) {
    const int N = c->p.N, H = c->F.H, k = c->p.k;
    const int ll = c->p.l_lev, lbl = c->p.logB_lev, lu = c->p.l_uni, lbu = c->p.logB_uni;
    const int maxl = ll > lu ? ll : lu;
    const size_t mH = (size_t)maxl * H;
    cplx *tbvec = malloc(sizeof(cplx) * mH), *tavec = malloc(sizeof(cplx) * mH * k);
    cplx *tv = malloc(sizeof(cplx) * H), *tvvec = malloc(sizeof(cplx) * (size_t)lu * H);
    cplx *tx = malloc(sizeof(cplx) * (size_t)H * (k + 1)), *ty = malloc(sizeof(cplx) * (size_t)H * (k + 1));
    uint64_t *y = malloc(sizeof(uint64_t) * (size_t)N * (k + 1)), *v = malloc(sizeof(uint64_t) * N);
    uint64_t *scr = malloc(sizeof(uint64_t) * (size_t)(maxl + 1) * N);
    const cplx *crs = (const cplx *)c->keys.crs;

    for (int idx = 0; idx < k; idx++) {            /* idx here = reference idx - 1 ; a_1..a_idx are live */
        const cplx *lk = levkeys + (size_t)idx * ll * 2 * H;
        const cplx *rlk = (const cplx *)c->keys.rlk[idx];
        decomp_fft_64(tbvec, acc, ll, lbl, &c->F, scr);
        for (int i = 0; i < idx; i++) decomp_fft_64(tavec + i * mH, acc + (size_t)(1 + i) * N, ll, lbl, &c->F, scr);
        const int iter = idx == 0 ? 1 : ll;
        /* :483-499 */
        memset(tx, 0, sizeof(cplx) * (size_t)H * (k + 1));
        memset(ty, 0, sizeof(cplx) * (size_t)H * (k + 1));
        for (int i = 0; i < iter; i++) {
            const cplx *kb = lk + (size_t)(i * 2 + 0) * H;
            for (int s = 0; s < H; s++) tx[s] = cadd(tx[s], cmul(tbvec[i * H + s], kb[s]));
        }
        for (int i = 0; i < idx; i++) for (int j = 0; j < iter; j++) {
            const cplx *kb = lk + (size_t)(j * 2 + 0) * H;
            cplx *t = tx + (size_t)(1 + i) * H;
            for (int s = 0; s < H; s++) t[s] = cadd(t[s], cmul(tavec[i * mH + j * H + s], kb[s]));
        }
        for (int i = 0; i < iter; i++) {
            const cplx *ka = lk + (size_t)(i * 2 + 1) * H;
            for (int s = 0; s < H; s++) ty[s] = cadd(ty[s], cmul(tbvec[i * H + s], ka[s]));
        }
        for (int i = 0; i < idx; i++) for (int j = 0; j < iter; j++) {
            const cplx *ka = lk + (size_t)(j * 2 + 1) * H;
            cplx *t = ty + (size_t)(1 + i) * H;
            for (int s = 0; s < H; s++) t[s] = cadd(t[s], cmul(tavec[i * mH + j * H + s], ka[s]));
        }
        /* :501-504 */
        ifftto_64(y, ty, &c->F);
        for (int i = 0; i < idx; i++) ifftto_64(y + (size_t)(1 + i) * N, ty + (size_t)(1 + i) * H, &c->F);
        /* :508-517 */
        decomp_fft_64(tbvec, y, lu, lbu, &c->F, scr);
        for (int i = 0; i < idx; i++) decomp_fft_64(tavec + i * mH, y + (size_t)(1 + i) * N, lu, lbu, &c->F, scr);
        /* u : :520-526 */
        memset(ty, 0, sizeof(cplx) * (size_t)H * (k + 1));
        for (int i = 0; i < lu; i++) {
            const cplx *dj = rlk + (size_t)(i * 3 + 0) * H;
            for (int s = 0; s < H; s++) ty[s] = cadd(ty[s], cmul(tbvec[i * H + s], dj[s]));
        }
        for (int i = 0; i < idx; i++) for (int j = 0; j < lu; j++) {
            const cplx *dj = rlk + (size_t)(j * 3 + 0) * H;
            cplx *t = ty + (size_t)(1 + i) * H;
            for (int s = 0; s < H; s++) t[s] = cadd(t[s], cmul(tavec[i * mH + j * H + s], dj[s]));
        }
        /* v : :529-535 */
        memset(tv, 0, sizeof(cplx) * H);
        for (int i = 0; i < lu; i++)
            for (int s = 0; s < H; s++) tv[s] = csub(tv[s], cmul(tbvec[i * H + s], crs[i * H + s]));
        for (int i = 0; i < idx; i++) for (int j = 0; j < lu; j++) {
            const cplx *pb = (const cplx *)c->keys.pubb[i] + (size_t)j * H;
            for (int s = 0; s < H; s++) tv[s] = cadd(tv[s], cmul(tavec[i * mH + j * H + s], pb[s]));
        }
        ifftto_64(v, tv, &c->F);
        decomp_fft_64(tvvec, v, lu, lbu, &c->F, scr);
        /* w : :547-550 */
        cplx *tyi = ty + (size_t)(1 + idx) * H;
        for (int i = 0; i < lu; i++) {
            const cplx *fb = rlk + (size_t)(i * 3 + 1) * H, *fa = rlk + (size_t)(i * 3 + 2) * H;
            for (int s = 0; s < H; s++) ty[s] = cadd(ty[s], cmul(tvvec[i * H + s], fb[s]));
            for (int s = 0; s < H; s++) tyi[s] = cadd(tyi[s], cmul(tvvec[i * H + s], fa[s]));
        }
        /* :553-556 */
        for (size_t s = 0; s < (size_t)H * (k + 1); s++) tx[s] = cadd(tx[s], ty[s]);
        for (int q = 0; q <= k; q++) ifftto_64(acc + (size_t)q * N, tx + (size_t)q * H, &c->F);
    }
    free(tbvec); free(tavec); free(tv); free(tvvec); free(tx); free(ty); free(y); free(v); free(scr);
}

void orc_phase2(const orc_ctx *c, const double *levkeys, uint32_t btilde, void *acc_out) {
    testvector_64(acc_out, c->p.N, c->p.k, btilde);
    phase2_64(c, (const cplx *)levkeys, acc_out);
}

void orc_blindrotate(const orc_ctx *c, const uint32_t *lwe, void *acc_out) {
    const int N = c->p.N, k = c->p.k, n = c->p.n;
    uint32_t *tilde = malloc(sizeof(uint32_t) * (size_t)(1 + n * k));
    orc_modswitch(c, lwe, tilde);
    if (is_kms(c)) {
        const size_t rowsz = (size_t)c->p.l_lev * 2 * c->F.H * 2;
        double *lev = calloc(rowsz * k, sizeof(double));
        /* bootstrapping.jl:376-378: one task per party in the reference; sequential here (same result) */
        for (int i = 0; i < k; i++) orc_phase1(c, i, tilde + 1 + (size_t)i * n, lev + rowsz * i);
        orc_phase2(c, lev, tilde[0], acc_out);
        free(lev);
    } else {
        testvector_32(acc_out, N, k, tilde[0]);
        if (c->p.scheme == ORC_CCS) blindrotate_ccs(c, tilde + 1, acc_out);
        else blindrotate_sk(c, tilde + 1, acc_out);
    }
    free(tilde);
}

/* ------------------------------------------------------------------ key switching */

static inline const uint32_t *ksk_row(const orc_ctx *c, int party, int coef, uint32_t digit, int level) {
    return c->keys.ksk[party] + ((((size_t)coef * c->Dk + (digit - 1)) * c->p.f + level) * (size_t)(c->p.n + 1));
}
static inline void lwe_addsub(uint32_t *b, uint32_t *a, const uint32_t *row, int n, int sub) {
    if (!sub) { *b += row[0]; for (int i = 0; i < n; i++) a[i] += row[1 + i]; }
    else      { *b -= row[0]; for (int i = 0; i < n; i++) a[i] -= row[1 + i]; }
}

/* CGGI :81-109, LMSS :170-229 (k = 1), CCS :333-364, KMS :564-594, KMS_BLOCK :664-695 */
void orc_keyswitch(const orc_ctx *c, const void *accv, uint32_t *out) {
    const int N = c->p.N, k = c->p.k, n = c->p.n, f = c->p.f, logD = c->p.logD;
    const int block = is_block(c);
    memset(out, 0, sizeof(uint32_t) * (size_t)(1 + n * k));
    uint32_t dig[ORC_MAXL];
    const uint32_t *acc32 = accv; const uint64_t *acc64 = accv;
#define ACC(q, cc) (c->bits == 64 ? (uint32_t)(acc64[(size_t)(q) * N + (cc)] >> 32) : acc32[(size_t)(q) * N + (cc)])
    out[0] = ACC(0, 0);
    for (int i = 0; i < k; i++) {
        uint32_t pb = 0, *pa = out + 1 + (size_t)i * n;   /* partctxt[i]; out.a block starts at zero */
        for (int cc = 0; cc < N; cc++) {
            /* sample extraction of coefficient 0: a'_1 = a[0], a'_j = -a[N-j+1] */
            const uint32_t val = cc == 0 ? ACC(1 + i, 0) : 0u - ACC(1 + i, N - cc);
            if (block && cc < n) { pa[cc] += val; continue; }   /* :678-681 / :177-190 copy */
            ks_digits_32(dig, val, f, logD, block);
            for (int lv = 0; lv < f; lv++) {
                const int32_t sd = (int32_t)dig[lv];
                if (sd > 0) lwe_addsub(&pb, pa, ksk_row(c, i, cc, (uint32_t)sd, lv), n, 0);
                else if (sd < 0) lwe_addsub(&pb, pa, ksk_row(c, i, cc, (uint32_t)(-sd), lv), n, 1);
            }
        }
        out[0] += pb;
    }
#undef ACC
}

void orc_bootstrap(const orc_ctx *c, uint32_t *lwe) {
    const size_t accsz = (size_t)(c->p.k + 1) * c->p.N * (c->bits / 8);
    void *acc = malloc(accsz);
    orc_blindrotate(c, lwe, acc);
    orc_keyswitch(c, acc, lwe);
    free(acc);
}

int orc_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

void orc_gate_batch(const orc_ctx *c, int op, const uint32_t *in1, const uint32_t *in2, uint32_t *out,
                    size_t batch, int nthreads) {
    const size_t len = 1 + (size_t)c->p.n * c->p.k;
#ifdef _OPENMP
    if (nthreads <= 0) nthreads = omp_get_max_threads();
#pragma omp parallel for schedule(dynamic, 1) num_threads(nthreads)
#endif
    for (long g = 0; g < (long)batch; g++) {
        if (op >= 0) orc_gate_linear(c, op, in1 + g * len, in2 + g * len, out + g * len);
        else if (out != in1) memcpy(out + g * len, in1 + g * len, len * sizeof(uint32_t));
        orc_bootstrap(c, out + g * len);
    }
}
