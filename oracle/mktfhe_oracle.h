/* mktfhe_oracle.h -- CPU oracle for the MKTFHE gate-bootstrapping hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load this library.  The product
 * (mktfhe_b200/, include/) never links, imports or calls it.
 *
 * It is a plain-C restatement of the reference's Float64 hot path
 * (/root/reference/src/{ring,ciphertext,tfhe}), function by function, with the same
 * butterfly schedule, the same operation order and no FMA contraction.
 *
 * PARITY PINNING: the reference holds no golden vectors or known-answer tests and Julia is
 * not installed, so ciphertext-level parity with a real reference run is UNPINNED.  What is
 * pinned (tests/test_oracle_*.py): (1) the transform against exact integer negacyclic
 * convolution and against the closed form slot j = p(exp(-i*pi*(4*brv(j)+1)/N)); (2) gadget
 * recomposition bounds; (3) the reference's own acceptance criterion (test/CGGI.jl:34,
 * test/LMSS.jl:34, test/CCS.jl:37, test/KMS.jl:37, test/KMSblock.jl:37): random chains of all
 * six gates decrypt to the plaintext circuit, at the five parameter sets those scripts use.
 *
 * Flat key layouts (the same arrays the product's C-ABI takes; SURVEY App. E order):
 *   brk  RGSW schemes (CGGI, LMSS, KMS, KMS_BLOCK), per party:
 *            [idx < n][basket: 0 = basketb, 1 = basketa[1]][j < l_gsw][comp: 0 = b, 1 = a[1]][H] complex
 *   brk  CCS, per party:  [idx < n][j < l_uni][0 = d[j], 1 = f.stack[j].b, 2 = f.stack[j].a[1]][H] complex
 *   rlk  KMS*, per party: [j < l_uni][0 = d[j], 1 = f.stack[j].b, 2 = f.stack[j].a[1]][H] complex
 *   pubb KMS*, CCS, per party: [j < l_uni][H] complex
 *   crs  KMS*, CCS:       [j < l_uni][H] complex  (scheme.a, FFT form)
 *   ksk  per party:       [c < N][digit-1 < Dk][level < f][1 + n] uint32 (b, then a);
 *                          Dk = D-1 (CGGI, CCS, KMS) or D/2 (LMSS, KMS_BLOCK; rows c < n unused)
 *   complex = interleaved (re, im) doubles, reference slot order (fft.jl bit-reversed evaluation order).
 *   LWE ciphertext: uint32 [1 + n*k]: b, then a (party-major blocks of n).
 *   RLWE accumulator: torus [(k+1)][N]: b, a_1 .. a_k.  torus = uint64 for KMS*, uint32 otherwise.
 */
#ifndef MKTFHE_ORACLE_H
#define MKTFHE_ORACLE_H
#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

enum { ORC_CGGI = 0, ORC_LMSS = 1, ORC_CCS = 2, ORC_KMS = 3, ORC_KMS_BLOCK = 4 };
enum { ORC_NAND = 0, ORC_AND = 1, ORC_OR = 2, ORC_XOR = 3, ORC_XNOR = 4, ORC_NOR = 5 };

/* Mirrors the reference parameter structs (src/tfhe/scheme.jl:6-101). */
typedef struct {
    int32_t scheme;
    int32_t n;                 /* LWE dimension (= d*ell for block schemes) */
    int32_t d, ell;            /* block-binary key shape (0 when not block) */
    int32_t f, logD;           /* key-switching gadget */
    int32_t N;                 /* ring dimension */
    int32_t k;                 /* parties (CCS/KMS) or RLWE length (CGGI/LMSS; must be 1) */
    int32_t l_gsw, logB_gsw;
    int32_t l_lev, logB_lev;
    int32_t l_uni, logB_uni;
    double alpha, beta;
} orc_params;

typedef struct {
    const double *const *brk;    /* [k] */
    const double *const *rlk;    /* [k] or NULL */
    const double *const *pubb;   /* [k] or NULL */
    const uint32_t *const *ksk;  /* [k] */
    const double *crs;           /* or NULL */
} orc_keys;

typedef struct orc_ctx orc_ctx;

/* Borrowing constructor: key arrays must outlive the context. */
orc_ctx *orc_create(const orc_params *p, const orc_keys *keys);
void orc_destroy(orc_ctx *c);

/* Transform tables (fft.jl:26-44): each H complex, interleaved. */
void orc_fft_tables(int N, double *psi, double *psiinv, double *roots, double *rootsinv);
/* scheme.jl:121-146: out = 2N polys x H complex; entry a-1 = FFT(X^a - 1), entry 2N-1 = 0. */
void orc_monomials(int N, double *out);

/* Unit operations (bits = 32 / 64 selects the torus). */
void orc_fft(int N, int bits, const void *poly, double *out);
void orc_ifft(int N, int bits, double *in_consumed, void *poly);
void orc_decomp(int N, int bits, int l, int logB, const void *poly, void *digits /* [l][N] */);

/* bootstrapping.jl:8-9: tilde[0] = b~, tilde[1..n*k] = a~. */
void orc_modswitch(const orc_ctx *c, const uint32_t *lwe, uint32_t *tilde);
/* Gate linear part only (gate.jl), no bootstrap. */
void orc_gate_linear(const orc_ctx *c, int op, const uint32_t *in1, const uint32_t *in2, uint32_t *out);
/* One iteration of the phase-1 / CGGI loop body on one RLWE row (b, a):
 * acc += ifft(monomial[atilde] * (acc [.] brk[party][idx])).   bootstrapping.jl:47-74, 413-438 */
void orc_cmux_step(const orc_ctx *c, int party, int idx, uint32_t atilde, void *acc_row /* [2][N] torus */);
/* One block iteration of the LMSS / KMS_BLOCK loop on one RLWE row (bootstrapping.jl:124-163, 624-655):
 * at = the ell rotations of block `blk`. */
void orc_block_step(const orc_ctx *c, int party, int blk, const uint32_t *at, void *acc_row);
/* KMS / KMS_BLOCK phase 1 for one party (bootstrapping.jl:389-443, 599-659).
 * out: [rows][2][H] complex, rows = 1 for party 0 else l_lev. */
void orc_phase1(const orc_ctx *c, int party, const uint32_t *tildea_party, double *levkey_out);
/* Test vector + blind rotation: lwe (after the gate's linear part) -> accumulator. */
void orc_blindrotate(const orc_ctx *c, const uint32_t *lwe, void *acc_out);
/* KMS phase 2 alone: levkeys [k][l_lev][2][H] (party 0 uses row 0 only), btilde -> acc. */
void orc_phase2(const orc_ctx *c, const double *levkeys, uint32_t btilde, void *acc_out);
void orc_keyswitch(const orc_ctx *c, const void *acc, uint32_t *lwe_out);
/* bootstrapping! (bootstrapping.jl:4-27), in place. */
void orc_bootstrap(const orc_ctx *c, uint32_t *lwe);
/* Gate + bootstrap over a batch, OpenMP over gates (nthreads <= 0: all cores). */
void orc_gate_batch(const orc_ctx *c, int op, const uint32_t *in1, const uint32_t *in2, uint32_t *out,
                    size_t batch, int nthreads);
int orc_max_threads(void);

#ifdef __cplusplus
}
#endif
#endif
