/* mktfhe_params.h -- parameter struct shared by the GPU library (mktfhe_b200.h) and the host
 * key-generation library (mktfhe_host.h).
 *
 * One flat struct replaces the five reference parameter structs
 *   TFHEparams_bin   /root/reference/src/tfhe/scheme.jl:6-19
 *   TFHEparams_block /root/reference/src/tfhe/scheme.jl:22-36
 *   CCSparams        /root/reference/src/tfhe/scheme.jl:40-54
 *   KMSparams        /root/reference/src/tfhe/scheme.jl:57-77
 *   KMSparams_block  /root/reference/src/tfhe/scheme.jl:80-101
 * with the same field names; `scheme` selects which of them it stands for.  Named presets
 * (src/tfhe/params.jl) live in mktfhe_b200/params.py and julia/MKTFHEB200.jl.
 */
#ifndef MKTFHE_PARAMS_H
#define MKTFHE_PARAMS_H
#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

enum mktfhe_scheme { MKTFHE_CGGI = 0, MKTFHE_LMSS = 1, MKTFHE_CCS = 2, MKTFHE_KMS = 3, MKTFHE_KMS_BLOCK = 4 };
/* gate opcodes: /root/reference/src/tfhe/gate.jl:1-52 */
enum mktfhe_gate { MKTFHE_NAND = 0, MKTFHE_AND = 1, MKTFHE_OR = 2, MKTFHE_XOR = 3, MKTFHE_XNOR = 4, MKTFHE_NOR = 5,
                   MKTFHE_NOT = 6 /* circuit levels only: negation, no bootstrap (gate.jl:55-58) */ };

typedef struct mktfhe_params {
    int32_t scheme;            /* enum mktfhe_scheme */
    int32_t n;                 /* LWE dimension (= d*ell for block schemes) */
    int32_t d, ell;            /* block-binary key shape; 0,0 when not a block scheme */
    int32_t f, logD;           /* key-switching gadget length / log2 base */
    int32_t N;                 /* RLWE ring dimension */
    int32_t k;                 /* number of parties (CCS, KMS*) or RLWE length (CGGI, LMSS: must be 1) */
    int32_t l_gsw, logB_gsw;   /* RGSW gadget (CGGI, LMSS, KMS*) */
    int32_t l_lev, logB_lev;   /* LEV gadget (KMS*) */
    int32_t l_uni, logB_uni;   /* UniEnc gadget (CCS, KMS*) */
    double alpha;              /* LWE noise std-dev, Torus32 units */
    double beta;               /* RLWE noise std-dev, RLWE torus units */
} mktfhe_params;

/* RLWE torus width in bits: 64 for KMS / KMS_BLOCK (params.jl:47-125), 32 otherwise. */
static inline int mktfhe_torus_bits(const mktfhe_params *p) {
    return (p->scheme == MKTFHE_KMS || p->scheme == MKTFHE_KMS_BLOCK) ? 64 : 32;
}
/* ksk rows per ring coefficient: D-1 (unbalanced digits) or D/2 (balanced, block schemes). */
static inline int mktfhe_ksk_rows(const mktfhe_params *p) {
    const int D = 1 << p->logD;
    return (p->scheme == MKTFHE_LMSS || p->scheme == MKTFHE_KMS_BLOCK) ? D / 2 : D - 1;
}
/* Flat key sizes per party (layouts: see mktfhe_b200.h). */
static inline size_t mktfhe_brk_doubles(const mktfhe_params *p) {
    const size_t polys = p->scheme == MKTFHE_CCS ? (size_t)3 * p->l_uni : (size_t)4 * p->l_gsw;
    return (size_t)p->n * polys * (size_t)p->N;            /* N/2 complex = N doubles per poly */
}
static inline size_t mktfhe_rlk_doubles(const mktfhe_params *p) { return (size_t)3 * p->l_uni * p->N; }
static inline size_t mktfhe_pubb_doubles(const mktfhe_params *p) { return (size_t)p->l_uni * p->N; }
static inline size_t mktfhe_crs_doubles(const mktfhe_params *p) { return (size_t)p->l_uni * p->N; }
static inline size_t mktfhe_ksk_words(const mktfhe_params *p) {
    return (size_t)p->N * mktfhe_ksk_rows(p) * p->f * (size_t)(p->n + 1);
}
static inline size_t mktfhe_lwe_words(const mktfhe_params *p) { return (size_t)1 + (size_t)p->n * p->k; }

#ifdef __cplusplus
}
#endif
#endif
