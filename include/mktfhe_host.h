/* mktfhe_host.h -- host-side (CPU, no CUDA) key generation, encryption and decryption.
 *
 * In a deployment with the reference, these steps stay in Julia: `setup`, `party_keygen`,
 * `lwe_encrypt`, `lwe_ith_encrypt`, `lwe_decrypt`, `CRS`
 * (/root/reference/src/tfhe/scheme.jl:151-410, src/tfhe/keygen.jl) -- they are the drop-in surface,
 * not the hot path.  Julia is not available in this image, so the Python mirror of that API
 * (mktfhe_b200/scheme.py) calls this library instead.  It builds key material the way the
 * reference does (same distributions, same structure, same Float64 transform for the uploaded
 * FFT form) from a SEEDED ChaCha20 stream, and emits the flat layouts mktfhe_b200.h consumes.
 *
 * Difference from the reference, stated once: keygen ring products a*s are computed exactly
 * mod 2^w by integer add/sub (keys are binary / ternary) instead of with the Float64x2 FFT
 * (scheme.jl:155,194,233,281,332).  The reference's Float64x2 path deviates from the exact product by
 * up to about 2^11 on Torus64; that rounding artefact is not reproduced.
 */
#ifndef MKTFHE_HOST_H
#define MKTFHE_HOST_H
#include "mktfhe_params.h"

#ifdef __cplusplus
extern "C" {
#endif

/* Randomness.  Every generator below exists in two forms:
 *   *_key(.., const uint8_t key32[32], ..)  production: a 256-bit ChaCha20 key from the OS CSPRNG.  Each party draws its OWN
 *        key for party_keygen (never related to the CRS key, which is public), and every encryption call gets a fresh key --
 *        like the reference, where every sampler call opens an unseeded ChaCha20Stream() (sampler.jl:2-34, lwe.jl:13).
 *   (.., uint64_t seed, ..)                 reproducible: the seed expands to a key through splitmix64.  For tests, benchmarks
 *        and golden vectors only: 64 bits are brute-forceable, and reusing a seed for two encryptions repeats mask and noise.
 */
/* CRS(params): scheme.jl:409-410.  crs_coeff: [l_uni][N] torus (uint32 CCS / uint64 KMS*);
 * crs_fft: [l_uni][N/2] complex = fft.(a, ffter) as in setup (scheme.jl:251,298,349). */
int mktfhe_host_crs(const mktfhe_params *p, uint64_t seed, void *crs_coeff, double *crs_fft);
int mktfhe_host_crs_key(const mktfhe_params *p, const uint8_t *key32, void *crs_coeff, double *crs_fft);

/* party_keygen / setup (scheme.jl:151-166,190-205,227-242,273-287,324-338 -> keygen.jl).
 * Outputs (any may be NULL to skip, except lwekey):
 *   lwekey  [n] uint32 (0/1)
 *   ringkey [N] torus: the RLWE key the key switch undoes (CGGI/LMSS/CCS ringkey, KMS* unikey)
 *   brk, rlk, pubb, ksk: flat layouts of mktfhe_b200.h.
 * crs_coeff is required for CCS / KMS*, ignored otherwise. */
int mktfhe_host_party_keygen(const mktfhe_params *p, uint64_t seed, int party, const void *crs_coeff,
                             uint32_t *lwekey, void *ringkey, double *brk, double *rlk, double *pubb,
                             uint32_t *ksk, int nthreads);
int mktfhe_host_party_keygen_key(const mktfhe_params *p, const uint8_t *key32, int party, const void *crs_coeff,
                                 uint32_t *lwekey, void *ringkey, double *brk, double *rlk, double *pubb,
                                 uint32_t *ksk, int nthreads);

/* lwe_encrypt (scheme.jl:352-368): single-key schemes.  out: [1 + n]. */
int mktfhe_host_lwe_encrypt(const mktfhe_params *p, uint64_t seed, int m, const uint32_t *lwekey, uint32_t *out);
/* lwe_ith_encrypt (scheme.jl:370-386): MK schemes, support on party i (0-based) only.  out: [1 + n*k]. */
int mktfhe_host_lwe_ith_encrypt(const mktfhe_params *p, uint64_t seed, int m, int i, const uint32_t *lwekey_i, uint32_t *out);
/* Bench/test helper with no reference counterpart: a fresh encryption supported on ALL k blocks
 * (the shape a bootstrapped MK ciphertext has).  lwekeys: [k][n]. */
int mktfhe_host_lwe_encrypt_full(const mktfhe_params *p, uint64_t seed, int m, const uint32_t *lwekeys, uint32_t *out);
int mktfhe_host_lwe_encrypt_key(const mktfhe_params *p, const uint8_t *key32, int m, const uint32_t *lwekey, uint32_t *out);
int mktfhe_host_lwe_ith_encrypt_key(const mktfhe_params *p, const uint8_t *key32, int m, int i, const uint32_t *lwekey_i, uint32_t *out);
int mktfhe_host_lwe_encrypt_full_key(const mktfhe_params *p, const uint8_t *key32, int m, const uint32_t *lwekeys, uint32_t *out);
/* Batched forms, OpenMP over ciphertexts (the caller side of the batched hot path).
 * kind: 0 = lwe_encrypt (lwekeys = [n]), 1 = lwe_ith_encrypt for `party` (lwekeys = that party's [n]), 2 = full support
 * (lwekeys = [k][n]).  bits: count bytes 0/1; out: [count][1 + n*k].  Seeded form: ciphertext g equals the single call with
 * seed0 + g.  Keyed form: one key for the batch, ciphertext g draws from its own ChaCha stream (nonce g + 1). */
int mktfhe_host_encrypt_batch(const mktfhe_params *p, uint64_t seed0, int kind, int party, const uint8_t *bits, size_t count,
                              const uint32_t *lwekeys, uint32_t *out, int nthreads);
int mktfhe_host_encrypt_batch_key(const mktfhe_params *p, const uint8_t *key32, int kind, int party, const uint8_t *bits, size_t count,
                                  const uint32_t *lwekeys, uint32_t *out, int nthreads);
/* lwe_decrypt / phase over `count` ciphertexts: bits_out [count] bytes, phases_out [count]. */
int mktfhe_host_decrypt_batch(const mktfhe_params *p, const uint32_t *lwekeys, const uint32_t *cts, size_t count, uint8_t *bits_out, int nthreads);
int mktfhe_host_phase_batch(const mktfhe_params *p, const uint32_t *lwekeys, const uint32_t *cts, size_t count, uint32_t *phases_out, int nthreads);
/* b + <a, s> over all blocks (lwe.jl:31-33); lwekeys: [k][n] (k = 1 for single key). */
uint32_t mktfhe_host_lwe_phase(const mktfhe_params *p, const uint32_t *lwekeys, const uint32_t *ct);
/* lwe_decrypt (scheme.jl:388-407): returns 0 / 1. */
int mktfhe_host_lwe_decrypt(const mktfhe_params *p, const uint32_t *lwekeys, const uint32_t *ct);

/* FFTransformer tables (fft.jl:26-44), each N/2 complex. */
void mktfhe_host_fft_tables(int N, double *psi, double *psiinv, double *roots, double *rootsinv);

#ifdef __cplusplus
}
#endif
#endif
