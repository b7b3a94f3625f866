/* mktfhe_b200.h -- C-ABI of the B200 (sm_100a) gate-bootstrapping library libmktfhe_b200.so.
 *
 * The reference (SNUCP/MKTFHE, Julia) has no FFI; this header defines the boundary a Julia `ccall`
 * binding uses (julia/MKTFHEB200.jl).  The cut is the reference's
 *   bootstrapping!(ctxt, scheme)        /root/reference/src/tfhe/bootstrapping.jl:4-27
 *   NAND/AND/OR/XOR/XNOR/NOR            /root/reference/src/tfhe/gate.jl:1-52
 * plus a one-time key upload hooked after `setup` (/root/reference/src/tfhe/scheme.jl:151,190,244,292,343).
 * Key generation, encryption and decryption stay on the host (the reference's own code, or
 * mktfhe_host.h when Julia is absent).
 *
 * Conventions
 *   - Plain pointers and sizes only.  Unless a name ends in _dev, pointers are HOST pointers, borrowed
 *     for the duration of the call; the library copies what it keeps and owns all device memory.
 *   - Every function returns 0 on success or a negative mktfhe_status; mktfhe_last_error() gives text.
 *     No exception crosses the boundary and the library never calls back into the host language.
 *   - One context belongs to one device and is used from one host thread at a time.  Multi-GPU = one
 *     context per device (one process per GPU under torchrun), or ONE front context over several devices
 *     (mktfhe_ctx_create_multi) that shards every batch itself; gates are independent, no collective is involved.
 *   - There is no CPU fallback: without a CUDA device every compute entry point fails with
 *     MKTFHE_ERR_CUDA.
 *
 * Flat layouts (SURVEY App. E traversal order; "complex" = interleaved (re, im) doubles in the
 * reference's slot order, i.e. exactly the Vector{ComplexF64} of a TransNativePoly; H = N/2):
 *   brk  RGSW schemes (CGGI, LMSS, KMS, KMS_BLOCK), per party
 *          [idx < n][basket: 0 = basketb, 1 = basketa[1]][j < l_gsw][comp: 0 = .b, 1 = .a[1]][H] complex
 *          = scheme.btk[p].brk[idx]   (keygen.jl:12-14,39-41,106-108,143-145; read at bootstrapping.jl:63-68,427-432)
 *   brk  CCS, per party   [idx < n][j < l_uni][0 = d[j], 1 = f.stack[j].b, 2 = f.stack[j].a[1]][H] complex
 *          = scheme.btk[p].brk[idx]::TransUniEnc   (keygen.jl:71-73; read at bootstrapping.jl:280-319)
 *   rlk  KMS*, per party  [j < l_uni][0 = d[j], 1 = f.stack[j].b, 2 = f.stack[j].a[1]][H] complex   (keygen.jl:103,139)
 *   pubb CCS, KMS*, per party  [j < l_uni][H] complex = btk[p].b   (keygen.jl:68,100,136)
 *   crs  CCS, KMS*        [j < l_uni][H] complex = scheme.a        (scheme.jl:251,298,349)
 *   ksk  per party        [c < N][digit-1 < Dk][level < f][1 + n] uint32: .b then .a of
 *          btk[p].ksk[digit, c+1].stack[level+1]; Dk = D-1 (CGGI, CCS, KMS) or D/2 (LMSS, KMS_BLOCK, where
 *          rows c < n are never read)   (keygen.jl:16-24,43-52,75-79,110-114,147-151)
 *   LWE ciphertext   uint32 [1 + n*k]: b, then a in party-major blocks of n   (lwe.jl:1-9, scheme.jl:379-386)
 *   RLWE accumulator torus [(k+1)][N]: b, a_1..a_k; torus = uint64 for KMS*, uint32 otherwise
 *   levkeys (KMS* phase-1 output)  [R][2][H] complex per gate, R = 1 + (k-1)*l_lev rows:
 *          party 0 row 0, then party p >= 1 rows 1 + (p-1)*l_lev + r   (bootstrapping.jl:400-409,441-442)
 */
#ifndef MKTFHE_B200_H
#define MKTFHE_B200_H
#include "mktfhe_params.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct mktfhe_ctx mktfhe_ctx;

enum mktfhe_status {
    MKTFHE_OK = 0,
    MKTFHE_ERR_PARAMS = -1,      /* unsupported or inconsistent parameters */
    MKTFHE_ERR_CUDA = -2,        /* CUDA runtime error, or no device */
    MKTFHE_ERR_STATE = -3,       /* keys missing / not finalized */
    MKTFHE_ERR_ARG = -4          /* null pointer, bad party index, bad opcode */
};

/* Arithmetic mode of the floating-point stages.
 *   STRICT: the reference's butterfly schedule, slot order and operation order with no FMA contraction;
 *           bit-identical to the CPU oracle on every ciphertext coefficient (tests/test_gpu_strict.py).
 *   FAST:   the production path: register-resident FFT with fused multiply-add and its own slot order.
 *           Integer stages stay bit-exact; accumulator coefficients differ from STRICT only by FFT
 *           rounding (per-step tolerance in tests/test_gpu_fast.py).  Default. */
enum mktfhe_mode { MKTFHE_MODE_STRICT = 0, MKTFHE_MODE_FAST = 1 };

/* ---- lifetime -------------------------------------------------------------------------------- */
/* Replaces: scheme construction in `setup` (scheme.jl:151-166,190-205,244-252,292-299,343-350). */
int mktfhe_ctx_create(const mktfhe_params *params, int device, mktfhe_ctx **out);
/* Multi-GPU in ONE process behind the same entry points: a front context over `ndev` devices
 * (devices == NULL: 0 .. ndev-1; ndev <= 0: every visible device).  mktfhe_upload_party_key / mktfhe_upload_common copy the
 * keys host -> first device once; mktfhe_finalize_keys replicates them device to device (cudaMemcpyPeer over NVLink) and
 * builds tables and FAST layouts on every device; mktfhe_gate_batch / mktfhe_bootstrap_batch cut the host batch into
 * contiguous slices, one per device, each run from its own host thread on its own stream (results are bit-identical to the
 * single-device call: gates are independent).  Parity hooks, the wire table and the _dev entry point stay single-device.
 * The only parallelism the reference has is Threads.@threads over parties (bootstrapping.jl:376-378,573). */
int mktfhe_ctx_create_multi(const mktfhe_params *params, int ndev, const int *devices, mktfhe_ctx **out);
/* Number of devices behind ctx (1 for a plain context); fills devices_out[0 .. min(n, cap)). */
int mktfhe_ctx_devices(const mktfhe_ctx *ctx, int *devices_out, int cap);
void mktfhe_ctx_destroy(mktfhe_ctx *ctx);
const char *mktfhe_last_error(const mktfhe_ctx *ctx);   /* ctx may be NULL: last creation error */
int mktfhe_set_mode(mktfhe_ctx *ctx, int mode);
int mktfhe_get_mode(const mktfhe_ctx *ctx);

/* ---- one-time key upload --------------------------------------------------------------------- */
/* Replaces: holding `btk[party]` in the scheme struct (scheme.jl:107-116,209-219,256-265,301-312).
 * rlk / pubb may be NULL for schemes that have none. party = 0 for CGGI / LMSS.
 * Sources may be host pointers or device pointers of any GPU (copied with cudaMemcpyDefault), so a key set
 * that arrived by NCCL broadcast is uploaded without a host round trip. */
int mktfhe_upload_party_key(mktfhe_ctx *ctx, int party, const double *brk, const double *rlk,
                            const double *pubb, const uint32_t *ksk);
/* Replaces: `fft(a, ffter)` stored as scheme.a (scheme.jl:251,298,349). NULL for CGGI / LMSS. */
int mktfhe_upload_common(mktfhe_ctx *ctx, const double *crs_fft);
/* ---- key generation on the device ----------------------------------------------------------------
 * Replaces: CRS (scheme.jl:409-410) + fft.(a, ffter) and party_keygen's BootKey constructors (keygen.jl:3-155: rgsw_encrypt per
 * key bit, unienc_encrypt, gen_b, lev_encrypt per ring coefficient) for the case that one process owns the seeds (benchmarks,
 * tests, a party generating its own keys on its own GPU).  The material is generated from the SAME seeded ChaCha20 streams as
 * mktfhe_host_party_keygen (include/mktfhe_host.h), with exact integer ring products and the reference's Float64 transform, so
 * for one seed it is byte-identical to the host library's (tests/test_gpu_keygen.py compares SHA-256) -- and it never crosses
 * PCIe.  key32 = 32-byte ChaCha20 key (production), or NULL to expand the 64-bit `seed` (reproducible runs).
 * mktfhe_keygen_common must precede mktfhe_keygen_party for CCS / KMS* (it generates and keeps the CRS); then mktfhe_finalize_keys.
 * The matching secret keys come from mktfhe_host_party_keygen with all evaluation-key outputs NULL (same seed). */
int mktfhe_keygen_common(mktfhe_ctx *ctx, uint64_t seed, const uint8_t *key32);
int mktfhe_keygen_party(mktfhe_ctx *ctx, int party, uint64_t seed, const uint8_t *key32);
/* Parity hook: copies a party's uploaded or generated keys back in the flat layouts (any pointer may be NULL). */
int mktfhe_download_party_key(mktfhe_ctx *ctx, int party, double *brk, double *rlk, double *pubb, uint32_t *ksk, double *crs_fft);
/* Builds the transform tables (fft.jl:26-44) and monomial table (scheme.jl:121-146) on the device and
 * the FAST-mode key layouts.  Must be called once after all uploads. */
int mktfhe_finalize_keys(mktfhe_ctx *ctx);

/* ---- the hot path ---------------------------------------------------------------------------- */
/* Replaces: NAND ... NOR (gate.jl:1-52) over a batch: out[g] = bootstrap(linear(in1[g], in2[g])).
 * batch = 1 is the reference's per-gate call.  Host buffers; H2D and D2H copies are inside the call. */
int mktfhe_gate_batch(mktfhe_ctx *ctx, int gate_op, const uint32_t *in1, const uint32_t *in2,
                      uint32_t *out, size_t batch);
/* Batches whose scratch (accumulators, RLEV rows: 0.2 MB per gate at KMS2party, 4.6 MB at KMS32party) exceeds the workspace
 * budget run in chunks inside the call; the budget is min(24 GiB, 40 % of device memory) or $MKTFHE_WORKSPACE_MB. */
/* Replaces: bootstrapping!(ctxt, scheme) (bootstrapping.jl:4-27) over a batch (out may alias in). */
int mktfhe_bootstrap_batch(mktfhe_ctx *ctx, const uint32_t *in, uint32_t *out, size_t batch);
/* Same with ciphertexts already resident in device memory of ctx's device (no copies, asynchronous on
 * the context stream; call mktfhe_sync before reading).  gate_op < 0 means bootstrap of in1 only. */
int mktfhe_gate_batch_dev(mktfhe_ctx *ctx, int gate_op, const uint32_t *in1_dev, const uint32_t *in2_dev,
                          uint32_t *out_dev, size_t batch);
int mktfhe_sync(mktfhe_ctx *ctx);
/* The CUDA stream (cudaStream_t) all work of this context is launched on, for event timing. */
void *mktfhe_stream(mktfhe_ctx *ctx);

/* ---- gate circuits (the caller side of the path) ----------------------------- */
/* The reference evaluates a circuit one gate call at a time (test/KMS.jl:28-36 chains NAND ... NOR calls); here the
 * circuit's wires live in a device-resident table of LWE records owned by the context, and one call evaluates one
 * LEVEL of independent gates:  for g < batch:  wires[dst[g]] = bootstrap(op[g](wires[src1[g]], wires[src2[g]])).
 *   ops[g]: MKTFHE_NAND .. MKTFHE_NOR (gate.jl:1-52), -1 = bootstrapping! of src1 alone, MKTFHE_NOT = NOT! of src1
 *   (negation only, no bootstrap, gate.jl:55-58).  ops / src1 / src2 / dst are HOST arrays of `batch` entries.
 * A dst wire must not also be a source of the same call (levels of a circuit in SSA form satisfy this).
 * Synchronous: the level has completed when the call returns. */
int mktfhe_wires_resize(mktfhe_ctx *ctx, size_t nwires);     /* (re)allocates the table, zero-filled; 0 frees it */
/* Copies `count` LWE records ([count][1 + n*k] uint32; host or device pointer) into / out of wires first .. first+count-1. */
int mktfhe_wires_write(mktfhe_ctx *ctx, size_t first, size_t count, const uint32_t *cts);
int mktfhe_wires_read(mktfhe_ctx *ctx, size_t first, size_t count, uint32_t *cts);
int mktfhe_gate_level(mktfhe_ctx *ctx, const int32_t *ops, const int32_t *src1, const int32_t *src2,
                      const int32_t *dst, size_t batch);

/* ---- parity / debug hooks (host buffers) ------------------------------------------------------- */
/* gate.jl linear part only. */
int mktfhe_gate_linear_batch(mktfhe_ctx *ctx, int gate_op, const uint32_t *in1, const uint32_t *in2,
                             uint32_t *out, size_t batch);
/* bootstrapping.jl:8-9: tilde[g] = [b~, a~...] (1 + n*k words). */
int mktfhe_modswitch_batch(mktfhe_ctx *ctx, const uint32_t *lwe, uint32_t *tilde, size_t batch);
/* Test vector + blind rotation (bootstrapping.jl:11-25): lwe -> accumulator [(k+1)][N] torus per gate. */
int mktfhe_blindrotate_batch(mktfhe_ctx *ctx, const uint32_t *lwe, void *acc_out, size_t batch);
/* KMS* phase 1 only (bootstrapping.jl:389-443, 599-659): lwe -> levkeys [R][2][H] complex per gate,
 * always in the reference's slot order. */
int mktfhe_phase1_batch(mktfhe_ctx *ctx, const uint32_t *lwe, double *levkeys_out, size_t batch);
/* Key switch only (bootstrapping.jl:81-109,170-229,333-364,564-594,664-695): accumulator -> LWE. */
int mktfhe_keyswitch_batch(mktfhe_ctx *ctx, const void *acc, uint32_t *lwe_out, size_t batch);
/* One loop iteration of phase 1 / CGGI on `batch` independent RLWE rows [2][N] torus, all against
 * brk[party][idx] with rotation atilde[g]:  acc += ifft(monomial[atilde] * (acc [.] brk))
 * (bootstrapping.jl:47-74, 413-438). */
int mktfhe_cmux_step_batch(mktfhe_ctx *ctx, int party, int idx, const uint32_t *atilde, void *acc_rows,
                           size_t batch);
/* One block iteration of the LMSS / KMS_BLOCK loop (bootstrapping.jl:124-163, 624-655) on `batch` independent rows:
 * atilde[g][ell] are the rotations of block `blk`'s key bits. */
int mktfhe_block_step_batch(mktfhe_ctx *ctx, int party, int blk, const uint32_t *atilde, void *acc_rows,
                            size_t batch);
/* One gadget product with the device functions FAST phase 2 is built from (KMS*, N = 2048):
 *   out[g][c] = native(ifft( Sum_{j<l} fft(D_j(polys[g])) (.) keys[j][c] )),  c < ncomp <= 3, keys [l][ncomp][H] complex in the
 *   reference slot order (the shape of a LEV product with levkey rows, bootstrapping.jl:483-499, of the u / v sums, :520-535, and
 *   of w, :538-550).  Lets a test feed the oracle's intermediate polynomials into single products of phase 2. */
int mktfhe_gadget_product_batch(mktfhe_ctx *ctx, int l, int logB, const uint64_t *polys, const double *keys, int ncomp,
                                uint64_t *out, size_t batch);
/* The same for the Torus32 schemes (N = 1024; polys / out uint32, l * logB <= 32): the shape of the CCS hybrid product's sums
 * (bootstrapping.jl:277-294 u_c / v_c from an accumulator component against d[j] and crs[j] | b[j], :313-320 w from v against f[j]),
 * built from the device functions of fast32::k_ccs_fast. */
int mktfhe_gadget_product32_batch(mktfhe_ctx *ctx, int l, int logB, const uint32_t *polys, const double *keys, int ncomp,
                                  uint32_t *out, size_t batch);
/* fftto! / ifftto! (fft.jl:57-63,74-81) and poly decompto! (gsw.jl:86-96) on `batch` polynomials.
 * bits = 32 / 64 selects the torus; spectra in the reference's slot order; STRICT arithmetic. */
int mktfhe_fft_batch(mktfhe_ctx *ctx, int bits, const void *polys, double *spectra, size_t batch);
int mktfhe_ifft_batch(mktfhe_ctx *ctx, int bits, const double *spectra, void *polys, size_t batch);
int mktfhe_decomp_batch(mktfhe_ctx *ctx, int bits, int l, int logB, const void *polys, void *digits,
                        size_t batch);

/* ---- measurement ----------------------------------------------------------------------------- */
/* Device time (ms, CUDA events on the context stream) spent in each stage of the most recent
 * gate/bootstrap batch call, and the number of kernel launches it made. */
enum mktfhe_stage { MKTFHE_STAGE_PREP = 0, MKTFHE_STAGE_PHASE1 = 1, MKTFHE_STAGE_PHASE2 = 2,
                    MKTFHE_STAGE_KEYSWITCH = 3, MKTFHE_STAGE_COUNT = 4 };
int mktfhe_last_stage_ms(mktfhe_ctx *ctx, float *ms_out /* [MKTFHE_STAGE_COUNT] */, int *launches_out);
/* Measured FP64 FMA peak of the device (TFLOP/s, 2 flops per DFMA) from a register-only DFMA loop:
 * the FP64 roofline denominator MEASURED_PEAKS.json lacks. */
int mktfhe_measure_dfma_peak(mktfhe_ctx *ctx, double *tflops_out);

#ifdef __cplusplus
}
#endif
#endif
