// kernels_strict.cuh -- STRICT-mode blind rotation kernels: the reference's arithmetic, operation for
// operation (no FMA, same accumulation order), so that every accumulator coefficient equals the CPU
// oracle's bit for bit.  One CTA of 256 threads = G groups of H/8 threads; each group transforms one
// polynomial per round in its own padded shared-memory buffer (fft_strict.cuh).
//
//   k_rgsw_blindrotate   CGGI  blindrotate!  /root/reference/src/tfhe/bootstrapping.jl:32-76
//                        LMSS  blindrotate!  :114-165
//                        KMS   phase_1       :389-443      KMS_block phase_1  :599-659
//   k_kms_phase2         phase_2!            :448-558
//   k_ccs_blindrotate    CCS   blindrotate!  :234-328
#pragma once
#include "fft_strict.cuh"

enum { RG_MODE_KMS = 0, RG_MODE_SK = 1, RG_MODE_STEP = 2 };

struct RgswArgs {
    const uint32_t *tilde;        // [B][1 + n*k]: b~, a~   (RG_MODE_STEP: [B][ELL] rotations)
    const cplx *const *brk;       // [k] party key pointers, reference slot order
    const cplx *mono;             // [2N][H] monomial table (scheme.jl:121-146)
    FftTables tb;
    cplx *lev_out;                // RG_MODE_KMS: [B][R][2][H]
    void *acc_io;                 // RG_MODE_SK: out [B][2][N];  RG_MODE_STEP: in/out [B][2][N]
    int mode;
    int n, d, k, l, logB, l_lev, logB_lev, R, lwe_words;
    int step_party, step_idx, step_block;
};

// src/tfhe/bootstrapping.jl:11-22
template <class T> __device__ __forceinline__ T testvector_coef(int i0 /*0-based*/, uint32_t tb, int N) {
    const T e = (T)1 << (sizeof(T) * 8 - 3);
    const uint32_t i = (uint32_t)i0 + 1;
    if (tb <= (uint32_t)N) return i <= tb ? e : (T)((T)0 - e);
    return i <= tb - (uint32_t)N ? (T)((T)0 - e) : e;
}

template <class T, int H, int ELL>
__global__ void __launch_bounds__(MK_THREADS) k_rgsw_blindrotate(const RgswArgs a) {
    constexpr int N = 2 * H, GT = H / 8, G = MK_THREADS / GT, SL = H / MK_THREADS, PL = padded_len(H);
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T *accb = reinterpret_cast<T *>(smem_raw), *acca = accb + N;
    cplx *bufs = reinterpret_cast<cplx *>(smem_raw + 2 * N * sizeof(T));
    const int tid = threadIdx.x, grp = tid / GT, t = tid % GT;
    const int unit = blockIdx.x;

    int gate, party, row = 0;
    if (a.mode == RG_MODE_KMS) {
        gate = unit / a.R;
        const int r = unit % a.R;
        party = r == 0 ? 0 : 1 + (r - 1) / a.l_lev;
        row = r == 0 ? 0 : (r - 1) % a.l_lev;
    } else if (a.mode == RG_MODE_SK) { gate = unit; party = 0; }
    else { gate = unit; party = a.step_party; }

    // ---- accumulator init
    if (a.mode == RG_MODE_KMS) {          // trivial RLEV row of 1: bootstrapping.jl:402-408
        for (int i = tid; i < N; i += MK_THREADS) { accb[i] = 0; acca[i] = 0; }
        __syncthreads();
        if (tid == 0) accb[0] = (T)1 << (sizeof(T) * 8 - (row + 1) * a.logB_lev);
    } else if (a.mode == RG_MODE_SK) {    // test vector: bootstrapping.jl:11-23
        const uint32_t tb = a.tilde[(size_t)gate * a.lwe_words];
        for (int i = tid; i < N; i += MK_THREADS) { accb[i] = testvector_coef<T>(i, tb, N); acca[i] = 0; }
    } else {
        const T *src = reinterpret_cast<const T *>(a.acc_io) + (size_t)gate * 2 * N;
        for (int i = tid; i < N; i += MK_THREADS) { accb[i] = src[i]; acca[i] = src[N + i]; }
    }
    __syncthreads();

    const cplx *brk = a.brk[party];
    const int l = a.l, nd = 2 * a.l;
    const size_t per_idx = (size_t)4 * l * H;
    const uint32_t *at_src = a.mode == RG_MODE_STEP ? a.tilde + (size_t)gate * ELL
                                                    : a.tilde + (size_t)gate * a.lwe_words + 1 + (size_t)party * a.n;
    const int nblk = a.mode == RG_MODE_STEP ? 1 : (ELL == 1 ? a.n : a.d);

    for (int blk = 0; blk < nblk; blk++) {
        uint32_t at[ELL];
        bool any = false;
#pragma unroll
        for (int b = 0; b < ELL; b++) { at[b] = at_src[blk * ELL + b]; any |= at[b] > 0; }
        if (!any) continue;                                   // :48 / :413 ; block: whole-block no-op
        const int idx0 = (a.mode == RG_MODE_STEP ? a.step_idx : blk) * ELL;   // step mode: step_idx counts blocks when ELL > 1

        cplx tacc[ELL][2][SL];
#pragma unroll
        for (int b = 0; b < ELL; b++)
#pragma unroll
            for (int q = 0; q < SL; q++) tacc[b][0][q] = tacc[b][1][q] = make_double2(0.0, 0.0);

        for (int d0 = 0; d0 < nd; d0 += G) {
            const int dg = d0 + grp;
            const bool active = dg < nd;
            const T *src = dg < l ? accb : acca;
            DigitLoad<T, H> ld{src, a.tb.roots, dg < l ? dg : dg - l, l, a.logB};
            fft_forward_strict<H>(bufs + grp * PL, a.tb, t, active, ld);
            for (int g2 = 0; g2 < G; g2++) {
                const int dg2 = d0 + g2;
                if (dg2 >= nd) break;
#pragma unroll
                for (int q = 0; q < SL; q++) {
                    const int s = tid + q * MK_THREADS;
                    const cplx x = bufs[g2 * PL + PAD(s)];
#pragma unroll
                    for (int b = 0; b < ELL; b++) {
                        if (at[b] == 0) continue;
                        const cplx *kp = brk + (size_t)(idx0 + b) * per_idx + (size_t)(dg2 * 2) * H + s;
                        tacc[b][0][q] = cadd_s(tacc[b][0][q], cmul_s(x, __ldg(kp)));
                        tacc[b][1][q] = cadd_s(tacc[b][1][q], cmul_s(x, __ldg(kp + H)));
                    }
                }
            }
            __syncthreads();
        }
        // monomial: ELL == 1: mul!(monomial, tacc) (:71/:435); block: tacc2 += monomial*tacc (:157/:648)
#pragma unroll
        for (int q = 0; q < SL; q++) {
            const int s = tid + q * MK_THREADS;
            cplx rb, ra;
            if (ELL == 1) {
                const cplx m = __ldg(&a.mono[(size_t)(at[0] - 1) * H + s]);
                rb = cmul_s(m, tacc[0][0][q]); ra = cmul_s(m, tacc[0][1][q]);
            } else {
                rb = ra = make_double2(0.0, 0.0);
#pragma unroll
                for (int b = 0; b < ELL; b++) {
                    if (at[b] == 0) continue;
                    const cplx m = __ldg(&a.mono[(size_t)(at[b] - 1) * H + s]);
                    rb = cadd_s(rb, cmul_s(m, tacc[b][0][q]));
                    ra = cadd_s(ra, cmul_s(m, tacc[b][1][q]));
                }
            }
            bufs[PAD(s)] = rb;
            bufs[PL + PAD(s)] = ra;
        }
        __syncthreads();
        T *dst = grp == 0 ? accb : acca;
        fft_inverse_strict<H>(bufs + (grp < 2 ? grp : 0) * PL, a.tb, t, grp < 2, [&](int i, cplx z) {
            dst[i] = (T)(dst[i] + Torus<T>::native(z.x));
            dst[i + H] = (T)(dst[i + H] + Torus<T>::native(-z.y));
        });
    }

    // ---- output
    if (a.mode == RG_MODE_KMS) {          // fftto!(tacc, acc): :441 / :657
        RawLoad<T, H> ld{grp == 0 ? accb : acca, a.tb.roots};
        fft_forward_strict<H>(bufs + (grp < 2 ? grp : 0) * PL, a.tb, t, grp < 2, ld);
        cplx *out = a.lev_out + (size_t)unit * 2 * H;
        for (int s = tid; s < 2 * H; s += MK_THREADS) out[s] = bufs[(s / H) * PL + PAD(s % H)];
    } else {
        T *out = reinterpret_cast<T *>(a.acc_io) + (size_t)gate * 2 * N;
        for (int i = tid; i < N; i += MK_THREADS) { out[i] = accb[i]; out[N + i] = acca[i]; }
    }
}

template <class T, int H> constexpr size_t rgsw_smem_bytes() {
    return 2 * (2 * H) * sizeof(T) + (size_t)(MK_THREADS / (H / 8)) * padded_len(H) * sizeof(cplx);
}

// ---------------------------------------------------------------------------------------------------
// Shared by phase 2 and CCS: transform the first `nd` gadget digits of `src` (G at a time) and hand each
// spectrum to mac(j, buf) in digit order.  src may be shared or global memory written earlier by this CTA.
template <class T, int H, class Mac>
__device__ __forceinline__ void digits_fft_mac(const T *src, int nd, int l, int logB, cplx *bufs,
                                               const FftTables &tb, Mac &mac) {
    constexpr int GT = H / 8, G = MK_THREADS / GT, PL = padded_len(H);
    const int grp = threadIdx.x / GT, t = threadIdx.x % GT;
    for (int d0 = 0; d0 < nd; d0 += G) {
        const int j = d0 + grp;
        DigitLoad<T, H> ld{src, tb.roots, j, l, logB};
        fft_forward_strict<H>(bufs + grp * PL, tb, t, j < nd, ld);
        for (int g2 = 0; g2 < G; g2++)
            if (d0 + g2 < nd) mac(d0 + g2, bufs + g2 * PL);
        __syncthreads();
    }
}

struct Phase2Args {
    const uint32_t *tilde;        // [B][lwe_words]; only b~ is read
    const cplx *lev;              // [B][R][2][H] phase-1 output, reference slot order
    const cplx *const *rlk;       // [k]: [l_uni][3][H]
    const cplx *const *pubb;      // [k]: [l_uni][H]
    const cplx *crs;              // [l_uni][H]
    FftTables tb;
    uint64_t *acc;                // [B][(k+1)][N] out
    cplx *tx, *ty;                // scratch [B][(k+1)][H] each
    int k, l_lev, logB_lev, l_uni, logB_uni, R, lwe_words;
};

template <int H>
__global__ void __launch_bounds__(MK_THREADS) k_kms_phase2(const Phase2Args a) {
    typedef uint64_t T;
    constexpr int N = 2 * H, GT = H / 8, G = MK_THREADS / GT, SL = H / MK_THREADS, PL = padded_len(H);
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T *P = reinterpret_cast<T *>(smem_raw);
    cplx *bufs = reinterpret_cast<cplx *>(smem_raw + N * sizeof(T));
    const int tid = threadIdx.x, grp = tid / GT, t = tid % GT, gate = blockIdx.x;
    const int k = a.k, ll = a.l_lev, lu = a.l_uni;
    T *ACC = a.acc + (size_t)gate * (k + 1) * N;
    cplx *TX = a.tx + (size_t)gate * (k + 1) * H, *TY = a.ty + (size_t)gate * (k + 1) * H;
    const cplx *LEV = a.lev + (size_t)gate * a.R * 2 * H;

    {   // acc = (test vector, 0 ... 0): bootstrapping.jl:11-23
        const uint32_t tb = a.tilde[(size_t)gate * a.lwe_words];
        for (int i = tid; i < N; i += MK_THREADS) ACC[i] = testvector_coef<T>(i, tb, N);
        for (int i = tid; i < k * N; i += MK_THREADS) ACC[N + i] = 0;
    }
    __syncthreads();

    auto to_P = [&](int i, cplx z) { P[i] = Torus<T>::native(z.x); P[i + H] = Torus<T>::native(-z.y); };

    for (int idx = 0; idx < k; idx++) {
        const cplx *lk = LEV + (size_t)(idx == 0 ? 0 : 1 + (idx - 1) * ll) * 2 * H;
        const cplx *rlk = a.rlk[idx];
        const int iter = idx == 0 ? 1 : ll;                         // :481
        cplx v[SL];
#pragma unroll
        for (int q = 0; q < SL; q++) v[q] = make_double2(0.0, 0.0);

        for (int c = 0; c <= idx; c++) {                            // components b, a_1 .. a_{idx}
            cplx tx[SL], ty[SL], u[SL];
#pragma unroll
            for (int q = 0; q < SL; q++) tx[q] = ty[q] = u[q] = make_double2(0.0, 0.0);
            // LEV product with levkey[idx]: :483-499 (digits beyond `iter` are transformed by the
            // reference but never used)
            auto mac1 = [&](int j, const cplx *buf) {
#pragma unroll
                for (int q = 0; q < SL; q++) {
                    const int s = tid + q * MK_THREADS;
                    const cplx x = buf[PAD(s)];
                    tx[q] = cadd_s(tx[q], cmul_s(x, lk[(size_t)(j * 2 + 0) * H + s]));
                    ty[q] = cadd_s(ty[q], cmul_s(x, lk[(size_t)(j * 2 + 1) * H + s]));
                }
            };
            digits_fft_mac<T, H>(ACC + (size_t)c * N, iter, ll, a.logB_lev, bufs, a.tb, mac1);
#pragma unroll
            for (int q = 0; q < SL; q++) {
                const int s = tid + q * MK_THREADS;
                TX[(size_t)c * H + s] = tx[q];
                bufs[PAD(s)] = ty[q];
            }
            __syncthreads();
            fft_inverse_strict<H>(bufs, a.tb, t, grp == 0, to_P);   // y_c : :501-504
            // u and v : :520-535
            auto mac2 = [&](int j, const cplx *buf) {
#pragma unroll
                for (int q = 0; q < SL; q++) {
                    const int s = tid + q * MK_THREADS;
                    const cplx x = buf[PAD(s)];
                    u[q] = cadd_s(u[q], cmul_s(x, __ldg(&rlk[(size_t)(j * 3 + 0) * H + s])));
                    if (c == 0) v[q] = csub_s(v[q], cmul_s(x, __ldg(&a.crs[(size_t)j * H + s])));
                    else v[q] = cadd_s(v[q], cmul_s(x, __ldg(&a.pubb[c - 1][(size_t)j * H + s])));
                }
            };
            digits_fft_mac<T, H>(P, lu, lu, a.logB_uni, bufs, a.tb, mac2);
#pragma unroll
            for (int q = 0; q < SL; q++) TY[(size_t)c * H + tid + q * MK_THREADS] = u[q];
        }
        // v -> coefficient form -> digits -> w : :538-550
#pragma unroll
        for (int q = 0; q < SL; q++) bufs[PAD(tid + q * MK_THREADS)] = v[q];
        __syncthreads();
        fft_inverse_strict<H>(bufs, a.tb, t, grp == 0, to_P);
        cplx wb[SL], wa[SL];
#pragma unroll
        for (int q = 0; q < SL; q++) { wb[q] = TY[tid + q * MK_THREADS]; wa[q] = make_double2(0.0, 0.0); }
        auto mac3 = [&](int j, const cplx *buf) {
#pragma unroll
            for (int q = 0; q < SL; q++) {
                const int s = tid + q * MK_THREADS;
                const cplx x = buf[PAD(s)];
                wb[q] = cadd_s(wb[q], cmul_s(x, __ldg(&rlk[(size_t)(j * 3 + 1) * H + s])));
                wa[q] = cadd_s(wa[q], cmul_s(x, __ldg(&rlk[(size_t)(j * 3 + 2) * H + s])));
            }
        };
        digits_fft_mac<T, H>(P, lu, lu, a.logB_uni, bufs, a.tb, mac3);
        // tx += ty ; acc = ifft(tx) for every component (:553-556); components beyond idx+1 stay 0
        for (int c0 = 0; c0 <= idx + 1; c0 += G) {
            for (int g2 = 0; g2 < G; g2++) {
                const int c = c0 + g2;
                if (c > idx + 1) break;
#pragma unroll
                for (int q = 0; q < SL; q++) {
                    const int s = tid + q * MK_THREADS;
                    cplx val;
                    if (c == 0) val = cadd_s(TX[s], wb[q]);
                    else if (c == idx + 1) val = cadd_s(make_double2(0.0, 0.0), wa[q]);
                    else val = cadd_s(TX[(size_t)c * H + s], TY[(size_t)c * H + s]);
                    bufs[g2 * PL + PAD(s)] = val;
                }
            }
            __syncthreads();
            const int c = c0 + grp;
            T *dst = ACC + (size_t)(c <= idx + 1 ? c : 0) * N;
            fft_inverse_strict<H>(bufs + grp * PL, a.tb, t, c <= idx + 1, [&](int i, cplx z) {
                dst[i] = Torus<T>::native(z.x);
                dst[i + H] = Torus<T>::native(-z.y);
            });
        }
    }
}

template <int H> constexpr size_t phase2_smem_bytes() {
    return (size_t)(2 * H) * sizeof(uint64_t) + (size_t)(MK_THREADS / (H / 8)) * padded_len(H) * sizeof(cplx);
}

// ---------------------------------------------------------------------------------------------------
struct CcsArgs {
    const uint32_t *tilde;        // [B][lwe_words]
    const cplx *const *brk;       // [k]: [n][l_uni][3][H]
    const cplx *const *pubb;      // [k]: [l_uni][H]
    const cplx *crs;              // [l_uni][H]
    const cplx *mono;
    FftTables tb;
    uint32_t *acc;                // [B][(k+1)][N] out
    uint32_t *vscr;               // scratch [B][(k+1)][N]
    cplx *tacc;                   // scratch [B][(k+1)][H]
    int n, k, l_uni, logB_uni, lwe_words;
};

template <int H>
__global__ void __launch_bounds__(MK_THREADS) k_ccs_blindrotate(const CcsArgs a) {
    typedef uint32_t T;
    constexpr int N = 2 * H, GT = H / 8, G = MK_THREADS / GT, SL = H / MK_THREADS, PL = padded_len(H);
    extern __shared__ __align__(16) unsigned char smem_raw[];
    cplx *bufs = reinterpret_cast<cplx *>(smem_raw);
    const int tid = threadIdx.x, grp = tid / GT, t = tid % GT, gate = blockIdx.x;
    const int k = a.k, lu = a.l_uni, lb = a.logB_uni;
    T *ACC = a.acc + (size_t)gate * (k + 1) * N, *V = a.vscr + (size_t)gate * (k + 1) * N;
    cplx *TACC = a.tacc + (size_t)gate * (k + 1) * H;
    const uint32_t *tilde = a.tilde + (size_t)gate * a.lwe_words;

    for (int i = tid; i < N; i += MK_THREADS) ACC[i] = testvector_coef<T>(i, tilde[0], N);
    for (int i = tid; i < k * N; i += MK_THREADS) ACC[N + i] = 0;
    __syncthreads();

    for (int idx = 0; idx < k; idx++) {
        for (int i = 0; i < a.n; i++) {
            const uint32_t at = tilde[1 + (size_t)idx * a.n + i];
            if (at == 0) continue;                                     // :262
            const cplx *uni = a.brk[idx] + (size_t)i * 3 * lu * H;
            const int na = idx + 1;                                    // live a-components
            // pass 1: u_c (:277-284) and v_c (:286-300) for every live component
            for (int c = 0; c <= na; c++) {
                cplx u[SL], v[SL];
#pragma unroll
                for (int q = 0; q < SL; q++) u[q] = v[q] = make_double2(0.0, 0.0);
                auto mac = [&](int j, const cplx *buf) {
#pragma unroll
                    for (int q = 0; q < SL; q++) {
                        const int s = tid + q * MK_THREADS;
                        const cplx x = buf[PAD(s)];
                        u[q] = cadd_s(u[q], cmul_s(x, __ldg(&uni[(size_t)(j * 3 + 0) * H + s])));
                        if (c == 0) v[q] = csub_s(v[q], cmul_s(x, __ldg(&a.crs[(size_t)j * H + s])));
                        else v[q] = cadd_s(v[q], cmul_s(x, __ldg(&a.pubb[c - 1][(size_t)j * H + s])));
                    }
                };
                digits_fft_mac<T, H>(ACC + (size_t)c * N, lu, lu, lb, bufs, a.tb, mac);
#pragma unroll
                for (int q = 0; q < SL; q++) {
                    const int s = tid + q * MK_THREADS;
                    TACC[(size_t)c * H + s] = u[q];
                    bufs[PAD(s)] = v[q];
                }
                __syncthreads();
                T *dst = V + (size_t)c * N;
                fft_inverse_strict<H>(bufs, a.tb, t, grp == 0, [&](int ii, cplx z) {
                    dst[ii] = Torus<T>::native(z.x);
                    dst[ii + H] = Torus<T>::native(-z.y);
                });
            }
            // pass 2: w (:313-320) on top of u_b and u_{a_idx}
            cplx wb[SL], wa[SL];
#pragma unroll
            for (int q = 0; q < SL; q++) {
                wb[q] = TACC[tid + q * MK_THREADS];
                wa[q] = TACC[(size_t)na * H + tid + q * MK_THREADS];
            }
            for (int c = 0; c <= na; c++) {
                auto mac = [&](int j, const cplx *buf) {
#pragma unroll
                    for (int q = 0; q < SL; q++) {
                        const int s = tid + q * MK_THREADS;
                        const cplx x = buf[PAD(s)];
                        wb[q] = cadd_s(wb[q], cmul_s(x, __ldg(&uni[(size_t)(j * 3 + 1) * H + s])));
                        wa[q] = cadd_s(wa[q], cmul_s(x, __ldg(&uni[(size_t)(j * 3 + 2) * H + s])));
                    }
                };
                digits_fft_mac<T, H>(V + (size_t)c * N, lu, lu, lb, bufs, a.tb, mac);
            }
            // acc += ifft(monomial * tacc) (:322-324); components beyond na hold tacc = 0
            const cplx *mo = a.mono + (size_t)(at - 1) * H;
            for (int c0 = 0; c0 <= na; c0 += G) {
                for (int g2 = 0; g2 < G; g2++) {
                    const int c = c0 + g2;
                    if (c > na) break;
#pragma unroll
                    for (int q = 0; q < SL; q++) {
                        const int s = tid + q * MK_THREADS;
                        const cplx val = c == 0 ? wb[q] : (c == na ? wa[q] : TACC[(size_t)c * H + s]);
                        bufs[g2 * PL + PAD(s)] = cmul_s(__ldg(&mo[s]), val);
                    }
                }
                __syncthreads();
                const int c = c0 + grp;
                T *dst = ACC + (size_t)(c <= na ? c : 0) * N;
                fft_inverse_strict<H>(bufs + grp * PL, a.tb, t, c <= na, [&](int ii, cplx z) {
                    dst[ii] = (T)(dst[ii] + Torus<T>::native(z.x));
                    dst[ii + H] = (T)(dst[ii + H] + Torus<T>::native(-z.y));
                });
            }
        }
    }
}

template <int H> constexpr size_t ccs_smem_bytes() {
    return (size_t)(MK_THREADS / (H / 8)) * padded_len(H) * sizeof(cplx);
}

// ---------------------------------------------------------------------------------------------------
// Unit kernels for the parity hooks and for building the monomial table.

// scheme.jl:121-146: entry a-1 = FFT(X^a - 1), a = 1..2N-1; entry 2N-1 = 0.  One group per entry.
template <int H>
__global__ void __launch_bounds__(MK_THREADS) k_build_monomials(cplx *out, FftTables tb) {
    constexpr int N = 2 * H, GT = H / 8, G = MK_THREADS / GT, PL = padded_len(H);
    extern __shared__ __align__(16) unsigned char smem_raw[];
    cplx *bufs = reinterpret_cast<cplx *>(smem_raw);
    const int grp = threadIdx.x / GT, t = threadIdx.x % GT;
    const int a = blockIdx.x * G + grp + 1;                    // 1 .. 2N
    const bool active = a < 2 * N;
    auto coef = [&](int i) -> int {                            // coefficient i of the reference's tmppoly
        if (a < N) return i == 0 ? -1 : (i == a ? 1 : 0);
        if (a == N) return i == 0 ? -2 : 0;
        return i == 0 ? -1 : (i == a - N ? -1 : 0);
    };
    fft_forward_strict<H>(bufs + grp * PL, tb, t, active, [&](int i) {
        return cmul_s(make_double2((double)coef(i), (double)(-coef(i + H))), __ldg(&tb.roots[i]));
    });
    if (a <= 2 * N) {
        cplx *dst = out + (size_t)(a - 1) * H;
        for (int s = t; s < H; s += GT) dst[s] = active ? bufs[grp * PL + PAD(s)] : make_double2(0.0, 0.0);
    }
}

template <class T, int H>
__global__ void __launch_bounds__(MK_THREADS) k_fft_batch(const T *polys, cplx *out, FftTables tb, int batch) {
    constexpr int N = 2 * H, GT = H / 8, G = MK_THREADS / GT, PL = padded_len(H);
    extern __shared__ __align__(16) unsigned char smem_raw[];
    cplx *bufs = reinterpret_cast<cplx *>(smem_raw);
    const int grp = threadIdx.x / GT, t = threadIdx.x % GT;
    const int p = blockIdx.x * G + grp;
    const bool active = p < batch;
    RawLoad<T, H> ld{polys + (size_t)(active ? p : 0) * N, tb.roots};
    fft_forward_strict<H>(bufs + grp * PL, tb, t, active, ld);
    if (active) for (int s = t; s < H; s += GT) out[(size_t)p * H + s] = bufs[grp * PL + PAD(s)];
}

template <class T, int H>
__global__ void __launch_bounds__(MK_THREADS) k_ifft_batch(const cplx *in, T *polys, FftTables tb, int batch) {
    constexpr int N = 2 * H, GT = H / 8, G = MK_THREADS / GT, PL = padded_len(H);
    extern __shared__ __align__(16) unsigned char smem_raw[];
    cplx *bufs = reinterpret_cast<cplx *>(smem_raw);
    const int grp = threadIdx.x / GT, t = threadIdx.x % GT;
    const int p = blockIdx.x * G + grp;
    const bool active = p < batch;
    if (active) for (int s = t; s < H; s += GT) bufs[grp * PL + PAD(s)] = in[(size_t)p * H + s];
    __syncthreads();
    T *dst = polys + (size_t)(active ? p : 0) * N;
    fft_inverse_strict<H>(bufs + grp * PL, tb, t, active, [&](int i, cplx z) {
        dst[i] = Torus<T>::native(z.x);
        dst[i + H] = Torus<T>::native(-z.y);
    });
}

// gsw.jl:86-96 on a batch: digits [batch][l][N], stored wrapped like the reference.
template <class T>
__global__ void k_decomp_batch(const T *polys, T *digits, int N, int l, int logB, size_t total) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const size_t p = i / N, c = i % N;
    const T v = polys[i];
    for (int j = 0; j < l; j++)
        digits[(p * l + j) * N + c] = (T)(typename Torus<T>::S)gadget_digit<T>(v, j, l, logB);
}
