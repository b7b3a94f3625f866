// capi.cu -- C-ABI of include/mktfhe_b200.h: context, key upload, batch pipeline, parity hooks.
#include "../../include/mktfhe_b200.h"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <string>
#include <thread>
#include <vector>
#include <quadmath.h>

#include "common.cuh"
#include "fft_strict.cuh"
#include "kernels_strict.cuh"
#include "keyswitch.cuh"
#include "kernels_fast.cuh"
#include "kernels_fast32.cuh"
#include "kernels_fast32_w.cuh"
#include "keygen.cuh"

namespace {

std::string g_create_error;

struct StageEvents { cudaEvent_t e[MKTFHE_STAGE_COUNT + 1]; };

}  // namespace

struct mktfhe_ctx {
    mktfhe_params p;
    int device = 0, mode = MKTFHE_MODE_STRICT, sms = 148;
    int N = 0, H = 0, bits = 32, R = 1, nparties = 1;
    bool mk = false, kms = false, block = false;
    cudaStream_t stream = nullptr;
    // tables / keys (device)
    cplx *psi = nullptr, *psiinv = nullptr, *roots = nullptr, *rootsinv = nullptr, *mono = nullptr;
    std::vector<cplx *> brk, rlk, pubb;
    std::vector<uint32_t *> ksk;
    cplx *crs = nullptr;
    void *crs_coeff = nullptr;                       // device key generation: CRS in coefficient form, [l_uni][N] torus
    cplx **d_brk = nullptr, **d_rlk = nullptr, **d_pubb = nullptr;
    uint32_t **d_ksk = nullptr;
    bool finalized = false;
    FastKeys fast;
    FastKeys32 fast32;
    FastKeys32W fast32w;           // CGGI, half-warp transform (kernels_fast32_w.cuh)
    FastCcsKeys fastccs;
    // workspace for `cap` gates
    size_t cap = 0;
    uint32_t *w_in1 = nullptr, *w_in2 = nullptr, *w_out = nullptr, *w_lin = nullptr, *w_tilde = nullptr, *w_v = nullptr;
    void *w_acc = nullptr;
    cplx *w_lev = nullptr, *w_tx = nullptr, *w_ty = nullptr;
    // circuit wire table (mktfhe_wires_*): nwires LWE records, plus per-level index scratch for `idx_cap` gates
    uint32_t *wires = nullptr;
    size_t nwires = 0, idx_cap = 0;
    int32_t *w_idx = nullptr;                        // [4][idx_cap]: ops, src1, src2, dst
    // measurement
    std::vector<StageEvents> events;
    size_t events_used = 0;
    int launches = 0;
    std::string err;
    size_t mem_budget = (size_t)24 << 30;
    // multi-device front (mktfhe_ctx_create_multi): one child context per device, this object only dispatches
    std::vector<mktfhe_ctx *> children;

    FftTables tables() const { return FftTables{psi, psiinv, roots, rootsinv}; }
};

namespace {

int fail(mktfhe_ctx *c, int code, const std::string &msg) {
    if (c) c->err = msg; else g_create_error = msg;
    return code;
}
#define CK(call)                                                                                   \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess)                                                                     \
            return fail(ctx, MKTFHE_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_)); \
    } while (0)

template <class T> void dfree(T *&p) { if (p) { cudaFree(p); p = nullptr; } }
// temporary device buffer of a hook: released on every exit path (CK returns early on the first CUDA error)
template <class T> struct DevTmp {
    T *p = nullptr;
    DevTmp() = default;
    DevTmp(const DevTmp &) = delete;
    DevTmp &operator=(const DevTmp &) = delete;
    ~DevTmp() { if (p) cudaFree(p); }
    cudaError_t alloc(size_t bytes) { return cudaMalloc(&p, bytes); }
    operator T *() const { return p; }
};

size_t per_gate_bytes(const mktfhe_ctx *c) {
    const mktfhe_params &p = c->p;
    size_t b = 5 * mktfhe_lwe_words(&p) * 4;                                   // in1, in2, out, lin, tilde
    b += (size_t)(p.k + 1) * c->N * (c->bits / 8);                             // acc
    if (c->kms) b += (size_t)c->R * 2 * c->H * 16 + 2 * (size_t)(p.k + 1) * c->H * 16;
    if (p.scheme == MKTFHE_CCS) b += (size_t)(p.k + 1) * c->N * 4 + (size_t)(p.k + 1) * c->H * 16;
    return b;
}

void free_workspace(mktfhe_ctx *c) {
    dfree(c->w_in1); dfree(c->w_in2); dfree(c->w_out); dfree(c->w_lin); dfree(c->w_tilde); dfree(c->w_v);
    dfree(c->w_acc); dfree(c->w_lev); dfree(c->w_tx); dfree(c->w_ty);
    c->cap = 0;
}

void free_wires(mktfhe_ctx *c) {
    dfree(c->wires); dfree(c->w_idx);
    c->nwires = c->idx_cap = 0;
}

int ensure_workspace(mktfhe_ctx *ctx, size_t gates) {
    if (gates <= ctx->cap) return 0;
    free_workspace(ctx);
    const mktfhe_params &p = ctx->p;
    const size_t lw = mktfhe_lwe_words(&p);
    CK(cudaMalloc(&ctx->w_in1, gates * lw * 4));
    CK(cudaMalloc(&ctx->w_in2, gates * lw * 4));
    CK(cudaMalloc(&ctx->w_out, gates * lw * 4));
    CK(cudaMalloc(&ctx->w_lin, gates * lw * 4));
    CK(cudaMalloc(&ctx->w_tilde, gates * lw * 4));
    CK(cudaMalloc(&ctx->w_acc, gates * (size_t)(p.k + 1) * ctx->N * (ctx->bits / 8)));
    if (ctx->kms) {
        CK(cudaMalloc(&ctx->w_lev, gates * (size_t)ctx->R * 2 * ctx->H * sizeof(cplx)));
        CK(cudaMalloc(&ctx->w_tx, gates * (size_t)(p.k + 1) * ctx->H * sizeof(cplx)));
        CK(cudaMalloc(&ctx->w_ty, gates * (size_t)(p.k + 1) * ctx->H * sizeof(cplx)));
    }
    if (p.scheme == MKTFHE_CCS) {
        CK(cudaMalloc(&ctx->w_v, gates * (size_t)(p.k + 1) * ctx->N * 4));
        CK(cudaMalloc(&ctx->w_tx, gates * (size_t)(p.k + 1) * ctx->H * sizeof(cplx)));
    }
    ctx->cap = gates;
    return 0;
}

size_t chunk_gates(const mktfhe_ctx *c, size_t batch) {
    const size_t lim = c->mem_budget / per_gate_bytes(c);
    return batch < lim ? batch : (lim ? lim : 1);
}

// ---- kernel launchers (one per stage) ---------------------------------------------------------------
template <class T, int H, int ELL>
int launch_rgsw(mktfhe_ctx *ctx, const RgswArgs &a, size_t units) {
    auto kern = k_rgsw_blindrotate<T, H, ELL>;
    const size_t smem = rgsw_smem_bytes<T, H>();
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<(unsigned)units, MK_THREADS, smem, ctx->stream>>>(a);
    ctx->launches++;
    CK(cudaGetLastError());
    return 0;
}

int run_rgsw(mktfhe_ctx *ctx, RgswArgs a, size_t units) {
    const mktfhe_params &p = ctx->p;
    a.brk = ctx->d_brk; a.mono = ctx->mono; a.tb = ctx->tables();
    a.n = p.n; a.d = p.d; a.k = p.k; a.l = p.l_gsw; a.logB = p.logB_gsw;
    a.l_lev = p.l_lev; a.logB_lev = p.logB_lev; a.R = ctx->R; a.lwe_words = (int)mktfhe_lwe_words(&p);
    const bool blk = ctx->block && (a.mode != RG_MODE_STEP || a.step_block);
    if (ctx->bits == 64) return blk ? launch_rgsw<uint64_t, 1024, 3>(ctx, a, units) : launch_rgsw<uint64_t, 1024, 1>(ctx, a, units);
    return blk ? launch_rgsw<uint32_t, 512, 3>(ctx, a, units) : launch_rgsw<uint32_t, 512, 1>(ctx, a, units);
}

int run_prep(mktfhe_ctx *ctx, int op, const uint32_t *in1, const uint32_t *in2, uint32_t *lin, uint32_t *tilde, size_t gates,
             const int32_t *ops = nullptr, const int32_t *idx1 = nullptr, const int32_t *idx2 = nullptr) {
    const size_t total = gates * mktfhe_lwe_words(&ctx->p);
    int logN = 0; while ((1 << logN) < ctx->N) logN++;
    k_gate_prep<<<(unsigned)((total + 255) / 256), 256, 0, ctx->stream>>>(in1, in2, lin, tilde, op, (int)mktfhe_lwe_words(&ctx->p), total, 32 - logN - 1,
                                                                         ops, idx1, idx2);
    ctx->launches++;
    CK(cudaGetLastError());
    return 0;
}

int run_phase1(mktfhe_ctx *ctx, const uint32_t *tilde, cplx *lev, size_t gates) {
    if (ctx->mode == MKTFHE_MODE_FAST && fast_supported(ctx->p)) return fast_phase1(ctx->fast, ctx->p, tilde, lev, gates, ctx->stream, &ctx->launches, ctx->err);
    RgswArgs a{};
    a.tilde = tilde; a.lev_out = lev; a.mode = RG_MODE_KMS;
    return run_rgsw(ctx, a, gates * ctx->R);
}

int run_phase2(mktfhe_ctx *ctx, const uint32_t *tilde, const cplx *lev, uint64_t *acc, size_t gates) {
    const mktfhe_params &p = ctx->p;
    Phase2Args a{};
    a.tilde = tilde; a.lev = lev; a.rlk = ctx->d_rlk; a.pubb = ctx->d_pubb; a.crs = ctx->crs; a.tb = ctx->tables();
    a.acc = acc; a.tx = ctx->w_tx; a.ty = ctx->w_ty;
    a.k = p.k; a.l_lev = p.l_lev; a.logB_lev = p.logB_lev; a.l_uni = p.l_uni; a.logB_uni = p.logB_uni;
    a.R = ctx->R; a.lwe_words = (int)mktfhe_lwe_words(&p);
    auto kern = k_kms_phase2<1024>;
    const size_t smem = phase2_smem_bytes<1024>();
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<(unsigned)gates, MK_THREADS, smem, ctx->stream>>>(a);
    ctx->launches++;
    CK(cudaGetLastError());
    return 0;
}

int run_ccs(mktfhe_ctx *ctx, const uint32_t *tilde, uint32_t *acc, size_t gates) {
    const mktfhe_params &p = ctx->p;
    CcsArgs a{};
    a.tilde = tilde; a.brk = ctx->d_brk; a.pubb = ctx->d_pubb; a.crs = ctx->crs; a.mono = ctx->mono; a.tb = ctx->tables();
    a.acc = acc; a.vscr = ctx->w_v; a.tacc = ctx->w_tx;
    a.n = p.n; a.k = p.k; a.l_uni = p.l_uni; a.logB_uni = p.logB_uni; a.lwe_words = (int)mktfhe_lwe_words(&p);
    auto kern = k_ccs_blindrotate<512>;
    const size_t smem = ccs_smem_bytes<512>();
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<(unsigned)gates, MK_THREADS, smem, ctx->stream>>>(a);
    ctx->launches++;
    CK(cudaGetLastError());
    return 0;
}

int run_keyswitch(mktfhe_ctx *ctx, const void *acc, uint32_t *out, size_t gates) {
    const mktfhe_params &p = ctx->p;
    KsArgs a{};
    a.acc = acc; a.ksk = ctx->d_ksk; a.out = out;
    a.N = ctx->N; a.n = p.n; a.k = p.k; a.f = p.f; a.logD = p.logD; a.Dk = mktfhe_ksk_rows(&p);
    a.bits64 = ctx->bits == 64; a.block = ctx->block; a.rowp = (p.n + 1 + 3) / 4 * 4;
    if (ctx->mode == MKTFHE_MODE_STRICT || p.f * p.logD != 16 || p.logD != 2 || p.n + 1 > 128 * KS_J) {
        // reference loop order: one gate per CTA, parties in sequence
        const size_t smem = keyswitch_smem_bytes(ctx->N, p.f, p.n);
        k_keyswitch<<<(unsigned)gates, MK_THREADS, smem, ctx->stream>>>(a);
    } else {
        CK(cudaMemsetAsync(out, 0, gates * mktfhe_lwe_words(&p) * 4, ctx->stream));
        auto kern = ctx->block ? k_keyswitch_tiled<true> : k_keyswitch_tiled<false>;
        // Split the coefficient range over blockIdx.z so that the grid fills whole waves: cost of a split = rounds x
        // co-resident CTAs x coefficients per CTA (an SM's row throughput is shared by its resident CTAs).
        const int c_total = ctx->N - (ctx->block ? p.n : 0);
        const size_t tiles = (gates + KS_G - 1) / KS_G;
        CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)keyswitch_tiled_smem(c_total, a.rowp)));
        int best_split = 1;
        double best_cost = 1e300;
        for (int split = 1; split <= 8; split++) {
            const int c_per = (c_total + split - 1) / split;
            int resident = 0;
            CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&resident, kern, MK_THREADS, keyswitch_tiled_smem(c_per, a.rowp)));
            if (resident < 1) continue;
            const double ctas = (double)tiles * p.k * split, slots = (double)resident * ctx->sms;
            const double rounds = std::ceil(ctas / slots);
            const double cost = rounds * std::min((double)resident, std::ceil(ctas / ctx->sms)) * c_per;
            if (cost < best_cost * 0.999) { best_cost = cost; best_split = split; }
        }
        const int c_per = (c_total + best_split - 1) / best_split;
        const dim3 grid((unsigned)tiles, (unsigned)p.k, (unsigned)best_split);
        kern<<<grid, MK_THREADS, keyswitch_tiled_smem(c_per, a.rowp), ctx->stream>>>(a, (int)gates, best_split);
    }
    ctx->launches++;
    CK(cudaGetLastError());
    return 0;
}

// blind rotation of `gates` ciphertexts whose tilde is in w_tilde -> w_acc
// N = 1024 single-key schemes in FAST mode: CGGI runs the half-warp-transform kernel (kernels_fast32_w.cuh), LMSS the 64-thread
// kernel with the TMA key ring (kernels_fast32.cuh).  MKTFHE_FAST32_KERNEL = tmem | tma forces the 64-thread kernels for CGGI too.
int launch_fast32(mktfhe_ctx *ctx, const fast32::Args &fa) {
    static const bool force64 = []() { const char *e = getenv("MKTFHE_FAST32_KERNEL"); return e && (std::string(e) == "tmem" || std::string(e) == "tma"); }();
    if (ctx->p.scheme == MKTFHE_CGGI && ctx->fast32w.built && !force64) {
        fastw32::Args wa{};
        wa.tilde = fa.tilde; wa.acc_io = fa.acc_io; wa.step_mode = fa.step_mode; wa.step_idx = fa.step_idx; wa.units = fa.units;
        return fast32w_launch(ctx->fast32w, ctx->fast32.emono, ctx->p, wa, ctx->stream, &ctx->launches, ctx->err);
    }
    return fast32_launch(ctx->fast32, ctx->p, fa, ctx->stream, &ctx->launches, ctx->err);
}

int run_blindrotate(mktfhe_ctx *ctx, const uint32_t *tilde, size_t gates, StageEvents *ev) {
    const mktfhe_params &p = ctx->p;
    int rc;
    if (ctx->kms) {
        // FAST: both phases in the production kernels; the RLEV rows travel between them in thread order
        const bool fast2 = ctx->mode == MKTFHE_MODE_FAST && fast_supported(p) && fast_variant_tma();
        if (fast2) rc = fast_phase1(ctx->fast, p, tilde, ctx->w_lev, gates, ctx->stream, &ctx->launches, ctx->err, true);
        else rc = run_phase1(ctx, tilde, ctx->w_lev, gates);
        if (rc) return rc;
        if (ev) CK(cudaEventRecord(ev->e[2], ctx->stream));
        if (fast2) rc = fast_phase2(ctx->fast, p, tilde, ctx->w_lev, (uint64_t *)ctx->w_acc, ctx->w_tx, ctx->w_ty, gates, ctx->stream, &ctx->launches, ctx->err);
        else rc = run_phase2(ctx, tilde, ctx->w_lev, (uint64_t *)ctx->w_acc, gates);
        if (rc) return rc;
    } else if (p.scheme == MKTFHE_CCS) {
        if (ctx->mode == MKTFHE_MODE_FAST && fastccs_supported(p))
            rc = fastccs_launch(ctx->fastccs, ctx->fast32, p, tilde, (uint32_t *)ctx->w_acc, ctx->w_tx, gates, ctx->stream, &ctx->launches, ctx->err);
        else
            rc = run_ccs(ctx, tilde, (uint32_t *)ctx->w_acc, gates);
        if (rc) return rc;
        if (ev) CK(cudaEventRecord(ev->e[2], ctx->stream));
    } else {
        if (ctx->mode == MKTFHE_MODE_FAST && fast32_supported(p)) {
            fast32::Args fa{};
            fa.tilde = tilde; fa.acc_io = (uint32_t *)ctx->w_acc; fa.step_mode = 0; fa.units = gates;
            if ((rc = launch_fast32(ctx, fa))) return rc;
        } else {
            RgswArgs a{};
            a.tilde = tilde; a.acc_io = ctx->w_acc; a.mode = RG_MODE_SK;
            if ((rc = run_rgsw(ctx, a, gates))) return rc;
        }
        if (ev) CK(cudaEventRecord(ev->e[2], ctx->stream));
    }
    return 0;
}

int check_ready(mktfhe_ctx *ctx) {
    if (!ctx) return MKTFHE_ERR_ARG;
    if (!ctx->finalized) return fail(ctx, MKTFHE_ERR_STATE, "keys not finalized: call mktfhe_finalize_keys first");
    cudaError_t e = cudaSetDevice(ctx->device);
    if (e != cudaSuccess) return fail(ctx, MKTFHE_ERR_CUDA, std::string("cudaSetDevice: ") + cudaGetErrorString(e));
    return 0;
}

StageEvents *next_events(mktfhe_ctx *ctx) {
    if (ctx->events_used == ctx->events.size()) {
        StageEvents s;
        for (auto &e : s.e) if (cudaEventCreate(&e) != cudaSuccess) return nullptr;
        ctx->events.push_back(s);
    }
    return &ctx->events[ctx->events_used++];
}

void fft_tables_host(int N, std::vector<cplx> &psi, std::vector<cplx> &psiinv, std::vector<cplx> &roots, std::vector<cplx> &rootsinv) {
    // fft.jl:26-44.  BigFloat in the reference, binary128 here; both round to the same Float64.
    const int H = N / 2;
    psi.resize(H); psiinv.resize(H); roots.resize(H); rootsinv.resize(H);
    for (int j = 0; j < H; j++) {
        const __float128 pi = acosq((__float128)-1), th = pi * j / H, ph = pi * j / N;
        psi[j] = make_double2((double)cosq(th), (double)-sinq(th));
        psiinv[j] = make_double2((double)cosq(th), (double)sinq(th));
        roots[j] = make_double2((double)cosq(ph), (double)sinq(ph));
        rootsinv[j] = make_double2((double)(cosq(ph) / H), (double)(-sinq(ph) / H));
    }
    auto bitrev = [&](std::vector<cplx> &v) {
        for (int i = 1, j = 0; i < H; i++) {
            int bit = H >> 1;
            for (; j >= bit; bit >>= 1) j -= bit;
            j += bit;
            if (i < j) std::swap(v[i], v[j]);
        }
    };
    bitrev(psi); bitrev(psiinv);
}

template <class T> int upload(mktfhe_ctx *ctx, T *&dst, const void *src, size_t bytes) {
    dfree(dst);
    CK(cudaMalloc(&dst, bytes));
    CK(cudaMemcpy(dst, src, bytes, cudaMemcpyDefault));   // host or device source (UVA)
    return 0;
}

// transform tables of the reference (fft.jl:26-44); needed by finalize and by the device key generation
int ensure_tables(mktfhe_ctx *ctx) {
    if (ctx->psi && ctx->psiinv && ctx->roots && ctx->rootsinv) return 0;
    int rc;
    std::vector<cplx> psi, psiinv, roots, rootsinv;
    fft_tables_host(ctx->N, psi, psiinv, roots, rootsinv);
    const size_t tb = sizeof(cplx) * ctx->H;
    if ((rc = upload(ctx, ctx->psi, psi.data(), tb)) || (rc = upload(ctx, ctx->psiinv, psiinv.data(), tb)) ||
        (rc = upload(ctx, ctx->roots, roots.data(), tb)) || (rc = upload(ctx, ctx->rootsinv, rootsinv.data(), tb)))
        return rc;
    return 0;
}

// ---- device key generation (csrc/keygen.cuh) -------------------------------------------------------------------
kg::Key kg_key(uint64_t seed, const uint8_t *key32) {
    kg::Key k;
    if (key32) { memcpy(k.k, key32, 32); return k; }
    uint64_t x = seed;                                  // splitmix64 expansion, as ChaCha20::Key::from_seed (host_keygen.cpp)
    for (int i = 0; i < 4; i++) {
        uint64_t z = (x += 0x9E3779B97F4A7C15ull);
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
        z ^= z >> 31;
        k.k[2 * i] = (uint32_t)z; k.k[2 * i + 1] = (uint32_t)(z >> 32);
    }
    return k;
}

template <class T, int H> int kg_fft(mktfhe_ctx *ctx, const T *polys, cplx *out, size_t count) {
    const int G = MK_THREADS / (H / 8);
    const size_t smem = (size_t)G * padded_len(H) * sizeof(cplx);
    k_fft_batch<T, H><<<(unsigned)((count + G - 1) / G), MK_THREADS, smem, ctx->stream>>>(polys, out, ctx->tables(), (int)count);
    CK(cudaGetLastError());
    return 0;
}

template <class T, int NN> int kg_common(mktfhe_ctx *ctx, const kg::Key &key) {
    const mktfhe_params &p = ctx->p;
    const size_t count = (size_t)p.l_uni * NN;
    dfree(ctx->crs_coeff); dfree(ctx->crs);
    CK(cudaMalloc(&ctx->crs_coeff, count * sizeof(T)));
    CK(cudaMalloc(&ctx->crs, mktfhe_crs_doubles(&p) * 8));
    kg::k_kg_uniform<T><<<(unsigned)std::min<size_t>((count + 2047) / 2048, 1024), 256, 0, ctx->stream>>>(key, kg::stream_id(kg::S_CRS, 0, 0), (T *)ctx->crs_coeff, count);
    CK(cudaGetLastError());
    return kg_fft<T, NN / 2>(ctx, (const T *)ctx->crs_coeff, ctx->crs, p.l_uni);
}

template <class T, int NN> int kg_party(mktfhe_ctx *ctx, int party, const kg::Key &key) {
    using Row = kg::RowDesc<T>;
    const mktfhe_params &p = ctx->p;
    constexpr int WT = sizeof(T) / 4, H = NN / 2;
    const int n = p.n, bits = (int)sizeof(T) * 8;
    const bool ccs = p.scheme == MKTFHE_CCS;
    int rc;
    int8_t *d_lwe = nullptr, *d_ring = nullptr, *d_gsw = nullptr, *d_r = nullptr;
    T *tmp = nullptr;
    Row *d_rows = nullptr;
    auto cleanup = [&]() { dfree(d_lwe); dfree(d_ring); dfree(d_gsw); dfree(d_r); dfree(tmp); dfree(d_rows); };
#define KG(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { cleanup(); return fail(ctx, MKTFHE_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_)); } } while (0)
    KG(cudaMalloc(&d_lwe, n)); KG(cudaMalloc(&d_ring, NN)); KG(cudaMalloc(&d_gsw, NN));
    kg::k_kg_secrets<<<1, 256, (size_t)NN * 4, ctx->stream>>>(key, party, n, NN, ctx->block ? 1 : 0, p.d, p.ell, d_lwe, d_ring, d_gsw);
    KG(cudaGetLastError());
    std::vector<int8_t> h_lwe(n);
    KG(cudaMemcpyAsync(h_lwe.data(), d_lwe, n, cudaMemcpyDeviceToHost, ctx->stream));
    KG(cudaStreamSynchronize(ctx->stream));
    const int8_t *brk_key = ctx->kms ? d_gsw : d_ring;          // RGSW key: gswkey for KMS*, ringkey for CGGI / LMSS
    auto gvec = [&](int j, int logB) -> T { return (T)1 << (bits - (j + 1) * logB); };
    KG(cudaFuncSetAttribute(kg::k_kg_rows<T, NN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kg::rows_smem<T, NN>()));
    auto run_rows = [&](std::vector<Row> &rows) -> int {
        dfree(d_rows);
        KG(cudaMalloc(&d_rows, rows.size() * sizeof(Row)));
        KG(cudaMemcpyAsync(d_rows, rows.data(), rows.size() * sizeof(Row), cudaMemcpyHostToDevice, ctx->stream));
        kg::k_kg_rows<T, NN><<<(unsigned)rows.size(), 256, kg::rows_smem<T, NN>(), ctx->stream>>>(key, d_rows);
        KG(cudaGetLastError());
        KG(cudaStreamSynchronize(ctx->stream));                // `rows` (host) is read by the async copy
        return 0;
    };
    // unienc rows of one stream (unienc.jl:36-75): polys [j][d, f.b, f.a] into dst
    auto unienc_rows = [&](std::vector<Row> &rows, uint64_t stream, const int8_t *r_vec, const int8_t *msg_vec, T msg_scalar, T *dst) {
        const int l = p.l_uni;
        for (int j = 0; j < l; j++) {
            Row d{};
            d.stream = stream; d.e_off = (uint64_t)NN + (uint64_t)j * 2 * NN; d.a_given = (const T *)ctx->crs_coeff + (size_t)j * NN;
            d.s = r_vec; d.sigma = p.beta; d.gen_a = 0; d.negate = 0;
            d.out0 = dst + (size_t)(j * 3 + 0) * NN; d.out1 = nullptr;
            if (msg_vec) { d.msg_mode = 3; d.msg_vec = msg_vec; d.msg = gvec(j, p.logB_uni); }
            else { d.msg_mode = 1; d.msg = (T)(gvec(j, p.logB_uni) * msg_scalar); }
            rows.push_back(d);
        }
        for (int j = 0; j < l; j++) {
            Row f{};
            f.stream = stream; f.a_off = (uint64_t)NN + (uint64_t)l * 2 * NN + (uint64_t)j * ((uint64_t)NN * WT + 2 * NN);
            f.e_off = f.a_off + (uint64_t)NN * WT;
            f.s = d_ring; f.sigma = p.beta; f.gen_a = 1; f.negate = 1;
            f.msg_mode = 3; f.msg_vec = r_vec; f.msg = gvec(j, p.logB_uni);
            f.out0 = dst + (size_t)(j * 3 + 1) * NN; f.out1 = dst + (size_t)(j * 3 + 2) * NN;
            rows.push_back(f);
        }
    };

    // ---- public key b (unienc.jl:77-90; keygen.jl:68,100,136)
    if (ctx->mk) {
        if (!ctx->crs_coeff) { cleanup(); return fail(ctx, MKTFHE_ERR_STATE, "CRS missing: call mktfhe_keygen_common first"); }
        std::vector<Row> rows;
        KG(cudaMalloc(&tmp, (size_t)p.l_uni * NN * sizeof(T)));
        for (int j = 0; j < p.l_uni; j++) {
            Row r{};
            r.stream = kg::stream_id(kg::S_PUBB, party, 0); r.e_off = (uint64_t)j * 2 * NN;
            r.a_given = (const T *)ctx->crs_coeff + (size_t)j * NN; r.s = d_ring; r.sigma = p.beta; r.gen_a = 0; r.negate = 1;
            r.out0 = tmp + (size_t)j * NN;
            rows.push_back(r);
        }
        if ((rc = run_rows(rows))) return rc;
        dfree(ctx->pubb[party]);
        KG(cudaMalloc(&ctx->pubb[party], mktfhe_pubb_doubles(&p) * 8));
        if ((rc = kg_fft<T, H>(ctx, tmp, ctx->pubb[party], p.l_uni))) { cleanup(); return rc; }
        KG(cudaStreamSynchronize(ctx->stream));
        dfree(tmp);
    }
    // ---- rlk = UniEnc(gswkey) under the ring key (keygen.jl:103,139)
    if (ctx->kms) {
        std::vector<Row> rows;
        KG(cudaMalloc(&d_r, NN));
        kg::k_kg_ternary<<<1, 256, (size_t)NN * 4, ctx->stream>>>(key, kg::S_RLK, party, 0, NN, d_r);
        KG(cudaGetLastError());
        KG(cudaMalloc(&tmp, (size_t)3 * p.l_uni * NN * sizeof(T)));
        unienc_rows(rows, kg::stream_id(kg::S_RLK, party, 0), d_r, d_gsw, (T)0, tmp);
        if ((rc = run_rows(rows))) return rc;
        dfree(ctx->rlk[party]);
        KG(cudaMalloc(&ctx->rlk[party], mktfhe_rlk_doubles(&p) * 8));
        if ((rc = kg_fft<T, H>(ctx, tmp, ctx->rlk[party], (size_t)3 * p.l_uni))) { cleanup(); return rc; }
        KG(cudaStreamSynchronize(ctx->stream));
        dfree(tmp); dfree(d_r);
    }
    // ---- brk (keygen.jl:12-14,39-41,71-73,106-108,143-145)
    {
        const size_t polys_per = ccs ? (size_t)3 * p.l_uni : (size_t)4 * p.l_gsw;
        std::vector<Row> rows;
        KG(cudaMalloc(&tmp, (size_t)n * polys_per * NN * sizeof(T)));
        if (ccs) {
            KG(cudaMalloc(&d_r, (size_t)n * NN));
            kg::k_kg_ternary<<<n, 256, (size_t)NN * 4, ctx->stream>>>(key, kg::S_BRK, party, 0, NN, d_r);
            KG(cudaGetLastError());
            for (int i = 0; i < n; i++)
                unienc_rows(rows, kg::stream_id(kg::S_BRK, party, (uint64_t)i), d_r + (size_t)i * NN, nullptr, (T)h_lwe[i], tmp + (size_t)i * polys_per * NN);
        } else {
            const int l = p.l_gsw;
            for (int i = 0; i < n; i++)
                for (int r = 0; r < 2 * l; r++) {
                    const int basket = r / l, j = r % l;
                    Row d{};
                    d.stream = kg::stream_id(kg::S_BRK, party, (uint64_t)i);
                    d.a_off = (uint64_t)r * ((uint64_t)NN * WT + 2 * NN); d.e_off = d.a_off + (uint64_t)NN * WT;
                    d.s = brk_key; d.sigma = p.beta; d.gen_a = 1; d.negate = 1;
                    d.msg_mode = basket == 0 ? 1 : 2; d.msg = (T)(gvec(j, p.logB_gsw) * (T)h_lwe[i]);
                    d.out0 = tmp + ((size_t)i * polys_per + (size_t)r * 2 + 0) * NN; d.out1 = d.out0 + NN;
                    rows.push_back(d);
                }
        }
        if ((rc = run_rows(rows))) return rc;
        dfree(ctx->brk[party]);
        KG(cudaMalloc(&ctx->brk[party], mktfhe_brk_doubles(&p) * 8));
        if ((rc = kg_fft<T, H>(ctx, tmp, ctx->brk[party], (size_t)n * polys_per))) { cleanup(); return rc; }
        KG(cudaStreamSynchronize(ctx->stream));
        dfree(tmp); dfree(d_r);
    }
    // ---- ksk (keygen.jl:16-24,43-52,75-79,110-114,147-151), stored with the 16-byte aligned row stride of uploaded keys
    {
        const int Dk = mktfhe_ksk_rows(&p), nrows = Dk * p.f, rowp = (n + 1 + 3) / 4 * 4;
        const size_t words = (size_t)nrows * n + 4 * ((nrows + 1) / 2);
        dfree(ctx->ksk[party]);
        KG(cudaMalloc(&ctx->ksk[party], (size_t)NN * nrows * rowp * 4));
        KG(cudaFuncSetAttribute(kg::k_kg_ksk, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(words * 4)));
        kg::k_kg_ksk<<<NN, 256, words * 4, ctx->stream>>>(key, party, n, NN, Dk, p.f, p.logD, ctx->block ? 1 : 0, p.alpha, d_lwe, d_ring,
                                                       ctx->ksk[party], rowp);
        KG(cudaGetLastError());
        KG(cudaStreamSynchronize(ctx->stream));
    }
#undef KG
    cleanup();
    ctx->finalized = false;
    return 0;
}

// 16 independent FMA chains per thread, 16 warps per SM: measured 36.9-37.0 TFLOP/s on B200 (tools/dfma_ilp.cu shows
// the dependent-issue latency is ~9 cycles and that 2 warps per scheduler with 4 chains each already reach 93 %).
__global__ void k_dfma_peak(double *out, int iters) {
    double a[16];
#pragma unroll
    for (int i = 0; i < 16; i++) a[i] = threadIdx.x * 1e-9 + i;
    const double m = 1.0000001, c = 1e-7;
    for (int it = 0; it < iters; it++)
#pragma unroll
        for (int i = 0; i < 16; i++) a[i] = fma(a[i], m, c);
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < 16; i++) s += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <class T> int fft_hook(mktfhe_ctx *ctx, bool inverse, const void *in, void *out, size_t batch) {
    const int H = ctx->H, N = ctx->N;
    DevTmp<T> d_poly; DevTmp<cplx> d_spec;
    CK(d_poly.alloc(batch * N * sizeof(T)));
    CK(d_spec.alloc(batch * H * sizeof(cplx)));
    const int G = MK_THREADS / (H / 8);
    const size_t smem = (size_t)G * padded_len(H) * sizeof(cplx);
    const unsigned grid = (unsigned)((batch + G - 1) / G);
    if (!inverse) {
        CK(cudaMemcpyAsync(d_poly, in, batch * N * sizeof(T), cudaMemcpyHostToDevice, ctx->stream));
        if (H == 1024) k_fft_batch<T, 1024><<<grid, MK_THREADS, smem, ctx->stream>>>(d_poly, d_spec, ctx->tables(), (int)batch);
        else k_fft_batch<T, 512><<<grid, MK_THREADS, smem, ctx->stream>>>(d_poly, d_spec, ctx->tables(), (int)batch);
        CK(cudaMemcpyAsync(out, d_spec, batch * H * sizeof(cplx), cudaMemcpyDeviceToHost, ctx->stream));
    } else {
        CK(cudaMemcpyAsync(d_spec, in, batch * H * sizeof(cplx), cudaMemcpyHostToDevice, ctx->stream));
        if (H == 1024) k_ifft_batch<T, 1024><<<grid, MK_THREADS, smem, ctx->stream>>>(d_spec, d_poly, ctx->tables(), (int)batch);
        else k_ifft_batch<T, 512><<<grid, MK_THREADS, smem, ctx->stream>>>(d_spec, d_poly, ctx->tables(), (int)batch);
        CK(cudaMemcpyAsync(out, d_poly, batch * N * sizeof(T), cudaMemcpyDeviceToHost, ctx->stream));
    }
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(ctx->stream));
    return 0;
}

template <class T> int decomp_hook(mktfhe_ctx *ctx, int l, int logB, const void *polys, void *digits, size_t batch) {
    const int N = ctx->N;
    const size_t total = batch * N;
    DevTmp<T> d_in, d_out;
    CK(d_in.alloc(total * sizeof(T)));
    CK(d_out.alloc(total * l * sizeof(T)));
    CK(cudaMemcpyAsync(d_in, polys, total * sizeof(T), cudaMemcpyHostToDevice, ctx->stream));
    k_decomp_batch<T><<<(unsigned)((total + 255) / 256), 256, 0, ctx->stream>>>(d_in, d_out, N, l, logB, total);
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(digits, d_out, total * l * sizeof(T), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return 0;
}

// ---- multi-device front --------------------------------------------------------------------------------
// Gates are independent: a front context owns one child per device, keys are uploaded once to the first device
// and replicated device-to-device, a host batch is cut into contiguous slices and every child runs its slice on its own
// stream from its own host thread.  No collective on the hot path.
bool is_multi(const mktfhe_ctx *c) { return c && !c->children.empty(); }

void shard(size_t batch, size_t ndev, size_t r, size_t *lo, size_t *hi) {
    const size_t base = batch / ndev, rem = batch % ndev;
    *lo = r * base + (r < rem ? r : rem);
    *hi = *lo + base + (r < rem ? 1 : 0);
}

template <class F> int on_children(mktfhe_ctx *front, F fn) {
    const size_t n = front->children.size();
    std::vector<int> rcs(n, 0);
    std::vector<std::thread> th;
    for (size_t r = 1; r < n; r++) th.emplace_back([&, r]() { rcs[r] = fn(front->children[r], r); });
    rcs[0] = fn(front->children[0], 0);
    for (auto &t : th) t.join();
    for (size_t r = 0; r < n; r++)
        if (rcs[r]) { front->err = "device " + std::to_string(front->children[r]->device) + ": " + front->children[r]->err; return rcs[r]; }
    return 0;
}

template <class T> int clone_buf(mktfhe_ctx *ctx, T *&dst, const T *src, int src_dev, size_t bytes) {
    dfree(dst);
    if (!src) return 0;
    CK(cudaMalloc(&dst, bytes));
    CK(cudaMemcpyPeer(dst, ctx->device, src, src_dev, bytes));
    return 0;
}

// replicate the uploaded (not yet finalized) key material of `src` into `ctx` over NVLink / PCIe peer copies
int clone_keys(mktfhe_ctx *ctx, const mktfhe_ctx *src) {
    CK(cudaSetDevice(ctx->device));
    const mktfhe_params &p = ctx->p;
    int rc;
    const size_t rows = (size_t)ctx->N * mktfhe_ksk_rows(&p) * p.f, rowp = ((size_t)p.n + 1 + 3) / 4 * 4;
    for (int i = 0; i < ctx->nparties; i++) {
        if ((rc = clone_buf(ctx, ctx->brk[i], src->brk[i], src->device, mktfhe_brk_doubles(&p) * 8))) return rc;
        if ((rc = clone_buf(ctx, ctx->ksk[i], src->ksk[i], src->device, rows * rowp * 4))) return rc;
        if ((rc = clone_buf(ctx, ctx->rlk[i], src->rlk[i], src->device, mktfhe_rlk_doubles(&p) * 8))) return rc;
        if ((rc = clone_buf(ctx, ctx->pubb[i], src->pubb[i], src->device, mktfhe_pubb_doubles(&p) * 8))) return rc;
    }
    if ((rc = clone_buf(ctx, ctx->crs, src->crs, src->device, mktfhe_crs_doubles(&p) * 8))) return rc;
    ctx->finalized = false;
    return 0;
}

#define SINGLE_ONLY(ctx, what)                                                                                         \
    do {                                                                                                               \
        if (is_multi(ctx)) return fail(ctx, MKTFHE_ERR_ARG, what " is a single-device entry point: use one of the per-device contexts"); \
    } while (0)

}  // namespace

extern "C" {

const char *mktfhe_last_error(const mktfhe_ctx *ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }

int mktfhe_ctx_create(const mktfhe_params *params, int device, mktfhe_ctx **out) {
    mktfhe_ctx *ctx = nullptr;   // for CK/fail before allocation
    if (!params || !out) return fail(nullptr, MKTFHE_ERR_ARG, "null argument");
    *out = nullptr;
    const mktfhe_params &p = *params;
    const bool kms = p.scheme == MKTFHE_KMS || p.scheme == MKTFHE_KMS_BLOCK;
    const bool blk = p.scheme == MKTFHE_LMSS || p.scheme == MKTFHE_KMS_BLOCK;
    const bool mk = kms || p.scheme == MKTFHE_CCS;
    if (p.scheme < 0 || p.scheme > MKTFHE_KMS_BLOCK) return fail(nullptr, MKTFHE_ERR_PARAMS, "unknown scheme");
    if (kms ? p.N != 2048 : p.N != 1024)
        return fail(nullptr, MKTFHE_ERR_PARAMS, "ring dimension: KMS* needs N = 2048, CGGI/LMSS/CCS need N = 1024 (params.jl)");
    if (!mk && p.k != 1) return fail(nullptr, MKTFHE_ERR_PARAMS, "single-key schemes support RLWE length k = 1 only");
    if (blk && (p.ell != 3 || p.d * p.ell != p.n)) return fail(nullptr, MKTFHE_ERR_PARAMS, "block schemes need ell = 3 and n = d*ell");
    if (p.n + 1 > 3 * MK_THREADS || p.n < 1 || p.k < 1) return fail(nullptr, MKTFHE_ERR_PARAMS, "n out of range (1..767)");
    if (p.f * p.logD > 32 || p.f > 16 || p.logD < 1 || p.logD > 7) return fail(nullptr, MKTFHE_ERR_PARAMS, "key-switch gadget out of range");
    const int lmax = p.l_gsw > p.l_uni ? (p.l_gsw > p.l_lev ? p.l_gsw : p.l_lev) : (p.l_uni > p.l_lev ? p.l_uni : p.l_lev);
    if (lmax > MK_MAXL) return fail(nullptr, MKTFHE_ERR_PARAMS, "gadget length above 16");
    const int w = kms ? 64 : 32;
    if (p.l_gsw * p.logB_gsw > w || p.l_lev * p.logB_lev > w || p.l_uni * p.logB_uni > w)
        return fail(nullptr, MKTFHE_ERR_PARAMS, "gadget l*logB exceeds the torus width");
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return fail(nullptr, MKTFHE_ERR_CUDA, std::string("no CUDA device (there is no CPU fallback): ") + cudaGetErrorString(e));
    if (device < 0 || device >= ndev) return fail(nullptr, MKTFHE_ERR_ARG, "bad device index");
    e = cudaSetDevice(device);
    if (e != cudaSuccess) return fail(nullptr, MKTFHE_ERR_CUDA, cudaGetErrorString(e));

    ctx = new mktfhe_ctx;
    ctx->p = p; ctx->device = device;
    ctx->N = p.N; ctx->H = p.N / 2; ctx->bits = w;
    ctx->kms = kms; ctx->block = blk; ctx->mk = mk;
    ctx->nparties = mk ? p.k : 1;
    ctx->R = kms ? 1 + (p.k - 1) * p.l_lev : 1;
    ctx->brk.assign(ctx->nparties, nullptr); ctx->rlk.assign(ctx->nparties, nullptr);
    ctx->pubb.assign(ctx->nparties, nullptr); ctx->ksk.assign(ctx->nparties, nullptr);
    e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking);
    if (e != cudaSuccess) { delete ctx; return fail(nullptr, MKTFHE_ERR_CUDA, cudaGetErrorString(e)); }
    cudaDeviceGetAttribute(&ctx->sms, cudaDevAttrMultiProcessorCount, device);
    // Workspace budget: batches whose per-gate scratch exceeds it run in chunks.  Default: 24 GiB or 40 % of the device memory,
    // whichever is smaller; MKTFHE_WORKSPACE_MB overrides (tests use it to force chunking).
    {
        size_t free_b = 0, total_b = 0;
        if (cudaMemGetInfo(&free_b, &total_b) == cudaSuccess && total_b) ctx->mem_budget = std::min(ctx->mem_budget, total_b / 5 * 2);
        if (const char *e = getenv("MKTFHE_WORKSPACE_MB")) {
            const long mb = atol(e);
            if (mb > 0) ctx->mem_budget = (size_t)mb << 20;
        }
    }
    *out = ctx;
    return 0;
}

int mktfhe_ctx_create_multi(const mktfhe_params *params, int ndev, const int *devices, mktfhe_ctx **out) {
    if (!params || !out) return fail(nullptr, MKTFHE_ERR_ARG, "null argument");
    *out = nullptr;
    int avail = 0;
    cudaError_t e = cudaGetDeviceCount(&avail);
    if (e != cudaSuccess || avail == 0)
        return fail(nullptr, MKTFHE_ERR_CUDA, std::string("no CUDA device (there is no CPU fallback): ") + cudaGetErrorString(e));
    if (ndev <= 0) ndev = avail;
    std::vector<int> devs(ndev);
    for (int i = 0; i < ndev; i++) {
        devs[i] = devices ? devices[i] : i;
        if (devs[i] < 0 || devs[i] >= avail) return fail(nullptr, MKTFHE_ERR_ARG, "bad device index " + std::to_string(devs[i]));
        for (int j = 0; j < i; j++) if (devs[j] == devs[i]) return fail(nullptr, MKTFHE_ERR_ARG, "device listed twice");
    }
    mktfhe_ctx *front = new mktfhe_ctx;
    front->p = *params; front->device = devs[0];
    for (int i = 0; i < ndev; i++) {
        mktfhe_ctx *child = nullptr;
        const int rc = mktfhe_ctx_create(params, devs[i], &child);
        if (rc) { for (auto *c : front->children) mktfhe_ctx_destroy(c); delete front; return rc; }
        front->children.push_back(child);
    }
    // peer access from every device to the first one (key replication); failure only means staged copies
    for (int i = 1; i < ndev; i++) {
        int can = 0;
        if (cudaDeviceCanAccessPeer(&can, devs[i], devs[0]) == cudaSuccess && can) {
            cudaSetDevice(devs[i]);
            cudaError_t pe = cudaDeviceEnablePeerAccess(devs[0], 0);
            if (pe != cudaSuccess) cudaGetLastError();            // already enabled is fine
        }
    }
    const mktfhe_ctx *c0 = front->children[0];
    front->N = c0->N; front->H = c0->H; front->bits = c0->bits; front->R = c0->R; front->nparties = c0->nparties;
    front->mk = c0->mk; front->kms = c0->kms; front->block = c0->block;
    *out = front;
    return 0;
}

int mktfhe_ctx_devices(const mktfhe_ctx *ctx, int *devices_out, int cap) {
    if (!ctx) return MKTFHE_ERR_ARG;
    if (!is_multi(ctx)) { if (devices_out && cap > 0) devices_out[0] = ctx->device; return 1; }
    const int n = (int)ctx->children.size();
    for (int i = 0; i < n && i < cap && devices_out; i++) devices_out[i] = ctx->children[i]->device;
    return n;
}

void mktfhe_ctx_destroy(mktfhe_ctx *ctx) {
    if (!ctx) return;
    if (is_multi(ctx)) {
        for (auto *c : ctx->children) mktfhe_ctx_destroy(c);
        delete ctx;
        return;
    }
    cudaSetDevice(ctx->device);
    cudaDeviceSynchronize();
    free_workspace(ctx);
    free_wires(ctx);
    fast_free(ctx->fast);
    fast32_free(ctx->fast32);
    fast32w_free(ctx->fast32w);
    fastccs_free(ctx->fastccs);
    for (auto &q : ctx->brk) dfree(q);
    for (auto &q : ctx->rlk) dfree(q);
    for (auto &q : ctx->pubb) dfree(q);
    for (auto &q : ctx->ksk) dfree(q);
    dfree(ctx->crs_coeff);
    dfree(ctx->crs); dfree(ctx->psi); dfree(ctx->psiinv); dfree(ctx->roots); dfree(ctx->rootsinv); dfree(ctx->mono);
    dfree(ctx->d_brk); dfree(ctx->d_rlk); dfree(ctx->d_pubb); dfree(ctx->d_ksk);
    for (auto &s : ctx->events) for (auto &ev : s.e) cudaEventDestroy(ev);
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
}

int mktfhe_set_mode(mktfhe_ctx *ctx, int mode) {
    if (!ctx) return MKTFHE_ERR_ARG;
    if (mode != MKTFHE_MODE_STRICT && mode != MKTFHE_MODE_FAST) return fail(ctx, MKTFHE_ERR_ARG, "bad mode");
    for (auto *c : ctx->children) c->mode = mode;
    ctx->mode = mode;
    return 0;
}
int mktfhe_get_mode(const mktfhe_ctx *ctx) { return ctx ? (is_multi(ctx) ? ctx->children[0]->mode : ctx->mode) : MKTFHE_ERR_ARG; }

int mktfhe_upload_party_key(mktfhe_ctx *ctx, int party, const double *brk, const double *rlk, const double *pubb, const uint32_t *ksk) {
    if (!ctx) return MKTFHE_ERR_ARG;
    if (is_multi(ctx)) {            // one host -> device copy; mktfhe_finalize_keys replicates device to device
        const int rc = mktfhe_upload_party_key(ctx->children[0], party, brk, rlk, pubb, ksk);
        if (rc) ctx->err = ctx->children[0]->err;
        return rc;
    }
    if (party < 0 || party >= ctx->nparties) return fail(ctx, MKTFHE_ERR_ARG, "bad party index");
    if (!brk || !ksk) return fail(ctx, MKTFHE_ERR_ARG, "brk and ksk are required");
    if (ctx->kms && (!rlk || !pubb)) return fail(ctx, MKTFHE_ERR_ARG, "KMS needs rlk and pubb");
    if (ctx->p.scheme == MKTFHE_CCS && !pubb) return fail(ctx, MKTFHE_ERR_ARG, "CCS needs pubb");
    CK(cudaSetDevice(ctx->device));
    int rc;
    if ((rc = upload(ctx, ctx->brk[party], brk, mktfhe_brk_doubles(&ctx->p) * 8))) return rc;
    {   // ksk rows are stored with a 16-byte aligned stride (bulk copies in the tiled key switch)
        const size_t rows = (size_t)ctx->N * mktfhe_ksk_rows(&ctx->p) * ctx->p.f, row = (size_t)ctx->p.n + 1, rowp = (row + 3) / 4 * 4;
        dfree(ctx->ksk[party]);
        CK(cudaMalloc(&ctx->ksk[party], rows * rowp * 4));
        CK(cudaMemset(ctx->ksk[party], 0, rows * rowp * 4));
        CK(cudaMemcpy2D(ctx->ksk[party], rowp * 4, ksk, row * 4, row * 4, rows, cudaMemcpyDefault));
    }
    if (ctx->kms && (rc = upload(ctx, ctx->rlk[party], rlk, mktfhe_rlk_doubles(&ctx->p) * 8))) return rc;
    if (ctx->mk && (rc = upload(ctx, ctx->pubb[party], pubb, mktfhe_pubb_doubles(&ctx->p) * 8))) return rc;
    ctx->finalized = false;
    return 0;
}

int mktfhe_upload_common(mktfhe_ctx *ctx, const double *crs_fft) {
    if (!ctx) return MKTFHE_ERR_ARG;
    if (is_multi(ctx)) {
        const int rc = mktfhe_upload_common(ctx->children[0], crs_fft);
        if (rc) ctx->err = ctx->children[0]->err;
        return rc;
    }
    if (!ctx->mk) return 0;
    if (!crs_fft) return fail(ctx, MKTFHE_ERR_ARG, "crs_fft is required for CCS / KMS");
    CK(cudaSetDevice(ctx->device));
    ctx->finalized = false;
    return upload(ctx, ctx->crs, crs_fft, mktfhe_crs_doubles(&ctx->p) * 8);
}

int mktfhe_keygen_common(mktfhe_ctx *ctx, uint64_t seed, const uint8_t *key32) {
    if (!ctx) return MKTFHE_ERR_ARG;
    if (is_multi(ctx)) {
        const int rc = mktfhe_keygen_common(ctx->children[0], seed, key32);
        if (rc) ctx->err = ctx->children[0]->err;
        return rc;
    }
    if (!ctx->mk) return 0;
    CK(cudaSetDevice(ctx->device));
    int rc;
    if ((rc = ensure_tables(ctx))) return rc;
    const kg::Key key = kg_key(seed, key32);
    rc = ctx->bits == 64 ? kg_common<uint64_t, 2048>(ctx, key) : kg_common<uint32_t, 1024>(ctx, key);
    if (rc) return rc;
    CK(cudaStreamSynchronize(ctx->stream));
    ctx->finalized = false;
    return 0;
}

int mktfhe_keygen_party(mktfhe_ctx *ctx, int party, uint64_t seed, const uint8_t *key32) {
    if (!ctx) return MKTFHE_ERR_ARG;
    if (is_multi(ctx)) {            // generated on the first device; mktfhe_finalize_keys replicates device to device
        const int rc = mktfhe_keygen_party(ctx->children[0], party, seed, key32);
        if (rc) ctx->err = ctx->children[0]->err;
        return rc;
    }
    if (party < 0 || party >= ctx->nparties) return fail(ctx, MKTFHE_ERR_ARG, "bad party index");
    CK(cudaSetDevice(ctx->device));
    int rc;
    if ((rc = ensure_tables(ctx))) return rc;
    const kg::Key key = kg_key(seed, key32);
    return ctx->bits == 64 ? kg_party<uint64_t, 2048>(ctx, party, key) : kg_party<uint32_t, 1024>(ctx, party, key);
}

int mktfhe_download_party_key(mktfhe_ctx *ctx, int party, double *brk, double *rlk, double *pubb, uint32_t *ksk, double *crs_fft) {
    if (!ctx) return MKTFHE_ERR_ARG;
    if (is_multi(ctx)) return mktfhe_download_party_key(ctx->children[0], party, brk, rlk, pubb, ksk, crs_fft);
    if (party < 0 || party >= ctx->nparties) return fail(ctx, MKTFHE_ERR_ARG, "bad party index");
    CK(cudaSetDevice(ctx->device));
    const mktfhe_params &p = ctx->p;
    if (brk) { if (!ctx->brk[party]) return fail(ctx, MKTFHE_ERR_STATE, "no brk"); CK(cudaMemcpy(brk, ctx->brk[party], mktfhe_brk_doubles(&p) * 8, cudaMemcpyDeviceToHost)); }
    if (rlk) { if (!ctx->rlk[party]) return fail(ctx, MKTFHE_ERR_STATE, "no rlk"); CK(cudaMemcpy(rlk, ctx->rlk[party], mktfhe_rlk_doubles(&p) * 8, cudaMemcpyDeviceToHost)); }
    if (pubb) { if (!ctx->pubb[party]) return fail(ctx, MKTFHE_ERR_STATE, "no pubb"); CK(cudaMemcpy(pubb, ctx->pubb[party], mktfhe_pubb_doubles(&p) * 8, cudaMemcpyDeviceToHost)); }
    if (crs_fft) { if (!ctx->crs) return fail(ctx, MKTFHE_ERR_STATE, "no crs"); CK(cudaMemcpy(crs_fft, ctx->crs, mktfhe_crs_doubles(&p) * 8, cudaMemcpyDeviceToHost)); }
    if (ksk) {
        if (!ctx->ksk[party]) return fail(ctx, MKTFHE_ERR_STATE, "no ksk");
        const size_t rows = (size_t)ctx->N * mktfhe_ksk_rows(&p) * p.f, row = (size_t)p.n + 1, rowp = (row + 3) / 4 * 4;
        CK(cudaMemcpy2D(ksk, row * 4, ctx->ksk[party], rowp * 4, row * 4, rows, cudaMemcpyDeviceToHost));
    }
    return 0;
}

int mktfhe_finalize_keys(mktfhe_ctx *ctx) {
    if (!ctx) return MKTFHE_ERR_ARG;
    if (is_multi(ctx)) {
        // replicate (peer copies from the first device, one thread per target), then build tables and FAST layouts everywhere
        const int rc = on_children(ctx, [&](mktfhe_ctx *c, size_t r) -> int {
            if (r > 0) { const int rc2 = clone_keys(c, ctx->children[0]); if (rc2) return rc2; }
            return 0;
        });
        if (rc) return rc;
        const int rc3 = on_children(ctx, [&](mktfhe_ctx *c, size_t) -> int { return mktfhe_finalize_keys(c); });
        if (rc3) return rc3;
        ctx->mode = ctx->children[0]->mode; ctx->finalized = true;
        return 0;
    }
    CK(cudaSetDevice(ctx->device));
    for (int i = 0; i < ctx->nparties; i++)
        if (!ctx->brk[i] || !ctx->ksk[i]) return fail(ctx, MKTFHE_ERR_STATE, "party key missing: " + std::to_string(i));
    if (ctx->mk && !ctx->crs) return fail(ctx, MKTFHE_ERR_STATE, "common reference string missing");
    int rc;
    if ((rc = ensure_tables(ctx))) return rc;
    if ((rc = upload(ctx, ctx->d_brk, ctx->brk.data(), sizeof(void *) * ctx->nparties))) return rc;
    if ((rc = upload(ctx, ctx->d_ksk, ctx->ksk.data(), sizeof(void *) * ctx->nparties))) return rc;
    if ((rc = upload(ctx, ctx->d_rlk, ctx->rlk.data(), sizeof(void *) * ctx->nparties))) return rc;
    if ((rc = upload(ctx, ctx->d_pubb, ctx->pubb.data(), sizeof(void *) * ctx->nparties))) return rc;
    // monomial table, built with the same transform the reference uses (scheme.jl:121-146)
    dfree(ctx->mono);
    CK(cudaMalloc(&ctx->mono, sizeof(cplx) * (size_t)2 * ctx->N * ctx->H));
    if (ctx->H == 1024) {
        const int G = MK_THREADS / (1024 / 8);
        const size_t smem = (size_t)G * padded_len(1024) * sizeof(cplx);
        k_build_monomials<1024><<<(2 * ctx->N + G - 1) / G, MK_THREADS, smem, ctx->stream>>>(ctx->mono, ctx->tables());
    } else {
        const int G = MK_THREADS / (512 / 8);
        const size_t smem = (size_t)G * padded_len(512) * sizeof(cplx);
        k_build_monomials<512><<<(2 * ctx->N + G - 1) / G, MK_THREADS, smem, ctx->stream>>>(ctx->mono, ctx->tables());
    }
    CK(cudaGetLastError());
    if (fast_supported(ctx->p) && (rc = fast_build(ctx->fast, ctx->p, ctx->brk, ctx->rlk, ctx->pubb, ctx->crs, ctx->stream, ctx->err))) return rc;
    if (fast32_supported(ctx->p) && (rc = fast32_build(ctx->fast32, ctx->p, ctx->brk[0], ctx->stream, ctx->err))) return rc;
    if (fast32_supported(ctx->p) && ctx->p.scheme == MKTFHE_CGGI && (rc = fast32w_build(ctx->fast32w, ctx->p, ctx->brk[0], ctx->stream, ctx->err))) return rc;
    if (fastccs_supported(ctx->p)) {
        if ((rc = fast32_build(ctx->fast32, ctx->p, nullptr, ctx->stream, ctx->err))) return rc;      // transform tables only
        if ((rc = fastccs_build(ctx->fastccs, ctx->p, ctx->brk, ctx->pubb, ctx->crs, ctx->stream, ctx->err))) return rc;
    }
    ctx->mode = MKTFHE_MODE_FAST;          // production default; floating-point stages without a FAST kernel run the STRICT one
    CK(cudaStreamSynchronize(ctx->stream));
    ctx->finalized = true;
    return 0;
}

int mktfhe_sync(mktfhe_ctx *ctx) {
    if (!ctx) return MKTFHE_ERR_ARG;
    if (is_multi(ctx)) return on_children(ctx, [](mktfhe_ctx *c, size_t) -> int { return mktfhe_sync(c); });
    CK(cudaSetDevice(ctx->device));
    CK(cudaStreamSynchronize(ctx->stream));
    return 0;
}

void *mktfhe_stream(mktfhe_ctx *ctx) { return ctx ? (void *)(is_multi(ctx) ? ctx->children[0]->stream : ctx->stream) : nullptr; }

static int run_pipeline(mktfhe_ctx *ctx, int gate_op, const uint32_t *in1, const uint32_t *in2, uint32_t *out, size_t g,
                        const int32_t *ops = nullptr, const int32_t *idx1 = nullptr, const int32_t *idx2 = nullptr) {
    int rc;
    StageEvents *ev = next_events(ctx);
    if (!ev) return fail(ctx, MKTFHE_ERR_CUDA, "cudaEventCreate failed");
    CK(cudaEventRecord(ev->e[0], ctx->stream));
    if ((rc = run_prep(ctx, gate_op, in1, in2, nullptr, ctx->w_tilde, g, ops, idx1, idx2))) return rc;
    CK(cudaEventRecord(ev->e[1], ctx->stream));
    if ((rc = run_blindrotate(ctx, ctx->w_tilde, g, ev))) return rc;
    CK(cudaEventRecord(ev->e[3], ctx->stream));
    if ((rc = run_keyswitch(ctx, ctx->w_acc, out, g))) return rc;
    CK(cudaEventRecord(ev->e[4], ctx->stream));
    return 0;
}

int mktfhe_gate_batch_dev(mktfhe_ctx *ctx, int gate_op, const uint32_t *in1, const uint32_t *in2, uint32_t *out, size_t batch) {
    int rc;
    if (!ctx) return MKTFHE_ERR_ARG;
    SINGLE_ONLY(ctx, "mktfhe_gate_batch_dev");
    if ((rc = check_ready(ctx))) return rc;
    if (!in1 || !out || (gate_op >= 0 && !in2)) return fail(ctx, MKTFHE_ERR_ARG, "null ciphertext pointer");
    if (gate_op > MKTFHE_NOR) return fail(ctx, MKTFHE_ERR_ARG, "bad gate opcode");
    if (batch == 0) return 0;
    const size_t lw = mktfhe_lwe_words(&ctx->p);
    const size_t chunk = chunk_gates(ctx, batch);
    if ((rc = ensure_workspace(ctx, chunk))) return rc;
    ctx->events_used = 0; ctx->launches = 0;
    for (size_t g0 = 0; g0 < batch; g0 += chunk) {
        const size_t g = batch - g0 < chunk ? batch - g0 : chunk;
        if ((rc = run_pipeline(ctx, gate_op, in1 + g0 * lw, in2 ? in2 + g0 * lw : nullptr, out + g0 * lw, g))) return rc;
    }
    return 0;
}

int mktfhe_gate_batch(mktfhe_ctx *ctx, int gate_op, const uint32_t *in1, const uint32_t *in2, uint32_t *out, size_t batch) {
    int rc;
    if (is_multi(ctx)) {
        if (!ctx->finalized) return fail(ctx, MKTFHE_ERR_STATE, "keys not finalized: call mktfhe_finalize_keys first");
        if (!in1 || !out || (gate_op >= 0 && !in2)) return fail(ctx, MKTFHE_ERR_ARG, "null ciphertext pointer");
        if (gate_op > MKTFHE_NOR) return fail(ctx, MKTFHE_ERR_ARG, "bad gate opcode");
        const size_t lw = mktfhe_lwe_words(&ctx->p), nd = ctx->children.size();
        // contiguous slices of the caller's arrays, one per device, each on its own host thread and stream
        return on_children(ctx, [&](mktfhe_ctx *c, size_t r) -> int {
            size_t lo, hi;
            shard(batch, nd, r, &lo, &hi);
            if (hi == lo) { c->events_used = 0; c->launches = 0; return 0; }
            return mktfhe_gate_batch(c, gate_op, in1 + lo * lw, in2 ? in2 + lo * lw : nullptr, out + lo * lw, hi - lo);
        });
    }
    if ((rc = check_ready(ctx))) return rc;
    if (!in1 || !out || (gate_op >= 0 && !in2)) return fail(ctx, MKTFHE_ERR_ARG, "null ciphertext pointer");
    if (gate_op > MKTFHE_NOR) return fail(ctx, MKTFHE_ERR_ARG, "bad gate opcode");
    if (batch == 0) return 0;
    const size_t lw = mktfhe_lwe_words(&ctx->p);
    const size_t chunk = chunk_gates(ctx, batch);
    if ((rc = ensure_workspace(ctx, chunk))) return rc;
    ctx->events_used = 0; ctx->launches = 0;
    // ciphertexts are small (<= 72 KB each); stage one chunk at a time through the workspace
    for (size_t g0 = 0; g0 < batch; g0 += chunk) {
        const size_t g = batch - g0 < chunk ? batch - g0 : chunk;
        CK(cudaMemcpyAsync(ctx->w_in1, in1 + g0 * lw, g * lw * 4, cudaMemcpyHostToDevice, ctx->stream));
        if (gate_op >= 0) CK(cudaMemcpyAsync(ctx->w_in2, in2 + g0 * lw, g * lw * 4, cudaMemcpyHostToDevice, ctx->stream));
        if ((rc = run_pipeline(ctx, gate_op, ctx->w_in1, gate_op >= 0 ? ctx->w_in2 : nullptr, ctx->w_out, g))) return rc;
        CK(cudaMemcpyAsync(out + g0 * lw, ctx->w_out, g * lw * 4, cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
    }
    return 0;
}

int mktfhe_bootstrap_batch(mktfhe_ctx *ctx, const uint32_t *in, uint32_t *out, size_t batch) {
    return mktfhe_gate_batch(ctx, -1, in, nullptr, out, batch);
}

int mktfhe_last_stage_ms(mktfhe_ctx *ctx, float *ms_out, int *launches_out) {
    if (!ctx || !ms_out) return MKTFHE_ERR_ARG;
    if (is_multi(ctx)) {            // slowest device per stage, launches summed
        for (int s = 0; s < MKTFHE_STAGE_COUNT; s++) ms_out[s] = 0.f;
        int total = 0;
        for (auto *c : ctx->children) {
            float ms[MKTFHE_STAGE_COUNT]; int l = 0;
            const int rc = mktfhe_last_stage_ms(c, ms, &l);
            if (rc) { ctx->err = c->err; return rc; }
            for (int s = 0; s < MKTFHE_STAGE_COUNT; s++) ms_out[s] = std::max(ms_out[s], ms[s]);
            total += l;
        }
        if (launches_out) *launches_out = total;
        return 0;
    }
    CK(cudaSetDevice(ctx->device));
    CK(cudaStreamSynchronize(ctx->stream));
    for (int s = 0; s < MKTFHE_STAGE_COUNT; s++) ms_out[s] = 0.f;
    for (size_t i = 0; i < ctx->events_used; i++)
        for (int s = 0; s < MKTFHE_STAGE_COUNT; s++) {
            float ms = 0.f;
            CK(cudaEventElapsedTime(&ms, ctx->events[i].e[s], ctx->events[i].e[s + 1]));
            ms_out[s] += ms;
        }
    if (launches_out) *launches_out = ctx->launches;
    return 0;
}

// ---- gate circuits over a device-resident wire table ------------------------------------------------

int mktfhe_wires_resize(mktfhe_ctx *ctx, size_t nwires) {
    int rc;
    if (!ctx) return MKTFHE_ERR_ARG;
    SINGLE_ONLY(ctx, "mktfhe_wires_resize");
    if ((rc = check_ready(ctx))) return rc;
    CK(cudaStreamSynchronize(ctx->stream));
    dfree(ctx->wires);
    ctx->nwires = 0;
    if (nwires == 0) return 0;
    if (nwires > (size_t)INT32_MAX) return fail(ctx, MKTFHE_ERR_ARG, "wire count above 2^31 - 1");
    const size_t bytes = nwires * mktfhe_lwe_words(&ctx->p) * 4;
    CK(cudaMalloc(&ctx->wires, bytes));
    CK(cudaMemsetAsync(ctx->wires, 0, bytes, ctx->stream));
    ctx->nwires = nwires;
    return 0;
}

int mktfhe_wires_write(mktfhe_ctx *ctx, size_t first, size_t count, const uint32_t *cts) {
    int rc;
    if (!ctx) return MKTFHE_ERR_ARG;
    SINGLE_ONLY(ctx, "mktfhe_wires_write");
    if ((rc = check_ready(ctx))) return rc;
    if (first + count > ctx->nwires || first + count < first) return fail(ctx, MKTFHE_ERR_ARG, "wire range outside the table");
    if (count == 0) return 0;
    if (!cts) return fail(ctx, MKTFHE_ERR_ARG, "null ciphertext pointer");
    const size_t lw = mktfhe_lwe_words(&ctx->p);
    CK(cudaMemcpyAsync(ctx->wires + first * lw, cts, count * lw * 4, cudaMemcpyDefault, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return 0;
}

int mktfhe_wires_read(mktfhe_ctx *ctx, size_t first, size_t count, uint32_t *cts) {
    int rc;
    if (!ctx) return MKTFHE_ERR_ARG;
    SINGLE_ONLY(ctx, "mktfhe_wires_read");
    if ((rc = check_ready(ctx))) return rc;
    if (first + count > ctx->nwires || first + count < first) return fail(ctx, MKTFHE_ERR_ARG, "wire range outside the table");
    if (count == 0) return 0;
    if (!cts) return fail(ctx, MKTFHE_ERR_ARG, "null ciphertext pointer");
    const size_t lw = mktfhe_lwe_words(&ctx->p);
    CK(cudaMemcpyAsync(cts, ctx->wires + first * lw, count * lw * 4, cudaMemcpyDefault, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return 0;
}

int mktfhe_gate_level(mktfhe_ctx *ctx, const int32_t *ops, const int32_t *src1, const int32_t *src2, const int32_t *dst, size_t batch) {
    int rc;
    if (!ctx) return MKTFHE_ERR_ARG;
    SINGLE_ONLY(ctx, "mktfhe_gate_level");
    if ((rc = check_ready(ctx))) return rc;
    if (batch == 0) return 0;
    if (!ops || !src1 || !src2 || !dst) return fail(ctx, MKTFHE_ERR_ARG, "null index array");
    if (!ctx->wires) return fail(ctx, MKTFHE_ERR_STATE, "no wire table: call mktfhe_wires_resize first");
    // split into bootstrapped gates and NOTs; validate every index before touching the device
    std::vector<int32_t> b_ops, b_s1, b_s2, b_dst, n_src, n_dst;
    const int64_t nw = (int64_t)ctx->nwires;
    for (size_t g = 0; g < batch; g++) {
        const int op = ops[g];
        if (op < -1 || op > MKTFHE_NOT) return fail(ctx, MKTFHE_ERR_ARG, "bad gate opcode at gate " + std::to_string(g));
        const bool two = op >= 0 && op <= MKTFHE_NOR;
        if (src1[g] < 0 || src1[g] >= nw || dst[g] < 0 || dst[g] >= nw || (two && (src2[g] < 0 || src2[g] >= nw)))
            return fail(ctx, MKTFHE_ERR_ARG, "wire index outside the table at gate " + std::to_string(g));
        if (op == MKTFHE_NOT) { n_src.push_back(src1[g]); n_dst.push_back(dst[g]); }
        else { b_ops.push_back(op); b_s1.push_back(src1[g]); b_s2.push_back(two ? src2[g] : 0); b_dst.push_back(dst[g]); }
    }
    const size_t lw = mktfhe_lwe_words(&ctx->p), nb = b_ops.size(), nn = n_src.size();
    const size_t need = nb > nn ? nb : nn;
    if (need > ctx->idx_cap) {
        dfree(ctx->w_idx);
        ctx->idx_cap = 0;
        CK(cudaMalloc(&ctx->w_idx, 4 * need * sizeof(int32_t)));
        ctx->idx_cap = need;
    }
    int32_t *d_ops = ctx->w_idx, *d_s1 = d_ops + ctx->idx_cap, *d_s2 = d_s1 + ctx->idx_cap, *d_dst = d_s2 + ctx->idx_cap;
    ctx->events_used = 0; ctx->launches = 0;
    if (nb) {
        const size_t chunk = chunk_gates(ctx, nb);
        if ((rc = ensure_workspace(ctx, chunk))) return rc;
        CK(cudaMemcpyAsync(d_ops, b_ops.data(), nb * 4, cudaMemcpyHostToDevice, ctx->stream));
        CK(cudaMemcpyAsync(d_s1, b_s1.data(), nb * 4, cudaMemcpyHostToDevice, ctx->stream));
        CK(cudaMemcpyAsync(d_s2, b_s2.data(), nb * 4, cudaMemcpyHostToDevice, ctx->stream));
        CK(cudaMemcpyAsync(d_dst, b_dst.data(), nb * 4, cudaMemcpyHostToDevice, ctx->stream));
        for (size_t g0 = 0; g0 < nb; g0 += chunk) {
            const size_t g = nb - g0 < chunk ? nb - g0 : chunk;
            if ((rc = run_pipeline(ctx, 0, ctx->wires, ctx->wires, ctx->w_out, g, d_ops + g0, d_s1 + g0, d_s2 + g0))) return rc;
            const size_t total = g * lw;
            k_wire_scatter<<<(unsigned)((total + 255) / 256), 256, 0, ctx->stream>>>(ctx->w_out, ctx->wires, nullptr, d_dst + g0, (int)lw, total, 0);
            ctx->launches++;
            CK(cudaGetLastError());
        }
        CK(cudaStreamSynchronize(ctx->stream));            // the host index vectors die with this call
    }
    if (nn) {
        CK(cudaMemcpyAsync(d_s1, n_src.data(), nn * 4, cudaMemcpyHostToDevice, ctx->stream));
        CK(cudaMemcpyAsync(d_dst, n_dst.data(), nn * 4, cudaMemcpyHostToDevice, ctx->stream));
        const size_t total = nn * lw;
        k_wire_scatter<<<(unsigned)((total + 255) / 256), 256, 0, ctx->stream>>>(nullptr, ctx->wires, d_s1, d_dst, (int)lw, total, 1);
        ctx->launches++;
        CK(cudaGetLastError());
        CK(cudaStreamSynchronize(ctx->stream));
    }
    return 0;
}

// ---- parity / debug hooks -------------------------------------------------------------------------

int mktfhe_gate_linear_batch(mktfhe_ctx *ctx, int gate_op, const uint32_t *in1, const uint32_t *in2, uint32_t *out, size_t batch) {
    int rc;
    if (!ctx) return MKTFHE_ERR_ARG;
    SINGLE_ONLY(ctx, "mktfhe_gate_linear_batch");
    if ((rc = check_ready(ctx))) return rc;
    if (!in1 || !in2 || !out || gate_op < 0 || gate_op > MKTFHE_NOR) return fail(ctx, MKTFHE_ERR_ARG, "bad argument");
    const size_t lw = mktfhe_lwe_words(&ctx->p);
    if ((rc = ensure_workspace(ctx, batch))) return rc;
    CK(cudaMemcpyAsync(ctx->w_in1, in1, batch * lw * 4, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(ctx->w_in2, in2, batch * lw * 4, cudaMemcpyHostToDevice, ctx->stream));
    if ((rc = run_prep(ctx, gate_op, ctx->w_in1, ctx->w_in2, ctx->w_lin, nullptr, batch))) return rc;
    CK(cudaMemcpyAsync(out, ctx->w_lin, batch * lw * 4, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return 0;
}

int mktfhe_modswitch_batch(mktfhe_ctx *ctx, const uint32_t *lwe, uint32_t *tilde, size_t batch) {
    int rc;
    if (!ctx) return MKTFHE_ERR_ARG;
    SINGLE_ONLY(ctx, "mktfhe_modswitch_batch");
    if ((rc = check_ready(ctx))) return rc;
    if (!lwe || !tilde) return fail(ctx, MKTFHE_ERR_ARG, "null pointer");
    const size_t lw = mktfhe_lwe_words(&ctx->p);
    if ((rc = ensure_workspace(ctx, batch))) return rc;
    CK(cudaMemcpyAsync(ctx->w_in1, lwe, batch * lw * 4, cudaMemcpyHostToDevice, ctx->stream));
    if ((rc = run_prep(ctx, -1, ctx->w_in1, nullptr, nullptr, ctx->w_tilde, batch))) return rc;
    CK(cudaMemcpyAsync(tilde, ctx->w_tilde, batch * lw * 4, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return 0;
}

int mktfhe_blindrotate_batch(mktfhe_ctx *ctx, const uint32_t *lwe, void *acc_out, size_t batch) {
    int rc;
    if (!ctx) return MKTFHE_ERR_ARG;
    SINGLE_ONLY(ctx, "mktfhe_blindrotate_batch");
    if ((rc = check_ready(ctx))) return rc;
    if (!lwe || !acc_out) return fail(ctx, MKTFHE_ERR_ARG, "null pointer");
    const size_t lw = mktfhe_lwe_words(&ctx->p);
    if ((rc = ensure_workspace(ctx, batch))) return rc;
    CK(cudaMemcpyAsync(ctx->w_in1, lwe, batch * lw * 4, cudaMemcpyHostToDevice, ctx->stream));
    if ((rc = run_prep(ctx, -1, ctx->w_in1, nullptr, nullptr, ctx->w_tilde, batch))) return rc;
    if ((rc = run_blindrotate(ctx, ctx->w_tilde, batch, nullptr))) return rc;
    CK(cudaMemcpyAsync(acc_out, ctx->w_acc, batch * (size_t)(ctx->p.k + 1) * ctx->N * (ctx->bits / 8), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return 0;
}

int mktfhe_phase1_batch(mktfhe_ctx *ctx, const uint32_t *lwe, double *levkeys_out, size_t batch) {
    int rc;
    if (!ctx) return MKTFHE_ERR_ARG;
    SINGLE_ONLY(ctx, "mktfhe_phase1_batch");
    if ((rc = check_ready(ctx))) return rc;
    if (!ctx->kms) return fail(ctx, MKTFHE_ERR_PARAMS, "phase 1 exists for KMS / KMS_BLOCK only");
    if (!lwe || !levkeys_out) return fail(ctx, MKTFHE_ERR_ARG, "null pointer");
    const size_t lw = mktfhe_lwe_words(&ctx->p);
    if ((rc = ensure_workspace(ctx, batch))) return rc;
    CK(cudaMemcpyAsync(ctx->w_in1, lwe, batch * lw * 4, cudaMemcpyHostToDevice, ctx->stream));
    if ((rc = run_prep(ctx, -1, ctx->w_in1, nullptr, nullptr, ctx->w_tilde, batch))) return rc;
    if ((rc = run_phase1(ctx, ctx->w_tilde, ctx->w_lev, batch))) return rc;
    CK(cudaMemcpyAsync(levkeys_out, ctx->w_lev, batch * (size_t)ctx->R * 2 * ctx->H * sizeof(cplx), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return 0;
}

int mktfhe_keyswitch_batch(mktfhe_ctx *ctx, const void *acc, uint32_t *lwe_out, size_t batch) {
    int rc;
    if (!ctx) return MKTFHE_ERR_ARG;
    SINGLE_ONLY(ctx, "mktfhe_keyswitch_batch");
    if ((rc = check_ready(ctx))) return rc;
    if (!acc || !lwe_out) return fail(ctx, MKTFHE_ERR_ARG, "null pointer");
    const size_t lw = mktfhe_lwe_words(&ctx->p);
    if ((rc = ensure_workspace(ctx, batch))) return rc;
    CK(cudaMemcpyAsync(ctx->w_acc, acc, batch * (size_t)(ctx->p.k + 1) * ctx->N * (ctx->bits / 8), cudaMemcpyHostToDevice, ctx->stream));
    if ((rc = run_keyswitch(ctx, ctx->w_acc, ctx->w_out, batch))) return rc;
    CK(cudaMemcpyAsync(lwe_out, ctx->w_out, batch * lw * 4, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return 0;
}

static int step_impl(mktfhe_ctx *ctx, int party, int idx, const uint32_t *atilde, void *acc_rows, size_t batch, bool block_step) {
    int rc;
    if (!ctx) return MKTFHE_ERR_ARG;
    SINGLE_ONLY(ctx, "mktfhe_cmux_step_batch / mktfhe_block_step_batch");
    if ((rc = check_ready(ctx))) return rc;
    if (ctx->p.scheme == MKTFHE_CCS) return fail(ctx, MKTFHE_ERR_PARAMS, "CCS has no RGSW step");
    if (block_step && !ctx->block) return fail(ctx, MKTFHE_ERR_PARAMS, "block step needs LMSS / KMS_BLOCK");
    const int lim = block_step ? ctx->p.d : ctx->p.n, per = block_step ? ctx->p.ell : 1;
    if (party < 0 || party >= ctx->nparties || idx < 0 || idx >= lim || !atilde || !acc_rows) return fail(ctx, MKTFHE_ERR_ARG, "bad argument");
    const size_t row_bytes = (size_t)2 * ctx->N * (ctx->bits / 8);
    DevTmp<uint32_t> d_at; DevTmp<unsigned char> d_rows;
    CK(d_at.alloc(batch * 4 * per));
    CK(d_rows.alloc(batch * row_bytes));
    CK(cudaMemcpyAsync(d_at, atilde, batch * 4 * per, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(d_rows, acc_rows, batch * row_bytes, cudaMemcpyHostToDevice, ctx->stream));
    if (ctx->mode == MKTFHE_MODE_FAST && fast_supported(ctx->p) && (block_step == ctx->block)) {
        rc = fast_cmux_step(ctx->fast, ctx->p, party, idx, d_at, d_rows.p, batch, ctx->stream, &ctx->launches, ctx->err);
    } else if (ctx->mode == MKTFHE_MODE_FAST && fast32_supported(ctx->p) && (block_step == ctx->block)) {
        fast32::Args fa{};
        fa.tilde = d_at; fa.acc_io = (uint32_t *)d_rows.p; fa.step_mode = block_step ? 2 : 1; fa.step_idx = idx; fa.units = batch;
        rc = launch_fast32(ctx, fa);
    } else {
        RgswArgs a{};
        a.tilde = d_at; a.acc_io = d_rows.p; a.mode = RG_MODE_STEP; a.step_party = party; a.step_idx = idx; a.step_block = block_step;
        rc = run_rgsw(ctx, a, batch);
    }
    if (!rc) {
        CK(cudaMemcpyAsync(acc_rows, d_rows.p, batch * row_bytes, cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
    }
    return rc;
}

int mktfhe_cmux_step_batch(mktfhe_ctx *ctx, int party, int idx, const uint32_t *atilde, void *acc_rows, size_t batch) {
    return step_impl(ctx, party, idx, atilde, acc_rows, batch, false);
}

int mktfhe_block_step_batch(mktfhe_ctx *ctx, int party, int blk, const uint32_t *atilde, void *acc_rows, size_t batch) {
    return step_impl(ctx, party, blk, atilde, acc_rows, batch, true);
}

int mktfhe_fft_batch(mktfhe_ctx *ctx, int bits, const void *polys, double *spectra, size_t batch) {
    int rc;
    if (!ctx) return MKTFHE_ERR_ARG;
    SINGLE_ONLY(ctx, "mktfhe_fft_batch");
    if ((rc = check_ready(ctx))) return rc;
    if (!polys || !spectra || (bits != 32 && bits != 64)) return fail(ctx, MKTFHE_ERR_ARG, "bad argument");
    return bits == 64 ? fft_hook<uint64_t>(ctx, false, polys, spectra, batch) : fft_hook<uint32_t>(ctx, false, polys, spectra, batch);
}

int mktfhe_ifft_batch(mktfhe_ctx *ctx, int bits, const double *spectra, void *polys, size_t batch) {
    int rc;
    if (!ctx) return MKTFHE_ERR_ARG;
    SINGLE_ONLY(ctx, "mktfhe_ifft_batch");
    if ((rc = check_ready(ctx))) return rc;
    if (!polys || !spectra || (bits != 32 && bits != 64)) return fail(ctx, MKTFHE_ERR_ARG, "bad argument");
    return bits == 64 ? fft_hook<uint64_t>(ctx, true, spectra, polys, batch) : fft_hook<uint32_t>(ctx, true, spectra, polys, batch);
}

int mktfhe_decomp_batch(mktfhe_ctx *ctx, int bits, int l, int logB, const void *polys, void *digits, size_t batch) {
    int rc;
    if (!ctx) return MKTFHE_ERR_ARG;
    SINGLE_ONLY(ctx, "mktfhe_decomp_batch");
    if ((rc = check_ready(ctx))) return rc;
    if (!polys || !digits || (bits != 32 && bits != 64) || l < 1 || l > MK_MAXL * 2 || logB < 1 || l * logB > bits)
        return fail(ctx, MKTFHE_ERR_ARG, "bad argument");
    return bits == 64 ? decomp_hook<uint64_t>(ctx, l, logB, polys, digits, batch) : decomp_hook<uint32_t>(ctx, l, logB, polys, digits, batch);
}

int mktfhe_gadget_product_batch(mktfhe_ctx *ctx, int l, int logB, const uint64_t *polys, const double *keys, int ncomp, uint64_t *out, size_t batch) {
    int rc;
    if (!ctx) return MKTFHE_ERR_ARG;
    SINGLE_ONLY(ctx, "mktfhe_gadget_product_batch");
    if ((rc = check_ready(ctx))) return rc;
    if (!ctx->kms || !fast_supported(ctx->p)) return fail(ctx, MKTFHE_ERR_PARAMS, "the gadget product hook exists for the FAST KMS path (N = 2048) only");
    if (!polys || !keys || !out || l < 1 || l > MK_MAXL || logB < 1 || l * logB > 64 || ncomp < 1 || ncomp > 3) return fail(ctx, MKTFHE_ERR_ARG, "bad argument");
    if (batch == 0) return 0;
    uint64_t *d_in = nullptr, *d_out = nullptr; cplx *d_keys = nullptr;
    const size_t nk = (size_t)l * ncomp * ctx->H;
    auto cleanup = [&]() { cudaFree(d_in); cudaFree(d_out); cudaFree(d_keys); };
    cudaError_t e;
    if ((e = cudaMalloc(&d_in, batch * ctx->N * 8)) != cudaSuccess || (e = cudaMalloc(&d_out, batch * ncomp * ctx->N * 8)) != cudaSuccess ||
        (e = cudaMalloc(&d_keys, nk * sizeof(cplx))) != cudaSuccess ||
        (e = cudaMemcpyAsync(d_in, polys, batch * ctx->N * 8, cudaMemcpyHostToDevice, ctx->stream)) != cudaSuccess ||
        (e = cudaMemcpyAsync(d_keys, keys, nk * sizeof(cplx), cudaMemcpyHostToDevice, ctx->stream)) != cudaSuccess) {
        cleanup();
        return fail(ctx, MKTFHE_ERR_CUDA, cudaGetErrorString(e));
    }
    rc = fast_gadget_product(ctx->fast, d_in, d_keys, d_out, l, logB, ncomp, batch, ctx->stream, &ctx->launches, ctx->err);
    if (!rc) {
        if ((e = cudaMemcpyAsync(out, d_out, batch * ncomp * ctx->N * 8, cudaMemcpyDeviceToHost, ctx->stream)) != cudaSuccess ||
            (e = cudaStreamSynchronize(ctx->stream)) != cudaSuccess)
            rc = fail(ctx, MKTFHE_ERR_CUDA, cudaGetErrorString(e));
    }
    cleanup();
    return rc;
}

int mktfhe_gadget_product32_batch(mktfhe_ctx *ctx, int l, int logB, const uint32_t *polys, const double *keys, int ncomp, uint32_t *out, size_t batch) {
    int rc;
    if (!ctx) return MKTFHE_ERR_ARG;
    SINGLE_ONLY(ctx, "mktfhe_gadget_product32_batch");
    if ((rc = check_ready(ctx))) return rc;
    if (ctx->N != 1024 || !ctx->fast32.t2) return fail(ctx, MKTFHE_ERR_PARAMS, "the 32-bit gadget product hook exists for the FAST N = 1024 paths (CGGI, LMSS, CCS) only");
    if (!polys || !keys || !out || l < 1 || l > MK_MAXL || logB < 1 || l * logB > 32 || ncomp < 1 || ncomp > 3) return fail(ctx, MKTFHE_ERR_ARG, "bad argument");
    if (batch == 0) return 0;
    uint32_t *d_in = nullptr, *d_out = nullptr; cplx *d_keys = nullptr;
    const size_t nk = (size_t)l * ncomp * ctx->H;
    auto cleanup = [&]() { cudaFree(d_in); cudaFree(d_out); cudaFree(d_keys); };
    cudaError_t e;
    if ((e = cudaMalloc(&d_in, batch * ctx->N * 4)) != cudaSuccess || (e = cudaMalloc(&d_out, batch * ncomp * ctx->N * 4)) != cudaSuccess ||
        (e = cudaMalloc(&d_keys, nk * sizeof(cplx))) != cudaSuccess ||
        (e = cudaMemcpyAsync(d_in, polys, batch * ctx->N * 4, cudaMemcpyHostToDevice, ctx->stream)) != cudaSuccess ||
        (e = cudaMemcpyAsync(d_keys, keys, nk * sizeof(cplx), cudaMemcpyHostToDevice, ctx->stream)) != cudaSuccess) {
        cleanup();
        return fail(ctx, MKTFHE_ERR_CUDA, cudaGetErrorString(e));
    }
    rc = fast32_gadget_product(ctx->fast32, d_in, d_keys, d_out, l, logB, ncomp, batch, ctx->stream, &ctx->launches, ctx->err);
    if (!rc) {
        if ((e = cudaMemcpyAsync(out, d_out, batch * ncomp * ctx->N * 4, cudaMemcpyDeviceToHost, ctx->stream)) != cudaSuccess ||
            (e = cudaStreamSynchronize(ctx->stream)) != cudaSuccess)
            rc = fail(ctx, MKTFHE_ERR_CUDA, cudaGetErrorString(e));
    }
    cleanup();
    return rc;
}

int mktfhe_measure_dfma_peak(mktfhe_ctx *ctx, double *tflops_out) {
    if (!ctx || !tflops_out) return MKTFHE_ERR_ARG;
    if (is_multi(ctx)) return mktfhe_measure_dfma_peak(ctx->children[0], tflops_out);
    CK(cudaSetDevice(ctx->device));
    int sms = 0;
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, ctx->device));
    const int blocks = sms, threads = 512, iters = 1 << 15;
    DevTmp<double> d;
    CK(d.alloc(sizeof(double) * blocks * threads));
    struct Ev { cudaEvent_t e = nullptr; ~Ev() { if (e) cudaEventDestroy(e); } } e0, e1;
    CK(cudaEventCreate(&e0.e)); CK(cudaEventCreate(&e1.e));
    k_dfma_peak<<<blocks, threads, 0, ctx->stream>>>(d, iters);     // warm-up
    float best = 1e30f;
    for (int r = 0; r < 3; r++) {
        CK(cudaEventRecord(e0.e, ctx->stream));
        k_dfma_peak<<<blocks, threads, 0, ctx->stream>>>(d, iters);
        CK(cudaEventRecord(e1.e, ctx->stream));
        CK(cudaEventSynchronize(e1.e));
        float ms = 0.f;
        CK(cudaEventElapsedTime(&ms, e0.e, e1.e));
        if (ms < best) best = ms;
    }
    *tflops_out = (double)blocks * threads * iters * 16 * 2 / (best * 1e-3) / 1e12;
    return 0;
}

}  // extern "C"
