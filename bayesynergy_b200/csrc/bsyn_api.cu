// bsyn_api.cu -- the C ABI of include/bsyn.h: host-side packing of problems into HBM, kernel dispatch.
// Replaces the Rcpp module boundary rstan::stan_fit<model_gp_grid, ...> (reference
// src/stanExports_gp_grid.cc:10-30) -- see INTEGRATION.md.  No CPU fallback: every compute entry point
// needs a CUDA device.
#include "../../include/bsyn.h"

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "bsyn_nuts.cuh"
#include "bsyn_post.cuh"
#include "bsyn_bridge.cuh"
#include "bsyn_kernels_decl.h"

using namespace bsyn;

// defaults of the specialised sampler variants (measured: profiles/r02_*)
#ifndef BSYN_FAST_DEFAULT_NW
#define BSYN_FAST_DEFAULT_NW 12
#endif
#ifndef BSYN_DEFAULT_TICK_PERIOD
#define BSYN_DEFAULT_TICK_PERIOD 2
#endif
#ifndef BSYN_FAST_DEFAULT_TICK
#define BSYN_FAST_DEFAULT_TICK 1
#endif

static thread_local std::string g_err;
static std::atomic<long long> g_launches{0};   // batches on different host threads share it

#define CK(call)                                                                                      \
    do {                                                                                              \
        cudaError_t e_ = (call);                                                                      \
        if (e_ != cudaSuccess) {                                                                      \
            g_err = std::string(#call) + ": " + cudaGetErrorString(e_);                               \
            return BSYN_ERR_CUDA;                                                                     \
        }                                                                                             \
    } while (0)

static int fail(int code, const std::string &msg)
{
    g_err = msg;
    return code;
}


// device allocation that frees itself on every return path
template <class T> struct DevBuf {
    T *p = nullptr;
    DevBuf() = default;
    DevBuf(const DevBuf &) = delete;
    DevBuf &operator=(const DevBuf &) = delete;
    ~DevBuf() { if (p) cudaFree(p); }
    cudaError_t alloc(size_t n) { return cudaMalloc(&p, sizeof(T) * std::max<size_t>(n, 1)); }
    operator T *() const { return p; }
};

// a pair of timing events that is destroyed on every return path
struct EvPair {
    cudaEvent_t a = nullptr, b = nullptr;
    EvPair() = default;
    EvPair(const EvPair &) = delete;
    EvPair &operator=(const EvPair &) = delete;
    ~EvPair() { if (a) cudaEventDestroy(a); if (b) cudaEventDestroy(b); }
    cudaError_t create() { const cudaError_t e = cudaEventCreate(&a); return e != cudaSuccess ? e : cudaEventCreate(&b); }
};

struct bsyn_batch {
    int device = 0;
    int n = 0;
    int cpl = 0;
    Flags F{};
    bsyn_options opts{};
    std::vector<ProbDesc> descs;
    ProbDesc *d_descs = nullptr;
    double *d_dpool = nullptr;
    int *d_ipool = nullptr;
    int *d_tri = nullptr;
    int maxD = 0, maxOut = 0, maxN = 0, max_d_smem = 0, max_i_smem = 0;
    const VariantOps *fast = nullptr;   // compile-time specialised kernels when the options and every grid match one
    cudaStream_t stream = nullptr;
    int num_sms = 0;
    // buffers of the last sampling call (device)
    double *d_qdraws = nullptr, *d_sp = nullptr, *d_summary = nullptr, *d_chain_stats = nullptr;
    double *d_arena = nullptr, *d_dense_L = nullptr, *d_dense_m2 = nullptr, *d_dense_mf = nullptr;
    size_t dense_L_elems = 0, dense_m2_elems = 0, dense_mf_elems = 0;
    size_t qdraws_elems = 0, sp_elems = 0, arena_elems = 0, summary_elems = 0, cs_elems = 0;
    int last_chains = 0, last_nsave = 0, last_nskip = 0;   // last_nsave: draw slots per chain (the row stride of qdraws)
    double *d_resume = nullptr; size_t resume_elems = 0;
    int *d_list = nullptr, *d_nsp = nullptr;
    bool last_var_ns = false;   // the last run left a different number of draws per problem (d_nsp)
    unsigned long long *d_gradctr = nullptr;
    int *d_workctr = nullptr;
};

// ---------------------------------------------------------------------------------------------------
static double host_inv_Phi(double p)
{
    // bisection + Newton on Phi(x) = erfc(-x/sqrt2)/2 (only used once per batch: tauLPTN, gp_grid.stan:37)
    double lo = -40, hi = 40;
    for (int it = 0; it < 200; ++it) {
        double mid = 0.5 * (lo + hi);
        if (0.5 * erfc(-mid / sqrt(2.0)) < p) lo = mid; else hi = mid;
    }
    double x = 0.5 * (lo + hi);
    for (int it = 0; it < 4; ++it) {
        double f = 0.5 * erfc(-x / sqrt(2.0)) - p;
        double d = exp(-0.5 * x * x) / sqrt(2.0 * M_PI);
        if (d > 0) x -= f / d;
    }
    return x;
}

static int make_flags(const bsyn_options &o, Flags &F)
{
    if (o.model != BSYN_MODEL_GP_GRID && o.model != BSYN_MODEL_NOINTERACTION) return fail(BSYN_ERR_ARG, "options.model must be 0 or 1");
    F.model = o.model;
    F.est_la = o.est_la ? 1 : 0;
    F.robust = o.robust ? 1 : 0;
    F.pcprior = o.pcprior ? 1 : 0;
    F.het = o.heteroscedastic ? 1 : 0;
    F.kernel = o.kernel; F.nu_matern = o.nu_matern; F.est_alpha = o.est_alpha ? 1 : 0;
    if (o.model == BSYN_MODEL_GP_GRID) {
        if (o.kernel < 1 || o.kernel > 3) return fail(BSYN_ERR_ARG, "options.kernel must be 1 (RBF), 2 (Matern) or 3 (RQ)");
        if (o.kernel == 2 && (o.nu_matern < 1 || o.nu_matern > 3)) return fail(BSYN_ERR_ARG, "options.nu_matern must be 1, 2 or 3 for the Matern kernel");
        if (o.kernel == 3 && !o.est_alpha) return fail(BSYN_ERR_ARG, "rational quadratic kernel needs est_alpha = 1 (alpha[1] is read, gp_grid.stan:139)");
        if (o.kernel != 3 && o.est_alpha) return fail(BSYN_ERR_ARG, "est_alpha = 1 only makes sense with kernel = 3");
    } else {
        F.kernel = 0; F.nu_matern = 0; F.est_alpha = 0; F.pcprior = 0;
    }
    F.lambda = o.lambda;
    F.tau = 0; F.lamL = 0; F.lptn_c0 = 0;
    if (F.robust) {
        if (!(o.rho > 0 && o.rho < 1)) return fail(BSYN_ERR_ARG, "rho must lie in (0,1)");
        F.tau = host_inv_Phi((1.0 + o.rho) / 2.0);
        F.lamL = 2.0 / (1.0 - o.rho) * exp(-0.5 * F.tau * F.tau - 0.5 * log(2.0 * M_PI)) * F.tau * log(F.tau);
        F.lptn_c0 = -0.5 * F.tau * F.tau + log(F.tau) + (F.lamL + 1.0) * log(log(F.tau));
    }
    F.lam1 = F.lam2 = F.log_lam12 = 0;
    if (F.pcprior) {
        const double *h = o.pcprior_hypers;
        F.lam1 = -log(h[1]) * h[0];      // gp_grid.stan:43  (d_half = 1)
        F.lam2 = -log(h[3]) / h[2];      // gp_grid.stan:44
        F.log_lam12 = log(F.lam1) + log(F.lam2);
    }
    unsigned m = 0;
    if (F.est_la) m |= (1u << SL_LA1) | (1u << SL_LA2);
    m |= (1u << SL_M1) | (1u << SL_M2) | (1u << SL_TH1) | (1u << SL_TH2) | (1u << SL_SL1) | (1u << SL_SL2);
    if (F.model == 0) {
        m |= (1u << SL_B1) | (1u << SL_B2) | (1u << SL_ELL) | (1u << SL_SIGF);
        if (F.est_alpha) m |= (1u << SL_ALPHA);
    }
    m |= (1u << SL_S) | (1u << SL_S21) | (1u << SL_S22);
    F.active_mask = m;
    return BSYN_OK;
}

template <int CPL> static bool fits(int n1, int n2)
{
    return n1 * n2 <= Cfg<CPL>::GM && n1 <= Cfg<CPL>::NM && n2 <= Cfg<CPL>::NM && n1 + n2 <= 31;
}

static const VariantOps *variant(int cpl);
const VariantOps *bsyn_variant_default_1();
const VariantOps *bsyn_variant_default_2();
const VariantOps *bsyn_variant_default_4();
const VariantOps *bsyn_variant_default_8();
const VariantOps *bsyn_variant_robust_pc_4();
const VariantOps *bsyn_variant_default(int cpl)
{
    switch (cpl) {
        case 1: return bsyn_variant_default_1();
        case 2: return bsyn_variant_default_2();
        case 4: return bsyn_variant_default_4();
        default: return bsyn_variant_default_8();
    }
}

// shared memory layouts.  per_warp_problem: every warp has its own problem block (logp / write_array
// kernels); otherwise one block per CTA (sampler).
// the kernels a call uses: the specialised variant when the batch matches one (and the caller can take it), else the
// runtime-flags variant of the batch's cells-per-lane class
static const VariantOps *ops_for(const bsyn_batch *b, bool allow_fast = true)
{
    return (allow_fast && b->fast) ? b->fast : variant(b->cpl);
}

// vpark_vs > 0 (sampler): also reserve 3 D-vectors of vpark_vs doubles per warp for parking S, p_prev and p across an
// evaluation -- when that still fits in the 227 KB of an SM; otherwise vpark_off = -1 and the arena is used
static SmemLayout make_layout(const bsyn_batch *b, const VariantOps *ops, bool per_warp_problem, size_t *bytes, int warps = WPC,
                              int vpark_vs = 0)
{
    SmemLayout sl;
    const int tr = ops->tr;
    auto even = [](int v) { return (v + 1) & ~1; };   // LDS.128 / STS.128 need 16-byte aligned double offsets
    sl.tri_off = 0;
    sl.pd_off = even((tr + 1) / 2 + 1);
    sl.pd_stride = even(b->max_d_smem);
    sl.pi_stride = even((b->max_i_smem + 1) / 2 + 1);   // in doubles
    const int nblk = per_warp_problem ? warps : 1;
    sl.pi_off = sl.pd_off + nblk * sl.pd_stride;
    sl.ws_off = even(sl.pi_off + nblk * sl.pi_stride);
    sl.park_off = even(sl.ws_off + warps * ops->ws_doubles);
    size_t tot = (size_t)(sl.park_off + warps * PARK_DOUBLES);
    sl.vpark_off = -1;
    if (vpark_vs > 0 && sizeof(double) * (tot + (size_t)warps * 3 * vpark_vs) <= 227 * 1024) {
        sl.vpark_off = (int)tot;
        tot += (size_t)warps * 3 * vpark_vs;
    }
    *bytes = sizeof(double) * tot;
    return sl;
}

extern "C" const char *bsyn_last_error(void) { return g_err.c_str(); }
extern "C" int bsyn_version(void) { return BSYN_VERSION; }
extern "C" int64_t bsyn_launch_count(void) { return g_launches.load(); }

extern "C" int bsyn_batch_create(const bsyn_problem *problems, int32_t n, const bsyn_options *opts, int32_t device,
                                 bsyn_batch **out)
{
    if (!problems || !opts || !out || n <= 0) return fail(BSYN_ERR_ARG, "bsyn_batch_create: null argument or n <= 0");
    *out = nullptr;
    Flags F;
    int rc = make_flags(*opts, F);
    if (rc) return rc;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
        return fail(BSYN_ERR_CUDA, "no CUDA device available (this library has no CPU fallback)");
    if (device < 0 || device >= ndev) return fail(BSYN_ERR_ARG, "bsyn_batch_create: bad device ordinal");
    CK(cudaSetDevice(device));
    bsyn_batch *b = new bsyn_batch();
    struct Guard {   // every early return below destroys the half-built batch (stream, device buffers)
        bsyn_batch *b;
        ~Guard() { if (b) bsyn_batch_destroy(b); }
    } guard{b};
    b->device = device; b->n = n; b->F = F; b->opts = *opts;
    // pick the smallest kernel variant that fits every problem
    int cpl = 1;
    for (int k = 0; k < n; ++k) {
        const bsyn_problem &p = problems[k];
        if (p.n1 < 1 || p.n2 < 1 || p.N < 0 || !p.y || !p.ii_obs || !p.x1 || !p.x2) { return fail(BSYN_ERR_ARG, "problem " + std::to_string(k) + ": bad sizes or null pointer"); }
        const long long Nexp = (long long)(p.n1 + p.n2 + p.n1 * p.n2 + 1) * p.nrep - p.nmissing;
        if (Nexp != p.N) { return fail(BSYN_ERR_ARG, "problem " + std::to_string(k) + ": N != (n1+n2+n1*n2+1)*nrep - nmissing (gp_grid.stan:31)"); }
        while (cpl <= 8) {
            bool f = cpl == 1 ? fits<1>(p.n1, p.n2) : cpl == 2 ? fits<2>(p.n1, p.n2) : cpl == 4 ? fits<4>(p.n1, p.n2) : fits<8>(p.n1, p.n2);
            if (f) break;
            cpl *= 2;
        }
        if (cpl > 8) { return fail(BSYN_ERR_UNSUPPORTED, "problem " + std::to_string(k) + ": dose grid larger than 16 x 15 is not supported by the compiled kernel variants"); }
    }
    b->cpl = cpl;
    {   // compile-time specialised kernels (bsyn_f10.cu / bsyn_h10.cu): package-default options on 10 x 10 grids
        bool all10 = true;
        for (int k = 0; k < n; ++k) all10 = all10 && problems[k].n1 == 10 && problems[k].n2 == 10;
        const bool def_lik = F.het == 1 && F.robust == 0 && F.est_la == 1;
        const char *nf = getenv("BSYN_NO_FAST");
        const bool def_gp = F.model == 0 && F.kernel == 2 && F.nu_matern == 2 && F.pcprior == 0 && F.est_alpha == 0;
        if (F.model == 0 && F.kernel == 2 && F.nu_matern == 2 && F.pcprior == 1 && F.est_alpha == 0 && F.het == 1 && F.robust == 1 &&
            F.est_la == 1 && cpl == 4 && !(nf && atoi(nf)))
            b->fast = bsyn_variant_robust_pc_4();
        if (def_lik && !(nf && atoi(nf))) {
            if (def_gp) b->fast = all10 ? bsyn_variant_gp10() : bsyn_variant_default(cpl);
            if (F.model == 1 && all10) b->fast = bsyn_variant_h0_10();
        }
    }
    std::vector<double> dpool;
    std::vector<int> ipool;
    b->descs.resize(n);
    for (int k = 0; k < n; ++k) {
        const bsyn_problem &p = problems[k];
        ProbDesc &d = b->descs[k];
        const int n1 = p.n1, n2 = p.n2, N = p.N, ncf = (n1 + 1) * (n2 + 1), nm = n1 + n2;
        const int P1 = n1 * (n1 - 1) / 2, P2 = n2 * (n2 - 1) / 2, PP = P1 + P2;
        d.n1 = n1; d.n2 = n2; d.N = N; d.G = n1 * n2; d.ncf = ncf;
        const int gp = (F.model == 0);
        d.D = gp ? 13 + 2 * F.est_la + F.est_alpha + n1 * n2 : 9 + 2 * F.est_la;
        d.n_out = gp ? d.D + 3 * d.G + nm + 2 + N + 7 : d.D + d.G + nm + 2 + N + 4;
        if (dpool.size() & 1) dpool.push_back(0.0);
        d.off_d = (int)dpool.size();
        d.off_i = (int)ipool.size();
        for (int j = 0; j < n1; ++j) dpool.push_back(p.x1[j]);
        for (int i = 0; i < n2; ++i) dpool.push_back(p.x2[i]);
        // VUS trapezoid weights (include/VUS.stan): VUS(Z) = sum_ij w1[j] w2[i] Z[i,j]; 100/area folded into w1
        {
            double mn1 = p.x1[0], mx1 = p.x1[0], mn2 = p.x2[0], mx2 = p.x2[0];
            for (int j = 0; j < n1; ++j) { mn1 = std::min(mn1, p.x1[j]); mx1 = std::max(mx1, p.x1[j]); }
            for (int i = 0; i < n2; ++i) { mn2 = std::min(mn2, p.x2[i]); mx2 = std::max(mx2, p.x2[i]); }
            const double area = (mx1 - mn1) * (mx2 - mn2);
            for (int j = 0; j < n1; ++j) {
                double w = 0;
                if (j > 0) w += 0.5 * (p.x1[j] - p.x1[j - 1]);
                if (j + 1 < n1) w += 0.5 * (p.x1[j + 1] - p.x1[j]);
                dpool.push_back(100.0 * w / area);
            }
            for (int i = 0; i < n2; ++i) {
                double w = 0;
                if (i > 0) w += 0.5 * (p.x2[i] - p.x2[i - 1]);
                if (i + 1 < n2) w += 0.5 * (p.x2[i + 1] - p.x2[i]);
                dpool.push_back(w);
            }
        }
        // pair distances (gp_grid.stan:47-62) and their (a,b) codes, lower triangle a > b
        for (int a = 0; a < n1; ++a)
            for (int c = 0; c < a; ++c) { dpool.push_back(sqrt((p.x1[a] - p.x1[c]) * (p.x1[a] - p.x1[c]))); ipool.push_back((a << 8) | c); }
        for (int a = 0; a < n2; ++a)
            for (int c = 0; c < a; ++c) { dpool.push_back(sqrt((p.x2[a] - p.x2[c]) * (p.x2[a] - p.x2[c]))); ipool.push_back((a << 8) | c); }
        // per-cell sufficient statistics + observations sorted by cell
        std::vector<double> cnt(ncf, 0.0), sum(ncf, 0.0), ss(ncf, 0.0);
        std::vector<int> cstart(ncf + 1, 0);
        for (int t = 0; t < N; ++t) {
            const int c = p.ii_obs[t] - 1;
            if (c < 0 || c >= ncf) { return fail(BSYN_ERR_ARG, "problem " + std::to_string(k) + ": ii_obs out of range"); }
            if (!(p.y[t] == p.y[t])) { return fail(BSYN_ERR_ARG, "problem " + std::to_string(k) + ": y contains NaN (missing values must be removed, R/bayesynergy.R:153-154)"); }
            cnt[c] += 1.0; sum[c] += p.y[t];
        }
        for (int c = 0; c < ncf; ++c) { if (cnt[c] > 0) sum[c] /= cnt[c]; cstart[c + 1] = cstart[c] + (int)cnt[c]; }
        std::vector<double> ysort(N);
        std::vector<int> fill(cstart.begin(), cstart.end() - 1);
        for (int t = 0; t < N; ++t) {
            const int c = p.ii_obs[t] - 1;
            const double dy = p.y[t] - sum[c];
            ss[c] += dy * dy;
            ysort[fill[c]++] = p.y[t];
        }
        dpool.insert(dpool.end(), cnt.begin(), cnt.end());
        dpool.insert(dpool.end(), sum.begin(), sum.end());
        dpool.insert(dpool.end(), ss.begin(), ss.end());
        d.n_d_smem = 2 * nm + PP + 3 * ncf + (F.robust ? N : 0);
        dpool.insert(dpool.end(), ysort.begin(), ysort.end());
        for (int t = 0; t < N; ++t) dpool.push_back(p.y[t]);
        ipool.insert(ipool.end(), cstart.begin(), cstart.end());
        d.n_i_smem = PP + (F.robust ? ncf + 1 : 0);
        for (int t = 0; t < N; ++t) ipool.push_back(p.ii_obs[t] - 1);
        b->maxD = std::max(b->maxD, d.D); b->maxOut = std::max(b->maxOut, d.n_out); b->maxN = std::max(b->maxN, N);
        b->max_d_smem = std::max(b->max_d_smem, d.n_d_smem); b->max_i_smem = std::max(b->max_i_smem, d.n_i_smem);
    }
    // packed lower-triangular (r,c) table, row-major; valid for every n <= 16
    std::vector<int> tri;
    for (int r = 0; r < 16; ++r)
        for (int c = 0; c <= r; ++c) tri.push_back((r << 8) | c);
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, device));
    b->num_sms = prop.multiProcessorCount;
    CK(cudaStreamCreateWithFlags(&b->stream, cudaStreamNonBlocking));
    CK(cudaMalloc(&b->d_descs, sizeof(ProbDesc) * n));
    CK(cudaMalloc(&b->d_dpool, sizeof(double) * std::max<size_t>(dpool.size(), 1)));
    CK(cudaMalloc(&b->d_ipool, sizeof(int) * std::max<size_t>(ipool.size(), 1)));
    CK(cudaMalloc(&b->d_tri, sizeof(int) * tri.size()));
    CK(cudaMalloc(&b->d_gradctr, sizeof(unsigned long long)));
    CK(cudaMalloc(&b->d_workctr, sizeof(int)));
    CK(cudaMemcpy(b->d_descs, b->descs.data(), sizeof(ProbDesc) * n, cudaMemcpyHostToDevice));
    // the packed pools are pinned for the upload (page-locked DMA straight from the staging vectors)
    const bool pin_d = cudaHostRegister(dpool.data(), sizeof(double) * dpool.size(), cudaHostRegisterDefault) == cudaSuccess;
    const bool pin_i = cudaHostRegister(ipool.data(), sizeof(int) * ipool.size(), cudaHostRegisterDefault) == cudaSuccess;
    cudaGetLastError();
    cudaError_t ce = cudaMemcpyAsync(b->d_dpool, dpool.data(), sizeof(double) * dpool.size(), cudaMemcpyHostToDevice, b->stream);
    if (ce == cudaSuccess) ce = cudaMemcpyAsync(b->d_ipool, ipool.data(), sizeof(int) * ipool.size(), cudaMemcpyHostToDevice, b->stream);
    if (ce == cudaSuccess) ce = cudaStreamSynchronize(b->stream);
    if (pin_d) cudaHostUnregister(dpool.data());
    if (pin_i) cudaHostUnregister(ipool.data());
    if (ce != cudaSuccess) { g_err = std::string("upload of the problem pools: ") + cudaGetErrorString(ce); return BSYN_ERR_CUDA; }
    CK(cudaMemcpy(b->d_tri, tri.data(), sizeof(int) * tri.size(), cudaMemcpyHostToDevice));
    guard.b = nullptr;
    *out = b;
    return BSYN_OK;
}

extern "C" void bsyn_batch_destroy(bsyn_batch *b)
{
    if (!b) return;
    cudaSetDevice(b->device);
    cudaFree(b->d_descs); cudaFree(b->d_dpool); cudaFree(b->d_ipool); cudaFree(b->d_tri);
    cudaFree(b->d_qdraws); cudaFree(b->d_sp); cudaFree(b->d_summary); cudaFree(b->d_chain_stats); cudaFree(b->d_arena); cudaFree(b->d_dense_L); cudaFree(b->d_dense_m2); cudaFree(b->d_dense_mf);
    cudaFree(b->d_gradctr); cudaFree(b->d_workctr); cudaFree(b->d_resume); cudaFree(b->d_list); cudaFree(b->d_nsp);
    if (b->stream) cudaStreamDestroy(b->stream);
    delete b;
}

extern "C" int32_t bsyn_num_params(const bsyn_batch *b, int32_t k) { return (b && k >= 0 && k < b->n) ? b->descs[k].D : -1; }
extern "C" int32_t bsyn_num_outputs(const bsyn_batch *b, int32_t k) { return (b && k >= 0 && k < b->n) ? b->descs[k].n_out : -1; }
// Flat name of entry `index` of one write_array row of problem k (the first bsyn_num_params entries are the
// parameters): declaration order of src/stanExports_gp_grid.h:1588-1622 / :2256-2402, matrices column-major.
extern "C" int bsyn_output_name(const bsyn_batch *b, int32_t k, int32_t index, char *buf, int32_t buflen)
{
    if (!b || k < 0 || k >= b->n || !buf || buflen < 2) return fail(BSYN_ERR_ARG, "bsyn_output_name: bad argument");
    const ProbDesc &d = b->descs[k];
    if (index < 0 || index >= d.n_out) return fail(BSYN_ERR_ARG, "bsyn_output_name: index out of range");
    const bool gp = b->opts.model == BSYN_MODEL_GP_GRID;
    const int n1 = d.n1, n2 = d.n2, G = n1 * n2;
    int i = index;
    auto put = [&](const char *fmt, int a, int c) { snprintf(buf, (size_t)buflen, fmt, a, c); return BSYN_OK; };
    auto mat = [&](const char *name) {   // (row = drug-2 level, column = drug-1 level), column-major
        char f[32];
        snprintf(f, sizeof f, "%s[%%d,%%d]", name);
        return put(f, i % n2 + 1, i / n2 + 1);
    };
    auto vec = [&](const char *name) { char f[32]; snprintf(f, sizeof f, "%s[%%d]", name); return put(f, i + 1, 0); };
    static const char *pre[] = {"la_1[1]", "la_2[1]", "log10_ec50_1", "log10_ec50_2", "theta_1", "theta_2", "slope_1", "slope_2",
                                "b1", "b2", "ell", "sigma_f", "alpha[1]"};
    const int skip = b->opts.est_la ? 0 : 2;
    const int npre = 8 - skip + (gp ? 4 + (b->opts.est_alpha ? 1 : 0) : 0);
    if (i < npre) return put(pre[i + skip], 0, 0);
    i -= npre;
    if (gp) { if (i < G) return mat("z"); i -= G; }
    static const char *post[] = {"s", "s2_log10_ec50_1", "s2_log10_ec50_2"};
    if (i < 3) return put(post[i], 0, 0);
    i -= 3;
    if (i < G) return mat("p0");
    i -= G;
    if (i < n1) return vec("p01");
    i -= n1;
    if (i < n2) return vec("p02");
    i -= n2;
    if (gp) {
        if (i < G) return mat("Delta");
        i -= G;
        if (i < G) return mat("GP");
        i -= G;
    }
    if (i < 2) return put(i ? "ec50_2" : "ec50_1", 0, 0);
    i -= 2;
    if (i < d.N) return vec("CPO");
    i -= d.N;
    static const char *tail[] = {"dss_1", "dss_2", "rVUS_f", "rVUS_p0", "VUS_Delta", "VUS_syn", "VUS_ant"};
    return put(tail[i], 0, 0);
}

// Parameter blocks = what param_names() / param_dims() of the module report (get_param_names / get_dims,
// src/stanExports_gp_grid.h:1588-1704; nointeraction: :966-1040): every declared name with its dimensions, including the
// zero-length la_1 / la_2 / alpha vectors when they are not estimated, then "lp__".
struct BlockDef { const char *name; int kind; };   // kind: 0 scalar, 1 [est_la], 2 [est_alpha], 3 [n2,n1], 4 [n1], 5 [n2], 6 [N]
static const BlockDef g_blocks_gp[] = {
    {"la_1", 1}, {"la_2", 1}, {"log10_ec50_1", 0}, {"log10_ec50_2", 0}, {"theta_1", 0}, {"theta_2", 0}, {"slope_1", 0}, {"slope_2", 0},
    {"b1", 0}, {"b2", 0}, {"ell", 0}, {"sigma_f", 0}, {"alpha", 2}, {"z", 3}, {"s", 0}, {"s2_log10_ec50_1", 0}, {"s2_log10_ec50_2", 0},
    {"p0", 3}, {"p01", 4}, {"p02", 5}, {"Delta", 3}, {"GP", 3}, {"ec50_1", 0}, {"ec50_2", 0}, {"CPO", 6}, {"dss_1", 0}, {"dss_2", 0},
    {"rVUS_f", 0}, {"rVUS_p0", 0}, {"VUS_Delta", 0}, {"VUS_syn", 0}, {"VUS_ant", 0}, {"lp__", 0}};
static const BlockDef g_blocks_h0[] = {
    {"la_1", 1}, {"la_2", 1}, {"log10_ec50_1", 0}, {"log10_ec50_2", 0}, {"theta_1", 0}, {"theta_2", 0}, {"slope_1", 0}, {"slope_2", 0},
    {"s", 0}, {"s2_log10_ec50_1", 0}, {"s2_log10_ec50_2", 0}, {"p0", 3}, {"p01", 4}, {"p02", 5}, {"ec50_1", 0}, {"ec50_2", 0}, {"CPO", 6},
    {"dss_1", 0}, {"dss_2", 0}, {"rVUS_f", 0}, {"rVUS_p0", 0}, {"lp__", 0}};

extern "C" int32_t bsyn_num_param_blocks(const bsyn_batch *b)
{
    if (!b) return -1;
    return b->opts.model == BSYN_MODEL_GP_GRID ? (int)(sizeof g_blocks_gp / sizeof g_blocks_gp[0]) : (int)(sizeof g_blocks_h0 / sizeof g_blocks_h0[0]);
}
extern "C" int bsyn_param_block(const bsyn_batch *b, int32_t k, int32_t block, char *name, int32_t buflen, int32_t *ndims,
                                int32_t *dims, int32_t *flat_offset)
{
    if (!b || k < 0 || k >= b->n || block < 0 || block >= bsyn_num_param_blocks(b) || !name || buflen < 2 || !ndims || !dims || !flat_offset)
        return fail(BSYN_ERR_ARG, "bsyn_param_block: bad argument");
    const BlockDef *tab = b->opts.model == BSYN_MODEL_GP_GRID ? g_blocks_gp : g_blocks_h0;
    const ProbDesc &d = b->descs[k];
    int off = 0;
    for (int i = 0;; ++i) {
        int nd = 0, dm[2] = {0, 0}, len = 1;
        switch (tab[i].kind) {
            case 1: nd = 1; dm[0] = b->opts.est_la ? 1 : 0; len = dm[0]; break;
            case 2: nd = 1; dm[0] = b->opts.est_alpha ? 1 : 0; len = dm[0]; break;
            case 3: nd = 2; dm[0] = d.n2; dm[1] = d.n1; len = d.n1 * d.n2; break;
            case 4: nd = 1; dm[0] = d.n1; len = d.n1; break;
            case 5: nd = 1; dm[0] = d.n2; len = d.n2; break;
            case 6: nd = 1; dm[0] = d.N; len = d.N; break;
            default: break;
        }
        if (i == block) {
            snprintf(name, (size_t)buflen, "%s", tab[i].name);
            *ndims = nd; dims[0] = dm[0]; dims[1] = dm[1]; *flat_offset = off;   // lp__ sits behind the write_array row
            return BSYN_OK;
        }
        off += len;
    }
}

extern "C" int32_t bsyn_max_params(const bsyn_batch *b) { return b ? b->maxD : -1; }
extern "C" int32_t bsyn_max_outputs(const bsyn_batch *b) { return b ? b->maxOut : -1; }
extern "C" int32_t bsyn_num_problems(const bsyn_batch *b) { return b ? b->n : -1; }

static const VariantOps *variant(int cpl)
{
    switch (cpl) {
        case 1: return bsyn_variant_1();
        case 2: return bsyn_variant_2();
        case 4: return bsyn_variant_4();
        default: return bsyn_variant_8();
    }
}

static int launch_logp(bsyn_batch *b, int B, const int *d_idx, const double *d_u, int ldu, int jac, double *d_lp,
                       double *d_grad, int *d_status, double *d_dbg, int dbg_stride, cudaStream_t st)
{
    size_t bytes;
    const VariantOps *ops = ops_for(b);
    SmemLayout sl = make_layout(b, ops, true, &bytes);
    int grid = std::min((B + WPC - 1) / WPC, b->num_sms * 8);
    if (grid < 1) grid = 1;
    CK(ops->logp(d_dbg != nullptr, grid, bytes, st, b->F, b->d_descs, b->d_dpool, b->d_ipool, b->d_tri, sl, B,
                             d_idx, d_u, ldu, jac, d_lp, d_grad, d_status, d_dbg, dbg_stride));
    ++g_launches;
    return BSYN_OK;
}

extern "C" int bsyn_logp_grad_device(bsyn_batch *b, int32_t B, const int32_t *d_prob_idx, const double *d_u, int32_t ldu,
                                     int32_t jacobian, double *d_lp, double *d_grad, int32_t *d_status, void *stream)
{
    if (!b || B < 0 || !d_prob_idx || !d_u || !d_lp || !d_status || ldu < b->maxD) return fail(BSYN_ERR_ARG, "bsyn_logp_grad_device: bad argument");
    if (B == 0) return BSYN_OK;
    CK(cudaSetDevice(b->device));
    return launch_logp(b, B, d_prob_idx, d_u, ldu, jacobian, d_lp, d_grad, d_status, nullptr, 0, (cudaStream_t)stream);
}

// debug variant used by tests: also returns the kernel's intermediates (see bsyn_model.cuh DBG layout)
extern "C" int bsyn_logp_grad_debug(bsyn_batch *b, int32_t B, const int32_t *prob_idx, const double *u, int32_t ldu,
                                    int32_t jacobian, double *lp, double *grad, int32_t *status, double *dbg,
                                    int32_t dbg_stride);

static int logp_host(bsyn_batch *b, int B, const int *prob_idx, const double *u, int ldu, int jac, double *lp, double *grad,
                     int *status, double *dbg, int dbg_stride)
{
    if (!b || B < 0 || !prob_idx || !u || !lp || !status || ldu < b->maxD) return fail(BSYN_ERR_ARG, "bsyn_logp_grad: bad argument (ldu must be >= bsyn_max_params)");
    if (B == 0) return BSYN_OK;
    for (int k = 0; k < B; ++k)
        if (prob_idx[k] < 0 || prob_idx[k] >= b->n) return fail(BSYN_ERR_ARG, "bsyn_logp_grad: prob_idx out of range");
    CK(cudaSetDevice(b->device));
    DevBuf<int> d_idx, d_st;
    DevBuf<double> d_u, d_lp, d_g, d_dbg;
    CK(d_idx.alloc(B)); CK(d_st.alloc(B));
    CK(d_u.alloc((size_t)B * ldu)); CK(d_lp.alloc(B));
    if (grad) { CK(d_g.alloc((size_t)B * ldu)); CK(cudaMemsetAsync(d_g, 0, sizeof(double) * (size_t)B * ldu, b->stream)); }
    if (dbg) { CK(d_dbg.alloc((size_t)B * dbg_stride)); CK(cudaMemsetAsync(d_dbg, 0, sizeof(double) * (size_t)B * dbg_stride, b->stream)); }
    CK(cudaMemcpyAsync(d_idx, prob_idx, sizeof(int) * B, cudaMemcpyHostToDevice, b->stream));
    CK(cudaMemcpyAsync(d_u, u, sizeof(double) * (size_t)B * ldu, cudaMemcpyHostToDevice, b->stream));
    int rc = launch_logp(b, B, d_idx, d_u, ldu, jac, d_lp, d_g, d_st, d_dbg, dbg_stride, b->stream);
    if (rc == BSYN_OK) {
        CK(cudaMemcpyAsync(lp, d_lp, sizeof(double) * B, cudaMemcpyDeviceToHost, b->stream));
        CK(cudaMemcpyAsync(status, d_st, sizeof(int) * B, cudaMemcpyDeviceToHost, b->stream));
        if (grad) CK(cudaMemcpyAsync(grad, d_g, sizeof(double) * (size_t)B * ldu, cudaMemcpyDeviceToHost, b->stream));
        if (dbg) CK(cudaMemcpyAsync(dbg, d_dbg, sizeof(double) * (size_t)B * dbg_stride, cudaMemcpyDeviceToHost, b->stream));
        CK(cudaStreamSynchronize(b->stream));
    }
    return rc;
}

extern "C" int bsyn_logp_grad(bsyn_batch *b, int32_t B, const int32_t *prob_idx, const double *u, int32_t ldu,
                              int32_t jacobian, double *lp, double *grad, int32_t *status)
{
    return logp_host(b, B, prob_idx, u, ldu, jacobian, lp, grad, status, nullptr, 0);
}
extern "C" int bsyn_logp_grad_debug(bsyn_batch *b, int32_t B, const int32_t *prob_idx, const double *u, int32_t ldu,
                                    int32_t jacobian, double *lp, double *grad, int32_t *status, double *dbg,
                                    int32_t dbg_stride)
{
    return logp_host(b, B, prob_idx, u, ldu, jacobian, lp, grad, status, dbg, dbg_stride);
}
extern "C" int bsyn_debug_layout(const bsyn_batch *b, int32_t *nm, int32_t *nn, int32_t *gm)
{
    if (!b) return BSYN_ERR_ARG;
    int NM = ops_for(b)->nm;
    *nm = NM; *nn = NM * NM; *gm = 32 * b->cpl;
    return BSYN_OK;
}

extern "C" int bsyn_write_array(bsyn_batch *b, int32_t B, const int32_t *prob_idx, const double *u, int32_t ldu,
                                double *out, int32_t ldo, uint64_t seed, int32_t *status)
{
    if (!b || B < 0 || !prob_idx || !u || !out || !status || ldu < b->maxD || ldo < b->maxOut) return fail(BSYN_ERR_ARG, "bsyn_write_array: bad argument");
    if (B == 0) return BSYN_OK;
    for (int k = 0; k < B; ++k)
        if (prob_idx[k] < 0 || prob_idx[k] >= b->n) return fail(BSYN_ERR_ARG, "bsyn_write_array: prob_idx out of range");
    CK(cudaSetDevice(b->device));
    DevBuf<int> d_idx, d_st;
    DevBuf<double> d_u, d_out;
    CK(d_idx.alloc(B)); CK(d_st.alloc(B));
    CK(d_u.alloc((size_t)B * ldu)); CK(d_out.alloc((size_t)B * ldo));
    CK(cudaMemsetAsync(d_out, 0, sizeof(double) * (size_t)B * ldo, b->stream));
    CK(cudaMemcpyAsync(d_idx, prob_idx, sizeof(int) * B, cudaMemcpyHostToDevice, b->stream));
    CK(cudaMemcpyAsync(d_u, u, sizeof(double) * (size_t)B * ldu, cudaMemcpyHostToDevice, b->stream));
    size_t bytes;
    SmemLayout sl = make_layout(b, variant(b->cpl), true, &bytes);
    int grid = std::max(1, std::min((B + WPC - 1) / WPC, b->num_sms * 8));
    CK(variant(b->cpl)->write_array(grid, bytes, b->stream, b->F, b->d_descs, b->d_dpool, b->d_ipool, b->d_tri, sl, B, d_idx,
                                    d_u, ldu, d_out, ldo, seed, d_st));
    ++g_launches;
    CK(cudaMemcpyAsync(out, d_out, sizeof(double) * (size_t)B * ldo, cudaMemcpyDeviceToHost, b->stream));
    CK(cudaMemcpyAsync(status, d_st, sizeof(int) * B, cudaMemcpyDeviceToHost, b->stream));
    CK(cudaStreamSynchronize(b->stream));
    return BSYN_OK;
}

// transform_inits (h:803-1075): constrained -> unconstrained, host arithmetic (lub_free / lb_free)
extern "C" int bsyn_unconstrain(const bsyn_batch *b, int32_t k, const double *theta, double *u)
{
    if (!b || k < 0 || k >= b->n || !theta || !u) return fail(BSYN_ERR_ARG, "bsyn_unconstrain: bad argument");
    const ProbDesc &d = b->descs[k];
    const Flags &F = b->F;
    int pos = 0;
    auto lb = [&](double v) { if (!(v >= 0)) return false; u[pos] = log(v); ++pos; return true; };
    if (F.est_la)
        for (int t = 0; t < 2; ++t) {
            double v = theta[pos];
            if (!(v >= 0 && v <= 1)) return fail(BSYN_ERR_ARG, "bsyn_unconstrain: la out of [0,1]");
            u[pos] = log(v / (1 - v)); ++pos;
        }
    for (int t = 0; t < 4; ++t) { u[pos] = theta[pos]; ++pos; }
    for (int t = 0; t < 2; ++t) if (!lb(theta[pos])) return fail(BSYN_ERR_ARG, "bsyn_unconstrain: slope < 0");
    if (F.model == 0) {
        for (int t = 0; t < 4 + F.est_alpha; ++t) if (!lb(theta[pos])) return fail(BSYN_ERR_ARG, "bsyn_unconstrain: negative scale parameter");
        for (int t = 0; t < d.G; ++t) { u[pos] = theta[pos]; ++pos; }
    }
    for (int t = 0; t < 3; ++t) if (!lb(theta[pos])) return fail(BSYN_ERR_ARG, "bsyn_unconstrain: negative variance parameter");
    return BSYN_OK;
}

// ---------------------------------------------------------------------------------------------------
extern "C" void bsyn_sampler_defaults(bsyn_sampler_args *a)
{
    a->chains = 4; a->iter = 2000; a->warmup = 1000; a->thin = 1; a->seed = 1;
    a->init_r = 2.0; a->adapt_delta = 0.9; a->stepsize = 1.0; a->stepsize_jitter = 0.0;
    a->max_treedepth = 10; a->metric = 0;
    a->adapt_gamma = 0.05; a->adapt_kappa = 0.75; a->adapt_t0 = 10.0;
    a->adapt_init_buffer = 75; a->adapt_term_buffer = 50; a->adapt_window = 25;
    a->save_warmup = 0;
}

static int ensure(double **p, size_t *have, size_t need)
{
    if (*have >= need && *p) return BSYN_OK;
    if (*p) cudaFree(*p);
    *p = nullptr; *have = 0;
    if (cudaMalloc(p, sizeof(double) * std::max<size_t>(need, 1)) != cudaSuccess) return fail(BSYN_ERR_ALLOC, "device allocation failed");
    *have = need;
    return BSYN_OK;
}

struct DevSinks {
    double *draws = nullptr, *udraws = nullptr, *cpo = nullptr, *inv_metric = nullptr, *init_u = nullptr, *f_acc = nullptr;
    long long ld_draws = 0, ld_u = 0;
    int ld_cpo = 0, ld_im = 0, ld_f = 0;
};

static int check_sampler_args(const bsyn_sampler_args *a)
{
    if (a->chains < 1 || a->iter < 1 || a->warmup < 0 || a->warmup > a->iter || a->thin < 1 || a->max_treedepth < 1 || a->max_treedepth > 20)
        return fail(BSYN_ERR_ARG, "bsyn_sample: bad sampler arguments");
    if (a->metric != 0 && a->metric != 1) return fail(BSYN_ERR_ARG, "bsyn_sample: metric must be 0 (diag_e) or 1 (dense_e)");
    return BSYN_OK;
}

// how a launch of the sampler sits inside a multi-launch run (bsyn_sample_to_ess); the default is one stand-alone launch
struct RunPlan {
    const int *d_list = nullptr; int n_list = 0;   // problems of this launch (device list), null = all
    int isave0 = 0, ld_save = 0;                   // first draw slot, slots per chain (0: = the draws of this launch)
    bool resume_in = false, resume_out = false, summary = true;
};

static int run_sampler(bsyn_batch *b, const bsyn_sampler_args *a, const DevSinks &ds, int64_t *total_grad, float *elapsed_ms,
                       const RunPlan &plan = RunPlan())
{
    if (int rc0 = check_sampler_args(a)) return rc0;
    CK(cudaSetDevice(b->device));
    SamplerParams sp{};
    sp.n_problems = b->n; sp.chains = a->chains;
    sp.iter = a->iter; sp.warmup = a->warmup; sp.thin = a->thin; sp.max_depth = a->max_treedepth; sp.save_warmup = a->save_warmup;
    const int ns_samp = (a->iter - a->warmup + a->thin - 1) / a->thin, ns_warm = a->save_warmup ? (a->warmup + a->thin - 1) / a->thin : 0;
    sp.n_save = ns_samp + ns_warm;
    sp.ld_save = plan.ld_save ? plan.ld_save : sp.n_save;
    sp.isave0 = plan.isave0; sp.prob_list = plan.d_list; sp.n_list = plan.n_list;
    sp.resume_in = plan.resume_in; sp.resume_out = plan.resume_out;
    if (sp.isave0 + sp.n_save > sp.ld_save) return fail(BSYN_ERR_ARG, "bsyn_sample: draws of this launch do not fit behind isave0");
    if ((plan.resume_in || plan.resume_out) && a->metric != 0) return fail(BSYN_ERR_UNSUPPORTED, "resuming chains needs metric = diag_e");
    sp.init_r = a->init_r; sp.delta = a->adapt_delta; sp.stepsize0 = a->stepsize; sp.jitter = a->stepsize_jitter;
    sp.gamma = a->adapt_gamma; sp.kappa = a->adapt_kappa; sp.t0 = a->adapt_t0;
    sp.init_buffer = a->adapt_init_buffer; sp.term_buffer = a->adapt_term_buffer; sp.window = a->adapt_window;
    sp.seed = a->seed; sp.jacobian = 1; sp.metric = a->metric;
    const size_t nrows = (size_t)b->n * a->chains * sp.ld_save;
    const bool first = plan.isave0 == 0 && !plan.resume_in;    // later launches of a run keep the earlier draws
    int rc;
    if ((rc = ensure(&b->d_qdraws, &b->qdraws_elems, nrows * BSYN_NQ))) return rc;
    if ((rc = ensure(&b->d_sp, &b->sp_elems, nrows * BSYN_NSP))) return rc;
    if ((rc = ensure(&b->d_chain_stats, &b->cs_elems, (size_t)b->n * a->chains * 8))) return rc;
    if ((rc = ensure(&b->d_summary, &b->summary_elems, (size_t)b->n * BSYN_NQ * 2))) return rc;
    b->last_chains = a->chains; b->last_nsave = sp.ld_save; b->last_nskip = ns_warm; b->last_var_ns = false;
    sp.qdraws = b->d_qdraws; sp.sp = b->d_sp; sp.chain_stats = b->d_chain_stats;
    sp.draws = ds.draws; sp.ld_draws = ds.ld_draws; sp.udraws = ds.udraws; sp.ld_u = ds.ld_u;
    sp.cpo_acc = ds.cpo; sp.ld_cpo = ds.ld_cpo; sp.inv_metric = ds.inv_metric; sp.ld_im = ds.ld_im; sp.init_u = ds.init_u;
    sp.f_acc = ds.f_acc; sp.ld_f = ds.ld_f;
    // warps per CTA: the largest compiled variant (12, 6, 4, 1) whose shared memory fits on an SM.  A batch that
    // matches a compile-time specialisation (diag_e only) uses it: 12 or 16 warps per SM, with or without the tick barrier
    // (BSYN_NUTS_NW / BSYN_NUTS_TICK override the defaults for A/B measurements).
    size_t bytes = 0;
    int nw = 0, per_sm = 0, tick = 1;
    const VariantOps *ops = nullptr;
    const int want = getenv("BSYN_NUTS_NW") ? atoi(getenv("BSYN_NUTS_NW")) : 0;
    sp.tick_period = std::max(1, getenv("BSYN_TICK_PERIOD") ? atoi(getenv("BSYN_TICK_PERIOD")) : BSYN_DEFAULT_TICK_PERIOD);
    // the specialised variant first (when the batch has one and it has a kernel for this metric), then the fallback
    const VariantOps *cands_ops[2] = {(b->fast && (a->metric == 0 || b->F.robust)) ? b->fast : nullptr, variant(b->cpl)};
    for (const VariantOps *o : cands_ops) {
        if (!o || nw) continue;
        const bool fast = o != variant(b->cpl);
        const int tk = fast ? (getenv("BSYN_NUTS_TICK") ? atoi(getenv("BSYN_NUTS_TICK")) : BSYN_FAST_DEFAULT_TICK) : 1;
        for (int cand : {16, 14, 12, 6, 4, 1}) {
            if (want && cand != want) continue;
            if ((cand == 16 || cand == 14) && !(fast && want)) continue;          // A/B variants of the fixed-grid kernels, on request
            if (fast && cand != 16 && cand != 14 && cand != 12 && cand != 4) continue;
            if (a->metric == 1 && cand != 12 && cand != 4) continue;
            SmemLayout sl = make_layout(b, o, true, &bytes, cand, (b->cpl + 1) * 32);
            if (bytes > 227 * 1024) continue;
            if (o->nuts_occupancy(cand, a->metric, tk, bytes, &per_sm) != cudaSuccess) { cudaGetLastError(); continue; }
            if (per_sm >= 1) { nw = cand; ops = o; tick = tk; sp.sl = sl; break; }
        }
    }
    if (!nw) return fail(BSYN_ERR_UNSUPPORTED, "sampler kernel does not fit on an SM");
    const int n_items = (plan.d_list ? plan.n_list : b->n) * a->chains;
    const int grid = std::max(1, std::min((n_items + nw - 1) / nw, per_sm * b->num_sms));
    const int VS = (b->cpl + 1) * 32;
    sp.arena_stride = (long long)arena_vectors(a->max_treedepth) * VS * (a->metric == 1 ? 2 : 1);
    if ((rc = ensure(&b->d_arena, &b->arena_elems, (size_t)grid * nw * sp.arena_stride))) return rc;
    if (a->metric == 1) {   // per resident warp: M^-1 (float), its Cholesky factor and the Welford accumulator (double)
        const size_t DI = (size_t)VS, per = DI * DI, slots = (size_t)grid * nw;
        sp.dense_stride = (long long)per;
        if ((rc = ensure(&b->d_dense_L, &b->dense_L_elems, slots * per))) return rc;
        if ((rc = ensure(&b->d_dense_m2, &b->dense_m2_elems, slots * per))) return rc;
        if ((rc = ensure(&b->d_dense_mf, &b->dense_mf_elems, (slots * per + 1) / 2))) return rc;
        sp.dense_L = b->d_dense_L; sp.dense_m2 = b->d_dense_m2; sp.dense_minv = reinterpret_cast<float *>(b->d_dense_mf);
    }
    sp.arena = b->d_arena;
    if (plan.resume_in || plan.resume_out) {
        sp.resume_stride = 2LL * VS + 32;
        if (plan.resume_in && b->resume_elems < (size_t)b->n * a->chains * sp.resume_stride) return fail(BSYN_ERR_ARG, "no chain state to resume from");
        if ((rc = ensure(&b->d_resume, &b->resume_elems, (size_t)b->n * a->chains * sp.resume_stride))) return rc;
        sp.resume = b->d_resume;
    }
    sp.grad_counter = b->d_gradctr; sp.work_counter = b->d_workctr;
    CK(cudaMemsetAsync(b->d_gradctr, 0, sizeof(unsigned long long), b->stream));
    CK(cudaMemsetAsync(b->d_workctr, 0, sizeof(int), b->stream));
    if (first) {
        CK(cudaMemsetAsync(b->d_qdraws, 0, sizeof(double) * nrows * BSYN_NQ, b->stream));
        CK(cudaMemsetAsync(b->d_sp, 0, sizeof(double) * nrows * BSYN_NSP, b->stream));
        CK(cudaMemsetAsync(b->d_chain_stats, 0, sizeof(double) * (size_t)b->n * a->chains * 8, b->stream));
    }
    EvPair ev;
    CK(ev.create());
    CK(cudaEventRecord(ev.a, b->stream));
    CK(ops->nuts(nw, tick, grid, bytes, b->stream, b->F, b->d_descs, b->d_dpool, b->d_ipool, b->d_tri, sp));
    ++g_launches;
    // per-problem posterior summaries (mean, sd over all chains and saved sampling draws)
    if (plan.summary) {
        const int g2 = std::max(1, std::min((b->n + WPC - 1) / WPC, b->num_sms * 8));
        summary_kernel<<<g2, 32 * WPC, 0, b->stream>>>(b->d_qdraws, b->n, a->chains, sp.ld_save, ns_warm, sp.n_save - ns_warm, nullptr,
                                                       b->d_summary);
        ++g_launches;
        CK(cudaGetLastError());
    }
    CK(cudaEventRecord(ev.b, b->stream));
    CK(cudaStreamSynchronize(b->stream));
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, ev.a, ev.b));
    if (elapsed_ms) *elapsed_ms = ms;
    unsigned long long ng = 0;
    CK(cudaMemcpy(&ng, b->d_gradctr, sizeof(ng), cudaMemcpyDeviceToHost));
    if (total_grad) *total_grad = (int64_t)ng;
    return BSYN_OK;
}

extern "C" int bsyn_sample_device(bsyn_batch *b, const bsyn_sampler_args *args, int64_t *total_grad_evals, float *elapsed_ms,
                                  const double **d_qdraws, const double **d_sampler_params, const double **d_summary)
{
    if (!b || !args) return fail(BSYN_ERR_ARG, "bsyn_sample_device: null argument");
    DevSinks ds;
    int rc = run_sampler(b, args, ds, total_grad_evals, elapsed_ms);
    if (rc) return rc;
    if (d_qdraws) *d_qdraws = b->d_qdraws;
    if (d_sampler_params) *d_sampler_params = b->d_sp;
    if (d_summary) *d_summary = b->d_summary;
    return BSYN_OK;
}

// Bridge-sampling log marginal likelihood of every problem from the unconstrained draws d_udraws[n][chains][ns][ld_u]
// resident on the device (bsyn_bridge.cuh): proposal fit + point list, ONE batched value-only log_prob launch, iterative scheme.
static int device_bridge(bsyn_batch *b, const double *d_udraws, long long ld_u, int chains, int ns, uint64_t seed, double *logml,
                         int32_t *niter)
{
    if (b->maxD > BRIDGE_MAX_D) return fail(BSYN_ERR_UNSUPPORTED, "bridge sampling on the device supports up to 160 parameters");
    const int n_fit = ns / 2, n_it = ns - n_fit;
    if (n_fit < 8 || n_it < 8) return fail(BSYN_ERR_ARG, "bridge sampling needs at least 16 saved draws per chain");
    if ((long long)chains * n_it > BULK_MAX_DRAWS) return fail(BSYN_ERR_UNSUPPORTED, "bridge sampling: more than 8192 draws in the second halves");
    const size_t npt = (size_t)chains * n_it * 2, n = (size_t)b->n;
    const int ldp = b->maxD;
    DevBuf<double> d_pts, d_lq, d_lp, d_lml;
    DevBuf<int> d_idx, d_st, d_nit;
    if (d_pts.alloc(n * npt * ldp) != cudaSuccess || d_lq.alloc(n * npt) != cudaSuccess || d_lp.alloc(n * npt) != cudaSuccess ||
        d_lml.alloc(n) != cudaSuccess || d_idx.alloc(n * npt) != cudaSuccess || d_st.alloc(n * npt) != cudaSuccess || d_nit.alloc(n) != cudaSuccess) {
        cudaGetLastError();
        return fail(BSYN_ERR_ALLOC, "bridge sampling: device buffers");
    }
    BridgeParams bp{};
    bp.udraws = d_udraws; bp.ld_u = ld_u; bp.chains = chains; bp.ns = ns; bp.seed = seed;
    bp.pts = d_pts; bp.ld_pts = ldp; bp.pidx = d_idx; bp.lq = d_lq; bp.lp = d_lp; bp.lp_status = d_st; bp.logml = d_lml; bp.niter = d_nit;
    bp.maxiter = 1000; bp.tol = 1e-10;
    const size_t sb1 = bridge_prep_smem(b->maxD), sb2 = bridge_iter_smem(chains * n_it, b->maxD);
    CK(cudaFuncSetAttribute(bridge_prep_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sb1));
    CK(cudaFuncSetAttribute(bridge_iter_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sb2));
    const int grid = std::max(1, std::min(b->n, b->num_sms * 2));
    bridge_prep_kernel<<<std::max(1, std::min(b->n, b->num_sms)), BRIDGE_THREADS, sb1, b->stream>>>(b->d_descs, b->n, bp);
    ++g_launches;
    CK(cudaGetLastError());
    int rc = launch_logp(b, (int)(n * npt), d_idx, d_pts, ldp, 1, d_lp, nullptr, d_st, nullptr, 0, b->stream);
    if (rc) return rc;
    bridge_iter_kernel<<<grid, BRIDGE_THREADS, sb2, b->stream>>>(b->d_descs, b->n, bp);
    ++g_launches;
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(logml, d_lml, sizeof(double) * n, cudaMemcpyDeviceToHost, b->stream));
    if (niter) CK(cudaMemcpyAsync(niter, d_nit, sizeof(int) * n, cudaMemcpyDeviceToHost, b->stream));
    CK(cudaStreamSynchronize(b->stream));
    return BSYN_OK;
}

extern "C" int bsyn_sample(bsyn_batch *b, const bsyn_sampler_args *args, const bsyn_sample_sinks *sinks, int64_t *total_grad_evals)
{
    if (!b || !args || !sinks) return fail(BSYN_ERR_ARG, "bsyn_sample: null argument");
    if (int rc0 = check_sampler_args(args)) return rc0;   // before any buffer is sized from them
    if (sinks->cpo_mean && args->iter == args->warmup) return fail(BSYN_ERR_ARG, "bsyn_sample: cpo_mean needs at least one sampling iteration");
    CK(cudaSetDevice(b->device));
    const int ns_samp = (args->iter - args->warmup + args->thin - 1) / std::max(1, args->thin);
    const int ns_warm = args->save_warmup ? (args->warmup + args->thin - 1) / std::max(1, args->thin) : 0;
    const size_t n_save = (size_t)ns_samp + ns_warm;
    const size_t nrows = (size_t)b->n * args->chains * n_save;
    const size_t nch = (size_t)b->n * args->chains;
    DevSinks ds;
    struct Holder {   // the device sinks of this call: freed on every return path
        DevSinks &d;
        ~Holder() { cudaFree(d.draws); cudaFree(d.udraws); cudaFree(d.cpo); cudaFree(d.inv_metric); cudaFree(d.init_u); cudaFree(d.f_acc); }
    } holder{ds};
    auto alloc0 = [&](double **p, size_t n, const char *what) -> int {   // zero-filled device buffer
        if (cudaMalloc(p, sizeof(double) * std::max<size_t>(n, 1)) != cudaSuccess) {
            *p = nullptr;
            cudaGetLastError();
            return fail(BSYN_ERR_ALLOC, std::string("bsyn_sample: the ") + what + " buffer does not fit on the device");
        }
        if (cudaMemsetAsync(*p, 0, sizeof(double) * std::max<size_t>(n, 1), b->stream) != cudaSuccess) return fail(BSYN_ERR_CUDA, "bsyn_sample: memset");
        return BSYN_OK;
    };
    int rc = BSYN_OK;
    int max_ncf = 0;
    for (const ProbDesc &d : b->descs) max_ncf = std::max(max_ncf, d.ncf);
    if ((sinks->f_mean || sinks->surfaces) && sinks->ld_f < max_ncf) return fail(BSYN_ERR_ARG, "bsyn_sample: ld_f < (n1+1)(n2+1)");
    if ((sinks->surfaces || sinks->lpml) && (long long)ns_samp * args->chains > BULK_MAX_DRAWS)
        return fail(BSYN_ERR_UNSUPPORTED, "bsyn_sample: surfaces / lpml need <= 8192 pooled sampling draws per problem");
    if (sinks->draws && sinks->ld_draws < b->maxOut) return fail(BSYN_ERR_ARG, "bsyn_sample: ld_draws < bsyn_max_outputs");
    if (sinks->udraws && sinks->ld_udraws < b->maxD) return fail(BSYN_ERR_ARG, "bsyn_sample: ld_udraws < bsyn_max_params");
    if (sinks->cpo_mean && sinks->ld_cpo < b->maxN) return fail(BSYN_ERR_ARG, "bsyn_sample: ld_cpo < max N");
    if (sinks->draws || sinks->surfaces || sinks->lpml) {
        ds.ld_draws = sinks->draws ? sinks->ld_draws : b->maxOut;
        if ((rc = alloc0(&ds.draws, nrows * ds.ld_draws, "draws (use the summary / surfaces sinks for large screens)"))) return rc;
    }
    if (sinks->bridge_logml && args->save_warmup) return fail(BSYN_ERR_ARG, "bsyn_sample: bridge_logml needs save_warmup = 0");
    if (sinks->udraws || sinks->bridge_logml) {
        ds.ld_u = sinks->udraws ? sinks->ld_udraws : b->maxD;
        if ((rc = alloc0(&ds.udraws, nrows * ds.ld_u, "udraws"))) return rc;
    }
    if (sinks->cpo_mean) {
        ds.ld_cpo = b->maxN;
        if ((rc = alloc0(&ds.cpo, nch * std::max(1, ds.ld_cpo), "cpo"))) return rc;
    }
    if (sinks->inv_metric || sinks->init_u) ds.ld_im = b->maxD;
    if (sinks->inv_metric && (rc = alloc0(&ds.inv_metric, nch * ds.ld_im, "inv_metric"))) return rc;
    if (sinks->init_u && (rc = alloc0(&ds.init_u, nch * ds.ld_im, "init_u"))) return rc;
    if (sinks->f_mean) {
        ds.ld_f = max_ncf;
        if ((rc = alloc0(&ds.f_acc, nch * max_ncf, "f_mean"))) return rc;
    }
    float ms_total = 0.f;
    rc = run_sampler(b, args, ds, total_grad_evals, &ms_total);
    if (rc == BSYN_OK) {
        cudaError_t e = cudaSuccess;
        if (sinks->draws) e = cudaMemcpy(sinks->draws, ds.draws, sizeof(double) * nrows * ds.ld_draws, cudaMemcpyDeviceToHost);
        if (e == cudaSuccess && sinks->udraws) e = cudaMemcpy(sinks->udraws, ds.udraws, sizeof(double) * nrows * ds.ld_u, cudaMemcpyDeviceToHost);
        if (e == cudaSuccess && sinks->qdraws) e = cudaMemcpy(sinks->qdraws, b->d_qdraws, sizeof(double) * nrows * BSYN_NQ, cudaMemcpyDeviceToHost);
        if (e == cudaSuccess && sinks->sampler_params) e = cudaMemcpy(sinks->sampler_params, b->d_sp, sizeof(double) * nrows * BSYN_NSP, cudaMemcpyDeviceToHost);
        if (e == cudaSuccess && sinks->summary) e = cudaMemcpy(sinks->summary, b->d_summary, sizeof(double) * (size_t)b->n * BSYN_NQ * 2, cudaMemcpyDeviceToHost);
        if (e == cudaSuccess && sinks->chain_stats) e = cudaMemcpy(sinks->chain_stats, b->d_chain_stats, sizeof(double) * nch * 8, cudaMemcpyDeviceToHost);
        if (e == cudaSuccess && sinks->inv_metric) e = cudaMemcpy(sinks->inv_metric, ds.inv_metric, sizeof(double) * nch * ds.ld_im, cudaMemcpyDeviceToHost);
        if (e == cudaSuccess && sinks->init_u) e = cudaMemcpy(sinks->init_u, ds.init_u, sizeof(double) * nch * b->maxD, cudaMemcpyDeviceToHost);
        if (sinks->elapsed_ms) *sinks->elapsed_ms = ms_total;
        if (e == cudaSuccess && sinks->f_mean) {   // mean over chains and sampling draws of f per grid cell
            std::vector<double> tmp(nch * (size_t)max_ncf);
            e = cudaMemcpy(tmp.data(), ds.f_acc, sizeof(double) * tmp.size(), cudaMemcpyDeviceToHost);
            if (e == cudaSuccess)
                for (int k = 0; k < b->n; ++k)
                    for (int c = 0; c < b->descs[k].ncf; ++c) {
                        double sacc = 0;
                        for (int ch = 0; ch < args->chains; ++ch) sacc += tmp[((size_t)k * args->chains + ch) * max_ncf + c];
                        sinks->f_mean[(size_t)k * sinks->ld_f + c] = sacc / ((double)args->chains * ns_samp);
                    }
        }
        if (e == cudaSuccess && (sinks->surfaces || sinks->lpml)) {
            // posterior surfaces (mean / median per grid cell) and LPML reduced on the device from the resident draws
            DevBuf<double> d_surf, d_lpml;
            const size_t ldf = sinks->surfaces ? (size_t)sinks->ld_f : (size_t)max_ncf;
            if (d_surf.alloc((size_t)b->n * 3 * ldf) != cudaSuccess || d_lpml.alloc(b->n) != cudaSuccess) e = cudaErrorMemoryAllocation;
            if (e == cudaSuccess) {
                int S2 = 1;
                while (S2 < ns_samp * args->chains) S2 <<= 1;
                const size_t sb = sizeof(double) * ((size_t)S2 + 16);
                e = cudaFuncSetAttribute(surface_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sb);
                if (e == cudaSuccess) {
                    surface_kernel<<<std::max(1, std::min(b->n, b->num_sms * 2)), BULK_THREADS, sb, b->stream>>>(
                        ds.draws, ds.ld_draws, b->d_descs, b->n, args->chains, (int)n_save, ns_warm, ns_samp, b->F.model == 0, b->F.robust,
                        d_surf, (int)ldf, d_lpml);
                    ++g_launches;
                    e = cudaGetLastError();
                }
                if (e == cudaSuccess) e = cudaStreamSynchronize(b->stream);
                if (e == cudaSuccess && sinks->surfaces) e = cudaMemcpy(sinks->surfaces, d_surf, sizeof(double) * (size_t)b->n * 3 * ldf, cudaMemcpyDeviceToHost);
                if (e == cudaSuccess && sinks->lpml) e = cudaMemcpy(sinks->lpml, d_lpml, sizeof(double) * b->n, cudaMemcpyDeviceToHost);
            }
        }
        if (e == cudaSuccess && sinks->cpo_mean) {
            // mean over chains and sampling draws of CPO_n
            std::vector<double> tmp(nch * std::max(1, ds.ld_cpo));
            e = cudaMemcpy(tmp.data(), ds.cpo, sizeof(double) * tmp.size(), cudaMemcpyDeviceToHost);
            if (e == cudaSuccess) {
                for (int k = 0; k < b->n; ++k) {
                    double *o = sinks->cpo_mean + (size_t)k * sinks->ld_cpo;
                    for (int n = 0; n < b->descs[k].N; ++n) {
                        double s = 0;
                        for (int c = 0; c < args->chains; ++c) s += tmp[((size_t)k * args->chains + c) * ds.ld_cpo + n];
                        o[n] = s / ((double)args->chains * ns_samp);
                    }
                }
            }
        }
        if (e != cudaSuccess) rc = fail(BSYN_ERR_CUDA, std::string("copy back: ") + cudaGetErrorString(e));
        if (rc == BSYN_OK && sinks->bridge_logml)
            rc = device_bridge(b, ds.udraws, ds.ld_u, args->chains, (int)n_save, args->seed ^ 0x5bd1e995u, sinks->bridge_logml, sinks->bridge_niter);
    }
    return rc;
}

extern "C" void bsyn_advi_defaults(bsyn_advi_args *a)
{
    a->iter = 10000; a->grad_samples = 1; a->elbo_samples = 100; a->eval_elbo = 100;
    a->eta = 1.0; a->adapt_engaged = 1; a->adapt_iter = 50; a->tol_rel_obj = 0.01;
    a->output_samples = 1000; a->init_r = 2.0; a->seed = 1;
}

extern "C" int bsyn_advi(bsyn_batch *b, const bsyn_advi_args *a, const bsyn_advi_sinks *sinks, int64_t *total_evals,
                         float *elapsed_ms)
{
    if (!b || !a || !sinks) return fail(BSYN_ERR_ARG, "bsyn_advi: null argument");
    if (a->iter < 1 || a->elbo_samples < 1 || a->eval_elbo < 1 || a->adapt_iter < 1 || a->output_samples < 0 ||
        !(a->tol_rel_obj > 0) || !(a->eta > 0))
        return fail(BSYN_ERR_ARG, "bsyn_advi: bad arguments");
    if (a->grad_samples != 1) return fail(BSYN_ERR_UNSUPPORTED, "bsyn_advi: grad_samples must be 1");
    if (sinks->mean && sinks->ld_mean < b->maxD) return fail(BSYN_ERR_ARG, "bsyn_advi: ld_mean < bsyn_max_params");
    const int cb_size = (int)std::max(0.1 * a->iter / a->eval_elbo, 2.0);
    if (cb_size > 32) return fail(BSYN_ERR_UNSUPPORTED, "bsyn_advi: iter / eval_elbo > 320 is not supported");
    CK(cudaSetDevice(b->device));
    AdviParams vp{};
    vp.n_problems = b->n;
    vp.max_iter = a->iter; vp.elbo_samples = a->elbo_samples; vp.eval_elbo = a->eval_elbo;
    vp.adapt_engaged = a->adapt_engaged; vp.adapt_iter = a->adapt_iter; vp.output_samples = a->output_samples;
    vp.cb_size = cb_size; vp.eta = a->eta; vp.tol_rel_obj = a->tol_rel_obj; vp.init_r = a->init_r; vp.seed = a->seed;
    size_t bytes = 0;
    vp.sl = make_layout(b, variant(b->cpl), true, &bytes, ADVI_NW);
    int per_sm = 0;
    if (bytes > 227 * 1024 || variant(b->cpl)->advi_occupancy(bytes, &per_sm) != cudaSuccess || per_sm < 1) {
        cudaGetLastError();
        return fail(BSYN_ERR_UNSUPPORTED, "ADVI kernel does not fit on an SM");
    }
    const int grid = std::max(1, std::min((b->n + ADVI_NW - 1) / ADVI_NW, per_sm * b->num_sms));
    vp.DP = (b->maxD + 31) / 32 * 32;
    vp.state_stride = 6LL * vp.DP + (3LL * vp.DP * vp.DP) / 2;   // 6 vectors + L (double) + its adaGrad history (float)
    const size_t ns = (size_t)a->output_samples, nrows = (size_t)b->n * ns;
    double *d_state = nullptr, *d_mean = nullptr, *d_info = nullptr, *d_draws = nullptr, *d_u = nullptr, *d_cpo = nullptr;
    int rc = BSYN_OK;
    auto cleanup = [&]() { cudaFree(d_state); cudaFree(d_mean); cudaFree(d_info); cudaFree(d_draws); cudaFree(d_u); cudaFree(d_cpo); };
    auto alloc0 = [&](double **p, size_t n) {
        if (cudaMalloc(p, sizeof(double) * std::max<size_t>(n, 1)) != cudaSuccess) return false;
        return cudaMemsetAsync(*p, 0, sizeof(double) * std::max<size_t>(n, 1), b->stream) == cudaSuccess;
    };
    if (!alloc0(&d_state, (size_t)grid * ADVI_NW * vp.state_stride) || !alloc0(&d_mean, (size_t)b->n * b->maxD) ||
        !alloc0(&d_info, (size_t)b->n * BSYN_NVI)) { cleanup(); return fail(BSYN_ERR_ALLOC, "bsyn_advi: device allocation failed"); }
    if (sinks->draws) {
        if (sinks->ld_draws < b->maxOut) { cleanup(); return fail(BSYN_ERR_ARG, "bsyn_advi: ld_draws < bsyn_max_outputs"); }
        if (!alloc0(&d_draws, nrows * sinks->ld_draws)) { cleanup(); return fail(BSYN_ERR_ALLOC, "bsyn_advi: draws buffer does not fit on the device"); }
        vp.draws = d_draws; vp.ld_draws = sinks->ld_draws;
    }
    if (sinks->udraws) {
        if (sinks->ld_udraws < b->maxD) { cleanup(); return fail(BSYN_ERR_ARG, "bsyn_advi: ld_udraws < bsyn_max_params"); }
        if (!alloc0(&d_u, nrows * sinks->ld_udraws)) { cleanup(); return fail(BSYN_ERR_ALLOC, "bsyn_advi: udraws buffer"); }
        vp.udraws = d_u; vp.ld_u = sinks->ld_udraws;
    }
    if (sinks->cpo_mean) {
        if (sinks->ld_cpo < b->maxN) { cleanup(); return fail(BSYN_ERR_ARG, "bsyn_advi: ld_cpo < max N"); }
        if (!alloc0(&d_cpo, (size_t)b->n * std::max(1, b->maxN))) { cleanup(); return fail(BSYN_ERR_ALLOC, "bsyn_advi: cpo buffer"); }
        vp.cpo_acc = d_cpo; vp.ld_cpo = b->maxN;
    }
    if ((rc = ensure(&b->d_qdraws, &b->qdraws_elems, std::max<size_t>(nrows, 1) * BSYN_NQ))) { cleanup(); return rc; }
    if ((rc = ensure(&b->d_summary, &b->summary_elems, (size_t)b->n * BSYN_NQ * 2))) { cleanup(); return rc; }
    b->last_chains = 1; b->last_nsave = a->output_samples; b->last_nskip = 0;
    vp.state = d_state; vp.mean = d_mean; vp.ld_mean = b->maxD; vp.info = d_info; vp.qdraws = b->d_qdraws;
    vp.work_counter = b->d_workctr; vp.grad_counter = b->d_gradctr;
    cudaError_t e = cudaMemsetAsync(b->d_gradctr, 0, sizeof(unsigned long long), b->stream);
    if (e == cudaSuccess) e = cudaMemsetAsync(b->d_workctr, 0, sizeof(int), b->stream);
    if (e == cudaSuccess) e = cudaMemsetAsync(b->d_qdraws, 0, sizeof(double) * std::max<size_t>(nrows, 1) * BSYN_NQ, b->stream);
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    if (e == cudaSuccess) e = cudaEventCreate(&e0);
    if (e == cudaSuccess) e = cudaEventCreate(&e1);
    if (e == cudaSuccess) e = cudaEventRecord(e0, b->stream);
    if (e == cudaSuccess) {
        e = variant(b->cpl)->advi(grid, bytes, b->stream, b->F, b->d_descs, b->d_dpool, b->d_ipool, b->d_tri, vp);
        ++g_launches;
    }
    if (e == cudaSuccess && ns > 0) {
        const int g2 = std::max(1, std::min((b->n + WPC - 1) / WPC, b->num_sms * 8));
        summary_kernel<<<g2, 32 * WPC, 0, b->stream>>>(b->d_qdraws, b->n, 1, a->output_samples, 0, a->output_samples, nullptr, b->d_summary);
        ++g_launches;
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaEventRecord(e1, b->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(b->stream);
    float ms = 0;
    if (e == cudaSuccess) e = cudaEventElapsedTime(&ms, e0, e1);
    if (e0) cudaEventDestroy(e0);
    if (e1) cudaEventDestroy(e1);
    if (elapsed_ms) *elapsed_ms = ms;
    unsigned long long ng = 0;
    if (e == cudaSuccess) e = cudaMemcpy(&ng, b->d_gradctr, sizeof(ng), cudaMemcpyDeviceToHost);
    if (total_evals) *total_evals = (int64_t)ng;
    if (e == cudaSuccess && sinks->mean) {
        e = cudaMemcpy2D(sinks->mean, sizeof(double) * sinks->ld_mean, d_mean, sizeof(double) * b->maxD, sizeof(double) * b->maxD, b->n, cudaMemcpyDeviceToHost);
    }
    if (e == cudaSuccess && sinks->info) e = cudaMemcpy(sinks->info, d_info, sizeof(double) * (size_t)b->n * BSYN_NVI, cudaMemcpyDeviceToHost);
    if (e == cudaSuccess && sinks->draws) e = cudaMemcpy(sinks->draws, d_draws, sizeof(double) * nrows * sinks->ld_draws, cudaMemcpyDeviceToHost);
    if (e == cudaSuccess && sinks->udraws) e = cudaMemcpy(sinks->udraws, d_u, sizeof(double) * nrows * sinks->ld_udraws, cudaMemcpyDeviceToHost);
    if (e == cudaSuccess && sinks->qdraws) e = cudaMemcpy(sinks->qdraws, b->d_qdraws, sizeof(double) * nrows * BSYN_NQ, cudaMemcpyDeviceToHost);
    if (e == cudaSuccess && sinks->summary) e = cudaMemcpy(sinks->summary, b->d_summary, sizeof(double) * (size_t)b->n * BSYN_NQ * 2, cudaMemcpyDeviceToHost);
    if (e == cudaSuccess && sinks->cpo_mean) {
        std::vector<double> tmp((size_t)b->n * std::max(1, b->maxN));
        e = cudaMemcpy(tmp.data(), d_cpo, sizeof(double) * tmp.size(), cudaMemcpyDeviceToHost);
        if (e == cudaSuccess)
            for (int k = 0; k < b->n; ++k)
                for (int n = 0; n < b->descs[k].N; ++n)
                    sinks->cpo_mean[(size_t)k * sinks->ld_cpo + n] = tmp[(size_t)k * b->maxN + n] / std::max<size_t>(ns, 1);
    }
    cleanup();
    if (e != cudaSuccess) return fail(BSYN_ERR_CUDA, std::string("bsyn_advi: ") + cudaGetErrorString(e));
    return BSYN_OK;
}

extern "C" int bsyn_ess(bsyn_batch *b, int32_t q, double *ess, double *rhat)
{
    if (!b || q < 0 || q >= BSYN_NQ || !ess) return fail(BSYN_ERR_ARG, "bsyn_ess: bad argument");
    if (!b->d_qdraws || b->last_nsave - b->last_nskip < 8) return fail(BSYN_ERR_ARG, "bsyn_ess: no draws from a previous bsyn_sample call");
    if (b->last_var_ns) return fail(BSYN_ERR_UNSUPPORTED, "bsyn_ess: after bsyn_sample_to_ess use bsyn_bulk_ess (per-problem draw counts)");
    CK(cudaSetDevice(b->device));
    double *d_out = nullptr;
    CK(cudaMalloc(&d_out, sizeof(double) * 2 * b->n));
    const int grid = std::max(1, std::min((b->n + WPC - 1) / WPC, b->num_sms * 16));
    ess_kernel<<<grid, 32 * WPC, 0, b->stream>>>(b->d_qdraws, b->n, b->last_chains, b->last_nsave, b->last_nskip, q, d_out);
    ++g_launches;
    CK(cudaGetLastError());
    std::vector<double> tmp(2 * (size_t)b->n);
    CK(cudaMemcpyAsync(tmp.data(), d_out, sizeof(double) * tmp.size(), cudaMemcpyDeviceToHost, b->stream));
    CK(cudaStreamSynchronize(b->stream));
    cudaFree(d_out);
    for (int k = 0; k < b->n; ++k) { ess[k] = tmp[2 * k]; if (rhat) rhat[k] = tmp[2 * k + 1]; }
    return BSYN_OK;
}

// Rank-normalised bulk ESS / R-hat and tail ESS of generated scalar q for every problem (bulk_ess_kernel).
extern "C" int bsyn_bulk_ess(bsyn_batch *b, int32_t q, double *ess_bulk, double *rhat, double *ess_tail)
{
    if (!b || q < 0 || q >= BSYN_NQ || !ess_bulk) return fail(BSYN_ERR_ARG, "bsyn_bulk_ess: bad argument");
    const int ns = b->last_nsave - b->last_nskip;
    if (!b->d_qdraws || ns < 8) return fail(BSYN_ERR_ARG, "bsyn_bulk_ess: no draws from a previous bsyn_sample call");
    if ((long long)ns * b->last_chains > BULK_MAX_DRAWS)
        return fail(BSYN_ERR_UNSUPPORTED, "bsyn_bulk_ess: more than 8192 pooled draws per problem (use bsyn_ess)");
    CK(cudaSetDevice(b->device));
    DevBuf<double> d_out;
    CK(d_out.alloc(4 * (size_t)b->n));
    const size_t bytes = bulk_ess_smem_bytes(ns * b->last_chains);
    CK(cudaFuncSetAttribute(bulk_ess_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    const int grid = std::max(1, std::min(b->n, b->num_sms * 4));
    bulk_ess_kernel<<<grid, BULK_THREADS, bytes, b->stream>>>(b->d_qdraws, b->n, nullptr, b->last_chains, b->last_nsave,
                                                              b->last_nskip, ns, b->last_var_ns ? b->d_nsp : nullptr, q, d_out);
    ++g_launches;
    CK(cudaGetLastError());
    std::vector<double> tmp(4 * (size_t)b->n);
    CK(cudaMemcpyAsync(tmp.data(), d_out, sizeof(double) * tmp.size(), cudaMemcpyDeviceToHost, b->stream));
    CK(cudaStreamSynchronize(b->stream));
    for (int k = 0; k < b->n; ++k) {
        ess_bulk[k] = tmp[4 * k];
        if (rhat) rhat[k] = tmp[4 * k + 1];
        if (ess_tail) ess_tail[k] = tmp[4 * k + 2];
    }
    return BSYN_OK;
}

// ---------------------------------------------------------------------------------------------------
// ESS-target control (north_star: "sampled to an ESS of 400 per combination"): one warm-up + sampling launch over
// the whole batch, then extension launches over the problems whose rank-normalised bulk ESS (min over the two scalars
// q1, q2) is still below the target -- their chains resume from the adapted state and append `block` draws.
static __global__ void select_below_kernel(const double *__restrict__ e1, const double *__restrict__ e2, const int n_list,
                                           const int *__restrict__ list_in, const double target, const int ns_now,
                                           double *__restrict__ ess, int *__restrict__ ns_p, int *__restrict__ list_out,
                                           int *__restrict__ n_out)
{
    for (int li = blockIdx.x * blockDim.x + threadIdx.x; li < n_list; li += gridDim.x * blockDim.x) {
        const int k = list_in ? list_in[li] : li;
        const double a = e1[4 * k], c = e2[4 * k];
        const double m = (a == a && c == c) ? fmin(a, c) : nan("");
        ess[k] = m;
        ns_p[k] = ns_now;
        if (!(m >= target)) list_out[atomicAdd(n_out, 1)] = k;
    }
}

extern "C" void bsyn_ess_target_defaults(bsyn_ess_target *t)
{
    t->target_ess = 400.0; t->q1 = BSYN_Q_VUS_SYN; t->q2 = BSYN_Q_VUS_ANT; t->max_rounds = 6;
    t->retry_adapt_delta = 0.99; t->retry_warmup = 150; t->retry_rounds = 6;
}

extern "C" int bsyn_sample_to_ess(bsyn_batch *b, const bsyn_sampler_args *args, const bsyn_ess_target *tg,
                                  const bsyn_sample_sinks *sinks, double *ess, int32_t *draws_per_chain, bsyn_ess_run *run)
{
    if (!b || !args || !tg) return fail(BSYN_ERR_ARG, "bsyn_sample_to_ess: null argument");
    if (int rc0 = check_sampler_args(args)) return rc0;
    const int block = args->iter - args->warmup, q1 = tg->q1, q2 = tg->q2;
    const int max_rounds = tg->max_rounds, retry_rounds = std::max(0, tg->retry_rounds), R = max_rounds + retry_rounds;
    const double target_ess = tg->target_ess;
    if (args->thin != 1 || args->save_warmup || block < 8 || max_rounds < 1 || q1 < 0 || q1 >= BSYN_NQ || q2 < 0 || q2 >= BSYN_NQ ||
        !(target_ess > 0))
        return fail(BSYN_ERR_ARG, "bsyn_sample_to_ess: needs thin = 1, save_warmup = 0, iter - warmup >= 8, max_rounds >= 1");
    if (retry_rounds > 0 && (!(tg->retry_adapt_delta > 0 && tg->retry_adapt_delta < 1) || tg->retry_warmup < 20))
        return fail(BSYN_ERR_ARG, "bsyn_sample_to_ess: retry_adapt_delta must lie in (0,1) and retry_warmup must be >= 20");
    if ((long long)block * R * args->chains > BULK_MAX_DRAWS)
        return fail(BSYN_ERR_UNSUPPORTED, "bsyn_sample_to_ess: chains x (iter - warmup) x (max_rounds + retry_rounds) exceeds 8192 pooled draws");
    if (sinks && (sinks->draws || sinks->udraws || sinks->cpo_mean || sinks->inv_metric || sinks->init_u || sinks->f_mean || sinks->surfaces ||
                  sinks->lpml || sinks->bridge_logml))
        return fail(BSYN_ERR_UNSUPPORTED, "bsyn_sample_to_ess: only the qdraws / sampler_params / summary / chain_stats sinks");
    CK(cudaSetDevice(b->device));
    const int n = b->n, ld_save = block * R;
    DevBuf<double> d_e1, d_e2, d_ess;
    DevBuf<int> d_lists, d_cnt;
    CK(d_e1.alloc(4 * (size_t)n)); CK(d_e2.alloc(4 * (size_t)n)); CK(d_ess.alloc(n));
    CK(d_lists.alloc(2 * (size_t)n)); CK(d_cnt.alloc(1));
    if (!b->d_nsp) CK(cudaMalloc(&b->d_nsp, sizeof(int) * n));
    const size_t ebytes = bulk_ess_smem_bytes(block * R * args->chains);
    CK(cudaFuncSetAttribute(bulk_ess_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ebytes));
    DevSinks ds;
    int64_t tot_grad = 0;
    float tot_ms = 0.f, phase1_ms = 0.f;
    int n_active = n, r = 0, n_retried = 0;
    const int *list = nullptr;       // device list of the active problems (null = all)
    for (;; ++r) {
        bsyn_sampler_args a = *args;
        RunPlan plan;
        plan.ld_save = ld_save; plan.isave0 = r * block; plan.resume_out = (r + 1 < R); plan.summary = false;
        if (r > 0) {
            a.warmup = 0; a.iter = block; a.seed = args->seed + 0x9E3779B97F4A7C15ull * (uint64_t)r;
            plan.resume_in = true; plan.d_list = list; plan.n_list = n_active;
        }
        if (r == max_rounds) {
            // phase 2, the reference's remedy for fits with warnings (R/synergyscreen.R:149-157: adapt_delta = 0.99): the chains of
            // what is still below the target resume, re-adapt their STEP SIZE to the stricter adapt_delta for retry_warmup
            // iterations (init_stepsize + dual averaging; the adapted metric is kept: no metric window), then append blocks
            a.warmup = tg->retry_warmup; a.iter = tg->retry_warmup + block; a.adapt_delta = tg->retry_adapt_delta;
            a.adapt_init_buffer = tg->retry_warmup; a.adapt_window = 0; a.adapt_term_buffer = 0;
            n_retried = n_active;
            phase1_ms = tot_ms;
        } else if (r > max_rounds) {
            a.adapt_delta = tg->retry_adapt_delta;
        }
        int64_t ng = 0;
        float ms = 0.f;
        int rc = run_sampler(b, &a, ds, &ng, &ms, plan);
        if (rc) return rc;
        tot_grad += ng;
        // rank-normalised bulk ESS of q1 and q2 over the (r + 1) * block draws of the active problems
        const int ns_now = (r + 1) * block;
        EvPair ev;
        CK(ev.create());
        CK(cudaEventRecord(ev.a, b->stream));
        const int grid = std::max(1, std::min(n_active, b->num_sms * 4));
        const size_t eb = bulk_ess_smem_bytes(ns_now * args->chains);
        bulk_ess_kernel<<<grid, BULK_THREADS, eb, b->stream>>>(b->d_qdraws, n_active, list, args->chains, ld_save, 0, ns_now, nullptr, q1, d_e1);
        bulk_ess_kernel<<<grid, BULK_THREADS, eb, b->stream>>>(b->d_qdraws, n_active, list, args->chains, ld_save, 0, ns_now, nullptr, q2, d_e2);
        CK(cudaMemsetAsync(d_cnt, 0, sizeof(int), b->stream));
        int *list_out = d_lists.p + (size_t)((r & 1) ? n : 0);
        select_below_kernel<<<std::max(1, std::min((n_active + 255) / 256, 1024)), 256, 0, b->stream>>>(d_e1, d_e2, n_active, list, target_ess, ns_now, d_ess,
                                                                                                       b->d_nsp, list_out, d_cnt);
        g_launches += 3;
        CK(cudaGetLastError());
        CK(cudaEventRecord(ev.b, b->stream));
        int n_next = 0;
        CK(cudaMemcpyAsync(&n_next, d_cnt, sizeof(int), cudaMemcpyDeviceToHost, b->stream));
        CK(cudaStreamSynchronize(b->stream));
        float ms2 = 0.f;
        CK(cudaEventElapsedTime(&ms2, ev.a, ev.b));
        tot_ms += ms + ms2;
        if (n_next == 0 || r + 1 >= R) { ++r; break; }
        list = list_out; n_active = n_next;
    }
    if (r <= max_rounds) phase1_ms = tot_ms;
    b->last_chains = args->chains; b->last_nsave = ld_save; b->last_nskip = 0; b->last_var_ns = true;
    {   // summaries over each problem's own number of draws
        int rc;
        if ((rc = ensure(&b->d_summary, &b->summary_elems, (size_t)n * BSYN_NQ * 2))) return rc;
        const int g2 = std::max(1, std::min((n + WPC - 1) / WPC, b->num_sms * 8));
        summary_kernel<<<g2, 32 * WPC, 0, b->stream>>>(b->d_qdraws, n, args->chains, ld_save, 0, 0, b->d_nsp, b->d_summary);
        ++g_launches;
        CK(cudaGetLastError());
    }
    std::vector<int> nsp(n);
    CK(cudaMemcpyAsync(nsp.data(), b->d_nsp, sizeof(int) * n, cudaMemcpyDeviceToHost, b->stream));
    if (ess) CK(cudaMemcpyAsync(ess, d_ess, sizeof(double) * n, cudaMemcpyDeviceToHost, b->stream));
    if (sinks) {
        const size_t nrows = (size_t)n * args->chains * ld_save, nch = (size_t)n * args->chains;
        if (sinks->qdraws) CK(cudaMemcpyAsync(sinks->qdraws, b->d_qdraws, sizeof(double) * nrows * BSYN_NQ, cudaMemcpyDeviceToHost, b->stream));
        if (sinks->sampler_params) CK(cudaMemcpyAsync(sinks->sampler_params, b->d_sp, sizeof(double) * nrows * BSYN_NSP, cudaMemcpyDeviceToHost, b->stream));
        if (sinks->summary) CK(cudaMemcpyAsync(sinks->summary, b->d_summary, sizeof(double) * (size_t)n * BSYN_NQ * 2, cudaMemcpyDeviceToHost, b->stream));
        if (sinks->chain_stats) CK(cudaMemcpyAsync(sinks->chain_stats, b->d_chain_stats, sizeof(double) * nch * 8, cudaMemcpyDeviceToHost, b->stream));
    }
    CK(cudaStreamSynchronize(b->stream));
    if (draws_per_chain) for (int k = 0; k < n; ++k) draws_per_chain[k] = nsp[k];
    if (run) { run->total_grad_evals = tot_grad; run->elapsed_ms = tot_ms; run->phase1_ms = phase1_ms; run->rounds = r; run->n_retried = n_retried; }
    return BSYN_OK;
}
