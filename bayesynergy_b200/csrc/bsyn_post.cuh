// bsyn_post.cuh -- on-device post-processing of the generated scalars kept per draw:
// posterior mean / sd per combination (what one screenSummary row needs, R/synergyscreen.R:185-218) and
// the split-chain effective sample size / R-hat (Stan's compute_split_effective_sample_size /
// split potential scale reduction: Geyer initial positive + monotone sequence) [StanHeaders, un-vendored].
#pragma once
#include "bsyn_common.cuh"

namespace bsyn {

constexpr int NQ = 12;

// one warp per problem; qdraws[(k*chains + c)*ld_save + i][NQ]; the first n_skip saved draws of every chain
// (saved warmup) are skipped, the next ns (= ns_p[k] when ns_p != nullptr, problems sampled for different lengths,
// else ns_all) are used.  summary[k][NQ][2] = mean, sd (n-1 denominator over all chains pooled).
__global__ void summary_kernel(const double *__restrict__ qdraws, const int n_problems, const int chains,
                               const int ld_save, const int n_skip, const int ns_all, const int *__restrict__ ns_p,
                               double *__restrict__ summary)
{
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    const int nwarps = (gridDim.x * blockDim.x) >> 5;
    for (int k = warp; k < n_problems; k += nwarps) {
        const int ns = ns_p ? ns_p[k] : ns_all;
        const long long tot = (long long)chains * ns;
        const double *base = qdraws + (size_t)k * chains * ld_save * NQ;
        for (int q = 0; q < NQ; ++q) {
            double s = 0.0;
            for (long long r = lane; r < tot; r += 32) {
                const int c = (int)(r / ns), i = (int)(r - (long long)c * ns) + n_skip;
                s += base[((size_t)c * ld_save + i) * NQ + q];
            }
            const double mean = warp_sum(s) / (double)tot;
            double v = 0.0;
            for (long long r = lane; r < tot; r += 32) {
                const int c = (int)(r / ns), i = (int)(r - (long long)c * ns) + n_skip;
                const double d = base[((size_t)c * ld_save + i) * NQ + q] - mean;
                v = fma(d, d, v);
            }
            v = warp_sum(v);
            if (lane == 0) {
                summary[((size_t)k * NQ + q) * 2] = mean;
                summary[((size_t)k * NQ + q) * 2 + 1] = tot > 1 ? sqrt(v / (double)(tot - 1)) : 0.0;
            }
        }
    }
}

// Split-chain ESS (and split R-hat) of generated scalar q, one warp per problem.  out[k][2] = ess, rhat.
__global__ void ess_kernel(const double *__restrict__ qdraws, const int n_problems, const int chains, const int n_save,
                           const int n_skip, const int q, double *__restrict__ out)
{
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    const int nwarps = (gridDim.x * blockDim.x) >> 5;
    const int ns = n_save - n_skip;
    const int n = ns / 2;                 // draws per split chain
    const int second = ns - n;            // start of the second half (ceil(ns/2))
    const int m = 2 * chains;
    constexpr int MAXM = 32;
    for (int k = warp; k < n_problems; k += nwarps) {
        if (n < 4 || m > MAXM) { if (lane == 0) { out[2 * k] = nan(""); out[2 * k + 1] = nan(""); } continue; }
        const double *base = qdraws + (size_t)k * chains * n_save * NQ + q;
        // split chain j: chain c = j/2, offset (j&1) ? second : 0
        double mean_j = 0.0;   // lane j (< m) keeps the mean of split chain j
        double var_j = 0.0;    // ... and its biased variance acov(0)
        for (int j = 0; j < m; ++j) {
            const double *x = base + ((size_t)(j >> 1) * n_save + n_skip + ((j & 1) ? second : 0)) * NQ;
            double s = 0.0;
            for (int i = lane; i < n; i += 32) s += x[(size_t)i * NQ];
            const double mu = warp_sum(s) / n;
            double v = 0.0;
            for (int i = lane; i < n; i += 32) { const double d = x[(size_t)i * NQ] - mu; v = fma(d, d, v); }
            v = warp_sum(v) / n;
            if (lane == j) { mean_j = mu; var_j = v; }
        }
        // mean_var = mean of unbiased chain variances; var_plus = mean_var*(n-1)/n + var(chain means)
        const double cv = (lane < m) ? var_j * n / (n - 1.0) : 0.0;
        const double mean_var = warp_sum(cv) / m;
        const double mm = warp_sum(lane < m ? mean_j : 0.0) / m;
        const double dm = (lane < m) ? mean_j - mm : 0.0;
        const double var_means = warp_sum(dm * dm) / (m - 1.0);
        const double var_plus = mean_var * (n - 1.0) / n + var_means;
        const double rhat = sqrt(var_plus / mean_var);
        if (!(var_plus > 0.0) || !(mean_var > 0.0)) { if (lane == 0) { out[2 * k] = nan(""); out[2 * k + 1] = nan(""); } continue; }
        // mean over split chains of the biased autocovariance at lag t
        auto mean_acov = [&](int t) -> double {
            double tot = 0.0;
            for (int j = 0; j < m; ++j) {
                const double *x = base + ((size_t)(j >> 1) * n_save + n_skip + ((j & 1) ? second : 0)) * NQ;
                const double mu = __shfl_sync(FULL, mean_j, j);
                double s = 0.0;
                for (int i = lane; i + t < n; i += 32) s = fma(x[(size_t)i * NQ] - mu, x[(size_t)(i + t) * NQ] - mu, s);
                tot += s;
            }
            return warp_sum(tot) / ((double)n * m);
        };
        double rho_even = 1.0;
        double rho_odd = 1.0 - (mean_var - mean_acov(1)) / var_plus;
        // pairs: pair 0 = (1, rho_1); acc = sum of monotone-adjusted pair sums of all pairs before the last one
        double acc = 0.0, prev_sum = 0.0, pend_sum = rho_even + rho_odd, pend_even = rho_even;
        int npair = 1;   // pairs computed so far (the pending one is pair npair-1)
        int s = 1;
        while (s < n - 4 && (rho_even + rho_odd) > 0.0) {
            // the pending pair is not the last: finalise it (monotone w.r.t. its predecessor)
            double ps = pend_sum;
            if (npair >= 2 && ps > prev_sum) ps = prev_sum;
            acc += ps; prev_sum = ps;
            rho_even = 1.0 - (mean_var - mean_acov(s + 1)) / var_plus;
            rho_odd = 1.0 - (mean_var - mean_acov(s + 2)) / var_plus;
            if (rho_even + rho_odd >= 0.0) { pend_sum = rho_even + rho_odd; pend_even = rho_even; }
            else { pend_sum = 0.0; pend_even = 0.0; }
            ++npair;
            s += 2;
        }
        // head(max_s) = all finalised pairs + the even element of the last (unadjusted) pair
        double tau = -1.0 + 2.0 * (acc + pend_even) + (rho_even > 0.0 ? rho_even : 0.0);
        const double ntot = (double)n * m;
        double ess = ntot / tau;
        const double cap = ntot * log10(ntot);
        if (!(ess < cap)) ess = cap;
        if (lane == 0) { out[2 * k] = ess; out[2 * k + 1] = rhat; }
    }
}

}  // namespace bsyn

namespace bsyn {

// ---------------------------------------------------------------------------------------------------
// Rank-normalised split-chain bulk ESS / R-hat and tail ESS (Vehtari, Gelman, Simpson, Carpenter & Buerkner 2021;
// what rstan >= 2.21 prints as Bulk_ESS / Tail_ESS and what SURVEY.md §8(d) defines the ESS metric on).
// One CTA per problem: the pooled draws of generated scalar q are sorted in shared memory (bitonic, value + position),
// replaced by the normal scores of their average ranks, z = Phi^-1((r - 3/8) / (S + 1/4)), and the split-chain
// estimator of ess_kernel (Geyer's initial positive + monotone sequence) is applied to z.  Tail ESS = the smaller of
// the split-chain ESS of the indicators I(x <= q05) and I(x <= q95), quantiles from the same sort.
// qdraws[(k*chains + c)*ld_save + i][NQ]; draws n_skip .. n_skip + ns - 1 of every chain are used, ns = ns_p[k] when
// ns_p != nullptr (problems sampled for different lengths), else ns_all.  prob_list (may be null) selects problems.
// out[k][4] = bulk ess, bulk rhat, tail ess, draws used.
constexpr int BULK_THREADS = 256;
constexpr int BULK_MAX_DRAWS = 8192;

__device__ __forceinline__ double block_sum(double v, double *red)
{
    v = warp_sum(v);
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    __syncthreads();
    if (l == 0) red[w] = v;
    __syncthreads();
    double t = (l < BULK_THREADS / 32) ? red[l] : 0.0;
    t = warp_sum(t);
    return t;
}

// split-chain ESS (and R-hat) of z[chains][ns] in shared memory, every thread of the CTA returns the same values
__device__ void split_ess_block(const double *z, const int chains, const int ns, double *red, double *mu_s, double &ess_out,
                                double &rhat_out)
{
    const int n = ns / 2, second = ns - n, m = 2 * chains;
    const double qnan = __longlong_as_double(0x7ff8000000000000LL);
    ess_out = qnan; rhat_out = qnan;
    if (n < 4 || m > 64) return;
    double mean_var = 0.0, mm = 0.0;
    for (int j = 0; j < m; ++j) {
        const double *x = z + (size_t)(j >> 1) * ns + ((j & 1) ? second : 0);
        double s = 0.0;
        for (int i = threadIdx.x; i < n; i += BULK_THREADS) s += x[i];
        const double mu = block_sum(s, red) / n;
        double v = 0.0;
        for (int i = threadIdx.x; i < n; i += BULK_THREADS) { const double d = x[i] - mu; v = fma(d, d, v); }
        v = block_sum(v, red) / n;
        __syncthreads();
        if (threadIdx.x == 0) mu_s[j] = mu;
        mean_var += v * n / (n - 1.0);
        mm += mu;
    }
    __syncthreads();
    mean_var /= m; mm /= m;
    double var_means = 0.0;
    for (int j = 0; j < m; ++j) { const double d = mu_s[j] - mm; var_means = fma(d, d, var_means); }
    var_means /= (m - 1.0);
    const double var_plus = mean_var * (n - 1.0) / n + var_means;
    if (!(var_plus > 0.0) || !(mean_var > 0.0)) return;
    rhat_out = sqrt(var_plus / mean_var);
    auto mean_acov = [&](int t) -> double {
        double tot = 0.0;
        for (int j = 0; j < m; ++j) {
            const double *x = z + (size_t)(j >> 1) * ns + ((j & 1) ? second : 0);
            const double mu = mu_s[j];
            for (int i = threadIdx.x; i + t < n; i += BULK_THREADS) tot = fma(x[i] - mu, x[i + t] - mu, tot);
        }
        return block_sum(tot, red) / ((double)n * m);
    };
    double rho_even = 1.0;
    double rho_odd = 1.0 - (mean_var - mean_acov(1)) / var_plus;
    double acc = 0.0, prev_sum = 0.0, pend_sum = rho_even + rho_odd, pend_even = rho_even;
    int npair = 1, s = 1;
    while (s < n - 4 && (rho_even + rho_odd) > 0.0) {
        double ps = pend_sum;
        if (npair >= 2 && ps > prev_sum) ps = prev_sum;
        acc += ps; prev_sum = ps;
        rho_even = 1.0 - (mean_var - mean_acov(s + 1)) / var_plus;
        rho_odd = 1.0 - (mean_var - mean_acov(s + 2)) / var_plus;
        if (rho_even + rho_odd >= 0.0) { pend_sum = rho_even + rho_odd; pend_even = rho_even; }
        else { pend_sum = 0.0; pend_even = 0.0; }
        ++npair;
        s += 2;
    }
    const double tau = -1.0 + 2.0 * (acc + pend_even) + (rho_even > 0.0 ? rho_even : 0.0);
    const double ntot = (double)n * m;
    double ess = ntot / tau;
    const double cap = ntot * log10(ntot);
    if (!(ess < cap)) ess = cap;
    ess_out = ess;
}

__global__ void __launch_bounds__(BULK_THREADS) bulk_ess_kernel(const double *__restrict__ qdraws, const int n_list,
                                                                 const int *__restrict__ prob_list, const int chains,
                                                                 const int ld_save, const int n_skip, const int ns_all,
                                                                 const int *__restrict__ ns_p, const int q,
                                                                 double *__restrict__ out)
{
    extern __shared__ double bsm[];
    for (int li = blockIdx.x; li < n_list; li += gridDim.x) {
        const int k = prob_list ? prob_list[li] : li;
        const int ns = ns_p ? ns_p[k] : ns_all;
        const int S = chains * ns;
        int S2 = 1;
        while (S2 < S) S2 <<= 1;
        // layout: key[S2] | z[S] | idx[S2] (ints) | red[8] | mu[64]
        double *key = bsm, *z = bsm + S2;
        int *idx = reinterpret_cast<int *>(z + S + (S & 1));
        double *red = reinterpret_cast<double *>(idx + S2), *mu_s = red + 8;
        __syncthreads();
        if (S < 16 || S > BULK_MAX_DRAWS) {
            if (threadIdx.x == 0) { const double qn = nan(""); out[4 * k] = qn; out[4 * k + 1] = qn; out[4 * k + 2] = qn; out[4 * k + 3] = (double)S; }
            continue;
        }
        const double *base = qdraws + (size_t)k * chains * ld_save * NQ + q;
        for (int e = threadIdx.x; e < S2; e += BULK_THREADS) {
            double v = INFINITY;
            if (e < S) { const int c = e / ns, i = e - c * ns; v = base[((size_t)c * ld_save + n_skip + i) * NQ]; }
            key[e] = v; idx[e] = e;
        }
        __syncthreads();
        for (int kk = 2; kk <= S2; kk <<= 1) {
            for (int j = kk >> 1; j > 0; j >>= 1) {
                for (int e = threadIdx.x; e < S2; e += BULK_THREADS) {
                    const int p = e ^ j;
                    if (p > e) {
                        const bool up = (e & kk) == 0;
                        const double a = key[e], b = key[p];
                        const bool sw = up ? (a > b || (a == b && idx[e] > idx[p])) : (a < b || (a == b && idx[e] < idx[p]));
                        if (sw) { key[e] = b; key[p] = a; const int t = idx[e]; idx[e] = idx[p]; idx[p] = t; }
                    }
                }
                __syncthreads();
            }
        }
        // tail quantiles (type-7-free: order statistics at floor(0.05 S), floor(0.95 S))
        const double q05 = key[(int)(0.05 * S)], q95 = key[min(S - 1, (int)(0.95 * S))];
        // average ranks of ties -> normal scores, scattered back to [chain][draw] order
        for (int e = threadIdx.x; e < S; e += BULK_THREADS) {
            const double v = key[e];
            int lo = e, hi = e;
            while (lo > 0 && key[lo - 1] == v) --lo;
            while (hi + 1 < S && key[hi + 1] == v) ++hi;
            const double r = 0.5 * (lo + hi) + 1.0;
            z[idx[e]] = normcdfinv((r - 0.375) / (S + 0.25));
        }
        __syncthreads();
        double ess_b, rhat_b, ess_lo, ess_hi, dummy;
        split_ess_block(z, chains, ns, red, mu_s, ess_b, rhat_b);
        __syncthreads();
        for (int e = threadIdx.x; e < S; e += BULK_THREADS) { const int c = e / ns, i = e - c * ns; z[e] = (base[((size_t)c * ld_save + n_skip + i) * NQ] <= q05) ? 1.0 : 0.0; }
        __syncthreads();
        split_ess_block(z, chains, ns, red, mu_s, ess_lo, dummy);
        __syncthreads();
        for (int e = threadIdx.x; e < S; e += BULK_THREADS) { const int c = e / ns, i = e - c * ns; z[e] = (base[((size_t)c * ld_save + n_skip + i) * NQ] <= q95) ? 1.0 : 0.0; }
        __syncthreads();
        split_ess_block(z, chains, ns, red, mu_s, ess_hi, dummy);
        if (threadIdx.x == 0) {
            out[4 * k] = ess_b; out[4 * k + 1] = rhat_b;
            out[4 * k + 2] = (ess_lo == ess_lo && ess_hi == ess_hi) ? fmin(ess_lo, ess_hi) : nan("");
            out[4 * k + 3] = (double)S;
        }
    }
}
inline size_t bulk_ess_smem_bytes(int S)
{
    int S2 = 1;
    while (S2 < S) S2 <<= 1;
    return sizeof(double) * ((size_t)S2 + S + (S & 1) + 8 + 64) + sizeof(int) * (size_t)S2;
}

}  // namespace bsyn

namespace bsyn {

// ---------------------------------------------------------------------------------------------------
// Posterior surfaces and LPML from the full write_array draws (what bayesynergy() computes after rstan::extract,
// R/bayesynergy.R:244-291): per cell of the (n2+1) x (n1+1) dose grid
//   f     = mean over the draws (median when `robust`) of the response surface  [1 at (0,0), p02 / p01 on the edges,
//           p0 + Delta inside]
//   p0    = median of the non-interaction surface, Delta = median of the interaction surface (0 on the edges)
// and LPML = -sum_n log(mean over draws of CPO_n).  One CTA per problem; every median is a shared-memory bitonic sort of
// the chains x ns values of one cell (R's median: the mean of the two middle order statistics for an even count).
// draws[(k*chains + c)*ld_save + i][ld_draws]; out[k][3][ld_f] column-major cells fc = b + a*(n2+1); lpml[k].
__device__ __forceinline__ void block_bitonic_sort(double *key, const int S2)
{
    for (int kk = 2; kk <= S2; kk <<= 1)
        for (int j = kk >> 1; j > 0; j >>= 1) {
            for (int e = threadIdx.x; e < S2; e += BULK_THREADS) {
                const int p = e ^ j;
                if (p > e) {
                    const bool up = (e & kk) == 0;
                    const double a = key[e], b = key[p];
                    if (up ? (a > b) : (a < b)) { key[e] = b; key[p] = a; }
                }
            }
            __syncthreads();
        }
}

__global__ void __launch_bounds__(BULK_THREADS) surface_kernel(const double *__restrict__ draws, const long long ld_draws,
                                                                const ProbDesc *__restrict__ descs, const int n_problems,
                                                                const int chains, const int ld_save, const int n_skip,
                                                                const int ns, const int gp, const int robust,
                                                                double *__restrict__ out, const int ld_f,
                                                                double *__restrict__ lpml)
{
    extern __shared__ double bsm[];
    const int S = chains * ns;
    int S2 = 1;
    while (S2 < S) S2 <<= 1;
    double *key = bsm, *red = bsm + S2;
    for (int k = blockIdx.x; k < n_problems; k += gridDim.x) {
        const ProbDesc pd = descs[k];
        const int n1 = pd.n1, n2 = pd.n2, G = pd.G, D = pd.D, nm = n1 + n2;
        const int o_p0 = D, o_pm = D + G, o_delta = D + G + nm, o_cpo = (gp ? o_delta + 2 * G : o_delta) + 2;
        const double *base = draws + (size_t)k * chains * ld_save * ld_draws;
        auto rowp = [&](int e) { const int c = e / ns, i = e - c * ns; return base + ((size_t)c * ld_save + n_skip + i) * ld_draws; };
        for (int fc = 0; fc < pd.ncf; ++fc) {
            const int a = fc / (n2 + 1), bb = fc - a * (n2 + 1);      // a: drug-1 level (column), bb: drug-2 level (row)
            const bool inner = a > 0 && bb > 0;
            const int cell = inner ? (bb - 1) + (a - 1) * n2 : 0;
            for (int sfc = 0; sfc < 3; ++sfc) {
                double *dst = out + ((size_t)k * 3 + sfc) * ld_f + fc;
                if (sfc == 2 && (!inner || !gp)) { if (threadIdx.x == 0) *dst = 0.0; continue; }
                if (fc == 0 && sfc < 2) { if (threadIdx.x == 0) *dst = 1.0; continue; }
                __syncthreads();
                double sum = 0.0;
                for (int e = threadIdx.x; e < S2; e += BULK_THREADS) {
                    double v = INFINITY;
                    if (e < S) {
                        const double *r = rowp(e);
                        if (!inner) v = (a > 0) ? r[o_pm + a - 1] : r[o_pm + n1 + bb - 1];
                        else if (sfc == 0) v = r[o_p0 + cell] + (gp ? r[o_delta + cell] : 0.0);
                        else if (sfc == 1) v = r[o_p0 + cell];
                        else v = r[o_delta + cell];
                        sum += v;
                    }
                    key[e] = v;
                }
                if (sfc == 0 && !robust) {
                    const double tot = block_sum(sum, red);
                    if (threadIdx.x == 0) *dst = tot / S;
                    continue;
                }
                __syncthreads();
                block_bitonic_sort(key, S2);
                if (threadIdx.x == 0) *dst = (S & 1) ? key[S / 2] : 0.5 * (key[S / 2 - 1] + key[S / 2]);
            }
        }
        // LPML
        double acc = 0.0;
        for (int n = 0; n < pd.N; ++n) {
            double s = 0.0;
            for (int e = threadIdx.x; e < S; e += BULK_THREADS) s += rowp(e)[o_cpo + n];
            const double tot = block_sum(s, red);
            acc += log(tot / S);
        }
        if (threadIdx.x == 0) lpml[k] = -acc;
    }
}

}  // namespace bsyn
