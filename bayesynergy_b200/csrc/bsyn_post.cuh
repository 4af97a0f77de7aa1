// bsyn_post.cuh -- on-device post-processing of the generated scalars kept per draw:
// posterior mean / sd per combination (what one screenSummary row needs, R/synergyscreen.R:185-218) and
// the split-chain effective sample size / R-hat (Stan's compute_split_effective_sample_size /
// split potential scale reduction: Geyer initial positive + monotone sequence) [StanHeaders, un-vendored].
#pragma once
#include "bsyn_common.cuh"

namespace bsyn {

constexpr int NQ = 12;

// one warp per problem; qdraws[(k*chains + c)*n_save + i][NQ]; the first n_skip saved draws of every chain
// (saved warmup) are skipped.  summary[k][NQ][2] = mean, sd (n-1 denominator over all chains pooled).
__global__ void summary_kernel(const double *__restrict__ qdraws, const int n_problems, const int chains,
                               const int n_save, const int n_skip, double *__restrict__ summary)
{
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    const int nwarps = (gridDim.x * blockDim.x) >> 5;
    const int ns = n_save - n_skip;
    const long long tot = (long long)chains * ns;
    for (int k = warp; k < n_problems; k += nwarps) {
        const double *base = qdraws + (size_t)k * chains * n_save * NQ;
        for (int q = 0; q < NQ; ++q) {
            double s = 0.0;
            for (long long r = lane; r < tot; r += 32) {
                const int c = (int)(r / ns), i = (int)(r - (long long)c * ns) + n_skip;
                s += base[((size_t)c * n_save + i) * NQ + q];
            }
            const double mean = warp_sum(s) / (double)tot;
            double v = 0.0;
            for (long long r = lane; r < tot; r += 32) {
                const int c = (int)(r / ns), i = (int)(r - (long long)c * ns) + n_skip;
                const double d = base[((size_t)c * n_save + i) * NQ + q] - mean;
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
