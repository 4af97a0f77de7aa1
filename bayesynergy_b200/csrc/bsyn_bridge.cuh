// bsyn_bridge.cuh -- bridge-sampling marginal likelihood on the device.
//
// What it replaces: bridgesampling::bridge_sampler(fit, method = "normal") as bayesynergy() calls it for the Bayes
// factor (R/bayesynergy.R:332-337) [bridgesampling is an un-vendored R dependency; algorithm: Meng & Wong 1996 with the
// multivariate-normal proposal of Gronau et al. 2017].  Per experiment, from the unconstrained posterior draws that the
// sampler left in HBM:
//   bridge_prep_kernel  : proposal moments from the first half of every chain (mean, covariance, Cholesky factor in
//                         shared memory), proposal draws  x = mu + L z  (Philox), the proposal log-density at the
//                         posterior draws of the second halves (forward substitution) and at the proposal draws (-|z|^2/2),
//                         and the [posterior; proposal] point list for ONE batched log_prob launch over all experiments;
//   (bsyn_logp_grad_device, value-only path, on that list)
//   bridge_iter_kernel  : effective sample size of the second halves (median over the D columns of the split-chain ESS, as
//                         bridgesampling does with coda), then the iterative scheme
//                             r <- mean_prop[ l2 / (s1 l2 + s2 r) ] / mean_post[ 1 / (s1 l1 + s2 r) ]
//                         to a relative change of 1e-10.
// One CTA of 256 threads per experiment; D <= BRIDGE_MAX_D (the covariance lives in shared memory).
#pragma once
#include "bsyn_post.cuh"

namespace bsyn {

constexpr int BRIDGE_MAX_D = 160;
constexpr int BRIDGE_THREADS = BULK_THREADS;

struct BridgeParams {
    const double *udraws; long long ld_u;    // [n_problems][chains][ns][ld_u]
    int chains, ns;
    unsigned long long seed;
    double *pts; long long ld_pts;           // [n_problems][n_post + n_prop][ld_pts]: evaluation points, posterior draws first
    int *pidx;                               // [n_problems][n_post + n_prop]: problem index of every point
    double *lq;                              // [n_problems][n_post + n_prop]: proposal log-density at the points
    const double *lp;                        // [n_problems][n_post + n_prop]: model log-density at the points (iter kernel)
    const int *lp_status;
    double *logml; int *niter;               // [n_problems]
    int maxiter; double tol;
};

__device__ __forceinline__ int bridge_n_fit(const int ns) { return ns / 2; }

__global__ void __launch_bounds__(BRIDGE_THREADS) bridge_prep_kernel(const ProbDesc *__restrict__ descs, const int n_problems,
                                                                      const BridgeParams bp)
{
    extern __shared__ double bsm[];
    const int tid = threadIdx.x;
    for (int k = blockIdx.x; k < n_problems; k += gridDim.x) {
        const int D = descs[k].D, ns = bp.ns, chains = bp.chains;
        const int n_fit = bridge_n_fit(ns), n_it = ns - n_fit, n_post = chains * n_it, n_prop = n_post, R = chains * n_fit;
        double *Lm = bsm;                         // [D][D] row-major: covariance, then its lower Cholesky factor
        double *mu = bsm + D * D;                 // [D]
        double *tile = mu + D + (D & 1);          // [16][D] row tile of centred draws
        double *red = tile + 16 * D;              // [8]
        const double *U = bp.udraws + (size_t)k * chains * ns * bp.ld_u;
        auto fit_row = [&](int r) { const int c = r / n_fit, i = r - c * n_fit; return U + ((size_t)c * ns + i) * bp.ld_u; };
        auto post_row = [&](int r) { const int c = r / n_it, i = r - c * n_it; return U + ((size_t)c * ns + n_fit + i) * bp.ld_u; };
        __syncthreads();
        for (int d = tid; d < D; d += BRIDGE_THREADS) {
            double s = 0.0;
            for (int r = 0; r < R; ++r) s += fit_row(r)[d];
            mu[d] = s / R;
        }
        for (int e = tid; e < D * D; e += BRIDGE_THREADS) Lm[e] = 0.0;
        __syncthreads();
        // covariance, lower triangle, accumulated over row tiles staged in shared memory
        for (int r0 = 0; r0 < R; r0 += 16) {
            const int nr = min(16, R - r0);
            for (int e = tid; e < nr * D; e += BRIDGE_THREADS) { const int rr = e / D, d = e - rr * D; tile[rr * D + d] = fit_row(r0 + rr)[d] - mu[d]; }
            __syncthreads();
            for (int e = tid; e < D * D; e += BRIDGE_THREADS) {
                const int i = e / D, j = e - i * D;
                if (j <= i) {
                    double a = 0.0;
                    for (int rr = 0; rr < nr; ++rr) a = fma(tile[rr * D + i], tile[rr * D + j], a);
                    Lm[e] += a;
                }
            }
            __syncthreads();
        }
        for (int e = tid; e < D * D; e += BRIDGE_THREADS) {
            const int i = e / D, j = e - i * D;
            if (j <= i) Lm[e] = Lm[e] / (R - 1.0) + (i == j ? 1e-12 : 0.0);
        }
        __syncthreads();
        // Cholesky, right-looking, in place (lower)
        for (int c = 0; c < D; ++c) {
            const double dg = Lm[c * D + c];
            const double inv = rsqrt(fmax(dg, 1e-300));
            __syncthreads();
            for (int i = c + tid; i < D; i += BRIDGE_THREADS) Lm[i * D + c] *= inv;       // column c (diagonal becomes sqrt)
            __syncthreads();
            for (int e = tid; e < (D - c - 1) * (D - c - 1); e += BRIDGE_THREADS) {
                const int i = c + 1 + e / (D - c - 1), j = c + 1 + e % (D - c - 1);
                if (j <= i) Lm[i * D + j] = fma(-Lm[i * D + c], Lm[j * D + c], Lm[i * D + j]);
            }
            __syncthreads();
        }
        double ld = 0.0;
        for (int d = tid; d < D; d += BRIDGE_THREADS) ld += log(Lm[d * D + d]);
        const double logdet = block_sum(ld, red);                                      // sum log L_ii
        const double cnorm = -logdet - 0.5 * D * 1.8378770664093453;                    // - sum log L_ii - D/2 log(2 pi)
        double *pts = bp.pts + (size_t)k * (n_post + n_prop) * bp.ld_pts;
        double *lq = bp.lq + (size_t)k * (n_post + n_prop);
        int *pidx = bp.pidx + (size_t)k * (n_post + n_prop);
        // posterior draws: copy to the point list; proposal log-density by forward substitution  L y = x - mu
        for (int r = tid; r < n_post; r += BRIDGE_THREADS) {
            const double *x = post_row(r);
            double *o = pts + (size_t)r * bp.ld_pts;
            double y[BRIDGE_MAX_D];
            double q = 0.0;
            for (int i = 0; i < D; ++i) {
                const double xi = x[i];
                o[i] = xi;
                double a = xi - mu[i];
                for (int j = 0; j < i; ++j) a = fma(-Lm[i * D + j], y[j], a);
                y[i] = a / Lm[i * D + i];
                q = fma(y[i], y[i], q);
            }
            for (int i = D; i < bp.ld_pts; ++i) o[i] = 0.0;
            lq[r] = cnorm - 0.5 * q;
            pidx[r] = k;
        }
        // proposal draws  x = mu + L z,  log q(x) = cnorm - |z|^2 / 2
        RngKey rk;
        rk.k0 = (uint32_t)bp.seed; rk.k1 = (uint32_t)(bp.seed >> 32); rk.chain = (uint32_t)k;
        for (int r = tid; r < n_prop; r += BRIDGE_THREADS) {
            double *o = pts + (size_t)(n_post + r) * bp.ld_pts;
            double z[BRIDGE_MAX_D];
            double q = 0.0;
            for (int i = 0; i < D; i += 2) {
                double ua, ub, sn, cs;
                uint32_t ph[4];
                philox4x32((uint32_t)i, (uint32_t)r, 0x42u, rk.chain, rk.k0, rk.k1, ph);
                ua = u01_from_bits(ph[0], ph[1]); ub = u01_from_bits(ph[2], ph[3]);
                sincospi(2.0 * ub, &sn, &cs);
                const double rad = sqrt(-2.0 * log(ua));
                z[i] = rad * cs;
                if (i + 1 < D) z[i + 1] = rad * sn;
            }
            for (int i = 0; i < D; ++i) {
                double a = mu[i];
                for (int j = 0; j <= i; ++j) a = fma(Lm[i * D + j], z[j], a);
                o[i] = a;
                q = fma(z[i], z[i], q);
            }
            for (int i = D; i < bp.ld_pts; ++i) o[i] = 0.0;
            lq[n_post + r] = cnorm - 0.5 * q;
            pidx[n_post + r] = k;
        }
        __syncthreads();
    }
}
inline size_t bridge_prep_smem(int D) { return sizeof(double) * ((size_t)D * D + D + (D & 1) + 16 * (size_t)D + 8); }

__global__ void __launch_bounds__(BRIDGE_THREADS) bridge_iter_kernel(const ProbDesc *__restrict__ descs, const int n_problems,
                                                                      const BridgeParams bp)
{
    extern __shared__ double bsm[];
    const int tid = threadIdx.x;
    for (int k = blockIdx.x; k < n_problems; k += gridDim.x) {
        const int D = descs[k].D, ns = bp.ns, chains = bp.chains;
        const int n_fit = bridge_n_fit(ns), n_it = ns - n_fit, n_post = chains * n_it, n_prop = n_post;
        int S2 = 1;
        while (S2 < n_post) S2 <<= 1;
        double *col = bsm;                       // [chains][n_it] one column of the second halves / sort buffer [S2]
        double *essd = bsm + S2;                 // [D] rounded up to a power of two for the sort
        int D2 = 1;
        while (D2 < D) D2 <<= 1;
        double *red = essd + D2, *mu_s = red + 8;
        const double *U = bp.udraws + (size_t)k * chains * ns * bp.ld_u;
        // neff: median over the columns of the split-chain ESS of the second halves
        for (int d = 0; d < D; ++d) {
            __syncthreads();
            for (int e = tid; e < n_post; e += BRIDGE_THREADS) { const int c = e / n_it, i = e - c * n_it; col[e] = U[((size_t)c * ns + n_fit + i) * bp.ld_u + d]; }
            __syncthreads();
            double ess, rh;
            split_ess_block(col, chains, n_it, red, mu_s, ess, rh);
            if (tid == 0) essd[d] = (ess == ess) ? ess : INFINITY;       // NaN sorts last
        }
        __syncthreads();
        for (int e = D + tid; e < D2; e += BRIDGE_THREADS) essd[e] = INFINITY;
        __syncthreads();
        block_bitonic_sort(essd, D2);
        int nfin = 0;
        for (int d = 0; d < D; ++d) nfin += (essd[d] < INFINITY);
        double neff = (double)n_post;
        if (nfin > 0) neff = (nfin & 1) ? essd[nfin / 2] : 0.5 * (essd[nfin / 2 - 1] + essd[nfin / 2]);
        // l1 = lp - lq at the posterior draws, l2 at the proposal draws; lstar = median(l1)
        const double *lp = bp.lp + (size_t)k * (n_post + n_prop), *lq = bp.lq + (size_t)k * (n_post + n_prop);
        const int *st = bp.lp_status + (size_t)k * (n_post + n_prop);
        __syncthreads();
        for (int e = tid; e < S2; e += BRIDGE_THREADS) col[e] = (e < n_post) ? ((st[e] == 0 ? lp[e] : -INFINITY) - lq[e]) : INFINITY;
        __syncthreads();
        block_bitonic_sort(col, S2);
        const double lstar = (n_post & 1) ? col[n_post / 2] : 0.5 * (col[n_post / 2 - 1] + col[n_post / 2]);
        const double s1 = neff / (neff + n_prop), s2 = n_prop / (neff + n_prop);
        double r = 1.0;
        int it = 0;
        bool bad = false;
        for (it = 1; it <= bp.maxiter; ++it) {
            double num = 0.0, den = 0.0;
            for (int e = tid; e < n_prop; e += BRIDGE_THREADS) {
                const double e2 = exp((st[n_post + e] == 0 ? lp[n_post + e] : -INFINITY) - lq[n_post + e] - lstar);
                num += e2 / (s1 * e2 + s2 * r);
            }
            for (int e = tid; e < n_post; e += BRIDGE_THREADS) {
                const double e1 = exp((st[e] == 0 ? lp[e] : -INFINITY) - lq[e] - lstar);
                den += 1.0 / (s1 * e1 + s2 * r);
            }
            num = block_sum(num, red) / n_prop;
            den = block_sum(den, red) / n_post;
            const double rn = num / den;
            if (!(rn > 0.0) || !(rn < INFINITY)) { bad = true; break; }
            const bool conv = fabs(rn - r) / rn < bp.tol;
            r = rn;
            if (conv) break;
        }
        if (tid == 0) {
            bp.logml[k] = bad ? nan("") : log(r) + lstar;
            bp.niter[k] = it;
        }
        __syncthreads();
    }
}
inline size_t bridge_iter_smem(int n_post, int D)
{
    int S2 = 1, D2 = 1;
    while (S2 < n_post) S2 <<= 1;
    while (D2 < D) D2 <<= 1;
    return sizeof(double) * ((size_t)S2 + D2 + 8 + 64);
}

}  // namespace bsyn
