"""Host-side chain diagnostics: split-chain effective sample size and R-hat.

Same estimator as the device kernel (csrc/bsyn_post.cuh: ess_kernel) so GPU and CPU-baseline runs are scored
with identical arithmetic: Stan's compute_split_effective_sample_size (Geyer initial positive + monotone
sequence on the split chains; what rstan::summary reports as n_eff) [StanHeaders, un-vendored].
"""
import numpy as np


def split_ess_rhat(x):
    """x: [chains, draws].  Returns (ess, rhat)."""
    x = np.asarray(x, dtype=np.float64)
    chains, ns = x.shape
    n = ns // 2
    second = ns - n
    if n < 4:
        return float("nan"), float("nan")
    sp = np.concatenate([np.stack([x[c, :n], x[c, second:second + n]]) for c in range(chains)])   # [2*chains, n]
    m = sp.shape[0]
    mu = sp.mean(axis=1)
    d = sp - mu[:, None]
    acov0 = (d * d).sum(axis=1) / n
    mean_var = (acov0 * n / (n - 1.0)).mean()
    var_plus = mean_var * (n - 1.0) / n + mu.var(ddof=1)
    if not (var_plus > 0 and mean_var > 0):
        return float("nan"), float("nan")
    rhat = np.sqrt(var_plus / mean_var)

    def mean_acov(t):
        return (d[:, :n - t] * d[:, t:]).sum() / (n * m)

    rho_even = 1.0
    rho_odd = 1.0 - (mean_var - mean_acov(1)) / var_plus
    acc, prev_sum, pend_sum, pend_even, npair, s = 0.0, 0.0, rho_even + rho_odd, rho_even, 1, 1
    while s < n - 4 and (rho_even + rho_odd) > 0:
        ps = pend_sum
        if npair >= 2 and ps > prev_sum:
            ps = prev_sum
        acc += ps
        prev_sum = ps
        rho_even = 1.0 - (mean_var - mean_acov(s + 1)) / var_plus
        rho_odd = 1.0 - (mean_var - mean_acov(s + 2)) / var_plus
        if rho_even + rho_odd >= 0:
            pend_sum, pend_even = rho_even + rho_odd, rho_even
        else:
            pend_sum, pend_even = 0.0, 0.0
        npair += 1
        s += 2
    tau = -1.0 + 2.0 * (acc + pend_even) + (rho_even if rho_even > 0 else 0.0)
    ntot = float(n * m)
    ess = min(ntot / tau, ntot * np.log10(ntot)) if tau > 0 else ntot * np.log10(ntot)
    return float(ess), float(rhat)


def mcse_mean(x):
    """Monte Carlo standard error of the posterior mean of x[chains, draws]."""
    ess, _ = split_ess_rhat(x)
    return float(np.std(x, ddof=1) / np.sqrt(max(ess, 1.0)))


def bulk_ess_rhat(x):
    """Rank-normalised split-chain bulk-ESS and R-hat (Vehtari, Gelman, Simpson, Carpenter & Buerkner 2021; what newer
    rstan prints as Bulk_ESS): the pooled draws are replaced by normal scores of their fractional ranks, then the
    split-chain estimator above is applied.  x: [chains, draws]."""
    from scipy.stats import norm, rankdata
    x = np.asarray(x, dtype=np.float64)
    r = rankdata(x.reshape(-1), method="average").reshape(x.shape)
    z = norm.ppf((r - 0.375) / (x.size + 0.25))
    return split_ess_rhat(z)


def bulk_ess(x):
    return bulk_ess_rhat(x)[0]


def tail_ess(x):
    """Tail ESS (Vehtari et al. 2021): min over the split-chain ESS of I(x <= q05) and I(x <= q95); the quantiles are
    the order statistics the device kernel uses (bsyn_post.cuh: bulk_ess_kernel)."""
    x = np.asarray(x, dtype=np.float64)
    s = np.sort(x.reshape(-1))
    S = s.size
    q05, q95 = s[int(0.05 * S)], s[min(S - 1, int(0.95 * S))]
    e1, _ = split_ess_rhat((x <= q05).astype(np.float64))
    e2, _ = split_ess_rhat((x <= q95).astype(np.float64))
    return float(min(e1, e2)) if e1 == e1 and e2 == e2 else float("nan")
