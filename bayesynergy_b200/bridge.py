"""Bridge-sampling marginal likelihood and the Bayes factor of gp_grid vs nointeraction.

Mirrors what R/bayesynergy.R:332-337 does with `bridgesampling::bridge_sampler(fit)` /
`bridge_sampler(fitH0)` / `bf()` [bridgesampling is an un-vendored dependency; algorithm: Meng & Wong 1996 with
the multivariate-normal proposal of Gronau et al. 2017, "method = normal"]: the thousands of
`$log_prob(upars, TRUE, FALSE)` calls the R package makes through the module (src/stanExports_gp_grid.cc:24)
become ONE batched `bsyn_logp_grad` launch per model.
"""
from __future__ import annotations

import numpy as np

from . import _lib
from .diagnostics import split_ess_rhat


def _log_mvn(x, mu, L):
    """log N(x | mu, L L^T) for rows of x."""
    d = x.shape[1]
    zs = np.linalg.solve(L, (x - mu).T)          # [d, n]
    return -0.5 * (zs * zs).sum(axis=0) - np.log(np.diag(L)).sum() - 0.5 * d * np.log(2 * np.pi)


def _ess_columns(x):
    """Split-chain ESS (Geyer, as diagnostics.split_ess_rhat) of every column of x[chains, n, D] at once: FFT
    autocovariances, the initial-positive / monotone-sequence truncation vectorised over the columns."""
    chains, ns, D = x.shape
    n = ns // 2
    second = ns - n
    sp = np.concatenate([np.stack([x[c, :n], x[c, second:second + n]]) for c in range(chains)])     # [m, n, D]
    m = sp.shape[0]
    mu = sp.mean(axis=1, keepdims=True)
    d = sp - mu
    nfft = 1 << int(np.ceil(np.log2(2 * n)))
    f = np.fft.rfft(d, n=nfft, axis=1)
    acov = np.fft.irfft(f * np.conj(f), n=nfft, axis=1)[:, :n, :] / n                                # [m, n, D] biased
    mean_acov = acov.mean(axis=0)                                                                    # [n, D]
    mean_var = mean_acov[0] * n / (n - 1.0)
    var_plus = mean_var * (n - 1.0) / n + mu[:, 0, :].var(axis=0, ddof=1)
    with np.errstate(divide="ignore", invalid="ignore"):
        rho = 1.0 - (mean_var - mean_acov) / var_plus                                                # [n, D]
    rho[0] = 1.0
    npair = (n - 3) // 2                                    # pairs (rho[2p], rho[2p+1]) the sequential loop can reach
    pe, po = rho[0:2 * npair:2], rho[1:2 * npair:2]
    ps = pe + po
    neg = ps <= 0                                           # the loop stops at the first non-positive pair sum
    first_neg = np.where(neg.any(axis=0), neg.argmax(axis=0), npair)                                 # [D]
    idx = np.arange(npair)[:, None]
    psm = np.where(idx < first_neg[None, :], ps, 0.0)
    psm = np.minimum.accumulate(np.where(idx < first_neg[None, :], psm, np.inf), axis=0)             # monotone
    psm = np.where(idx < first_neg[None, :], psm, 0.0)
    tau = -1.0 + 2.0 * psm.sum(axis=0)
    # the estimator of diagnostics.split_ess_rhat keeps the even element of the first non-positive pair when it is > 0
    last_even = np.where(first_neg < npair, pe[np.minimum(first_neg, npair - 1), np.arange(D)], 0.0)
    tau = tau + np.where(last_even > 0, last_even, 0.0)
    ntot = float(n * m)
    with np.errstate(divide="ignore", invalid="ignore"):
        ess = np.where(tau > 0, ntot / tau, np.nan)
    return np.minimum(ess, ntot * np.log10(ntot))


def _proposal(u, seed):
    """Moments of the first half of every chain, proposal draws, and both sample sets of one problem."""
    chains, n, D = u.shape
    n_fit = n // 2
    fit = u[:, :n_fit].reshape(-1, D)            # first half of every chain: proposal moments
    it = u[:, n_fit:]                            # second half: the iterative scheme
    post = it.reshape(-1, D)
    mu = fit.mean(axis=0)
    cov = np.cov(fit, rowvar=False) + 1e-12 * np.eye(D)
    L = np.linalg.cholesky(cov)
    rng = np.random.default_rng(seed)
    prop = mu + rng.standard_normal((post.shape[0], D)) @ L.T
    return it, post, prop, mu, L


def _iterate(lp_post, lp_prop, post, prop, mu, L, neff, maxiter=1000, tol=1e-10):
    """Meng & Wong's iterative scheme (bridgesampling `method = "normal"`).  Returns (logml, niter)."""
    n_prop = prop.shape[0]
    l1 = lp_post - _log_mvn(post, mu, L)         # posterior samples
    l2 = lp_prop - _log_mvn(prop, mu, L)         # proposal samples
    s1 = neff / (neff + n_prop)
    s2 = n_prop / (neff + n_prop)
    lstar = np.median(l1)
    e1 = np.exp(l1 - lstar)
    e2 = np.exp(l2 - lstar)
    r = 1.0
    for k in range(1, maxiter + 1):
        num = np.mean(e2 / (s1 * e2 + s2 * r))
        den = np.mean(1.0 / (s1 * e1 + s2 * r))
        r_new = num / den
        if not np.isfinite(r_new) or r_new <= 0:
            return float("nan"), k
        if abs(r_new - r) / r_new < tol:
            r = r_new
            break
        r = r_new
    return float(np.log(r) + lstar), k


def bridge_logml(batch: _lib.Batch, udraws, problem=0, seed=1, maxiter=1000, tol=1e-10):
    """log marginal likelihood from unconstrained posterior draws udraws[chains, n, >=D] of `problem`.

    Returns (logml, niter).  log_prob is evaluated with jacobian = TRUE at the posterior and proposal samples.
    """
    out, nit = bridge_logml_many(batch, [udraws], problems=[problem], seed=seed, maxiter=maxiter, tol=tol)
    return out[0], nit[0]


def bridge_logml_many(batch: _lib.Batch, udraws_list, problems=None, seed=1, maxiter=1000, tol=1e-10):
    """log marginal likelihoods of many experiments of one batch: every `log_prob` call of every experiment (posterior
    and proposal samples alike) goes into ONE batched `bsyn_logp_grad` launch; the proposal fit and the iterative scheme
    run vectorised on the host per experiment.  udraws_list[k]: [chains, n, >= D_k].  Returns (logml[n], niter[n])."""
    problems = list(range(len(udraws_list))) if problems is None else list(problems)
    preps, pts, idx = [], [], []
    ld = batch.max_params
    for k, (u, prob) in enumerate(zip(udraws_list, problems)):
        D = batch.num_params(prob)
        u = np.asarray(u, dtype=np.float64)[:, :, :D]
        it, post, prop, mu, L = _proposal(u, seed + 7919 * k)
        ess = _ess_columns(it)
        neff = float(np.nanmedian(ess)) if np.isfinite(np.nanmedian(ess)) else float(post.shape[0])
        preps.append((post, prop, mu, L, neff))
        both = np.zeros((post.shape[0] + prop.shape[0], ld))
        both[:post.shape[0], :D] = post
        both[post.shape[0]:, :D] = prop
        pts.append(both)
        idx.append(np.full(both.shape[0], prob, dtype=np.int32))
    lp, _, st = batch.logp_grad(np.concatenate(pts), prob_idx=np.concatenate(idx), jacobian=True, want_grad=False)
    lp = np.where(st == 0, lp, -np.inf)
    out, nit, o = [], [], 0
    for post, prop, mu, L, neff in preps:
        n1, n2 = post.shape[0], prop.shape[0]
        a, k = _iterate(lp[o:o + n1], lp[o + n1:o + n1 + n2], post, prop, mu, L, neff, maxiter, tol)
        o += n1 + n2
        out.append(a)
        nit.append(k)
    return np.array(out), np.array(nit)


def bayes_factor_many(sds, udraws_gp, sampler_kw, device=0, seed=1, chunk=256):
    """Bayes factors of the interaction model over the no-interaction model for a list of experiments
    (R/bayesynergy.R:221-223, 332-337, per experiment; here batched): ONE nointeraction sampler launch over all
    experiments, then ONE batched log_prob launch per model for all bridge points (chunked to bound host memory).
    udraws_gp[k]: [chains, n, D] unconstrained gp_grid draws of experiment k.  Returns bf[n]."""
    n = len(sds)
    kw = dict(sampler_kw)
    b0 = _lib.Batch(sds, model="nointeraction", device=device)
    try:
        res0 = b0.sample(want_udraws=True, want_qdraws=False, want_sampler_params=False, **kw)
        lml0 = np.concatenate([bridge_logml_many(b0, [res0["udraws"][k] for k in range(lo, min(n, lo + chunk))],
                                                 problems=range(lo, min(n, lo + chunk)), seed=seed + 1)[0]
                               for lo in range(0, n, chunk)])
    finally:
        b0.close()
    b1 = _lib.Batch(sds, model="gp_grid", device=device)
    try:
        lml1 = np.concatenate([bridge_logml_many(b1, [udraws_gp[k] for k in range(lo, min(n, lo + chunk))],
                                                 problems=range(lo, min(n, lo + chunk)), seed=seed)[0]
                               for lo in range(0, n, chunk)])
    finally:
        b1.close()
    return np.exp(lml1 - lml0)


DEVICE_BRIDGE_MAX_D = 160      # BRIDGE_MAX_D of csrc/bsyn_bridge.cuh: the proposal covariance lives in shared memory


def bayes_factor_from_logml(sds, logml_gp, sampler_kw, device=0):
    """Bayes factors when the gp_grid launch already produced its bridge-sampling log marginal likelihoods on the device
    (bsyn_sample's `bridge_logml` sink): ONE nointeraction launch with the same sink, nothing but 2 n doubles comes back.
    Returns bf[n]."""
    b0 = _lib.Batch(sds, model="nointeraction", device=device)
    try:
        res0 = b0.sample(want_bridge=True, want_qdraws=False, want_sampler_params=False, **sampler_kw)
    finally:
        b0.close()
    return np.exp(np.asarray(logml_gp) - res0["bridge_logml"])


def bayes_factor(sd, udraws_full, sampler_kw, device=0, seed=1):
    """BF of the interaction model over the no-interaction model for one experiment (R/bayesynergy.R:332-337)."""
    return float(bayes_factor_many([sd], [udraws_full], sampler_kw, device=device, seed=seed)[0])
