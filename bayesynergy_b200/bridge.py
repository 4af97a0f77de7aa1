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


def bridge_logml(batch: _lib.Batch, udraws, problem=0, seed=1, maxiter=1000, tol=1e-10):
    """log marginal likelihood from unconstrained posterior draws udraws[chains, n, >=D] of `problem`.

    Returns (logml, niter).  log_prob is evaluated with jacobian = TRUE at the posterior and proposal samples.
    """
    D = batch.num_params(problem)
    u = np.asarray(udraws, dtype=np.float64)[:, :, :D]
    chains, n, _ = u.shape
    n_fit = n // 2
    fit = u[:, :n_fit].reshape(-1, D)            # first half of every chain: proposal moments
    it = u[:, n_fit:]                            # second half: the iterative scheme
    post = it.reshape(-1, D)
    mu = fit.mean(axis=0)
    cov = np.cov(fit, rowvar=False) + 1e-12 * np.eye(D)
    L = np.linalg.cholesky(cov)
    rng = np.random.default_rng(seed)
    n_prop = post.shape[0]
    prop = mu + rng.standard_normal((n_prop, D)) @ L.T
    idx = np.full(post.shape[0] + n_prop, problem, dtype=np.int32)
    lp, _, st = batch.logp_grad(np.concatenate([post, prop]), prob_idx=idx, jacobian=True, want_grad=False)
    lp = np.where(st == 0, lp, -np.inf)
    l1 = lp[:post.shape[0]] - _log_mvn(post, mu, L)      # posterior samples
    l2 = lp[post.shape[0]:] - _log_mvn(prop, mu, L)      # proposal samples
    # effective number of posterior samples (median over dimensions), as bridgesampling does with coda
    ess = np.nanmedian([split_ess_rhat(it[:, :, d])[0] for d in range(D)])
    neff = float(ess) if np.isfinite(ess) else float(post.shape[0])
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


def bayes_factor(sd, udraws_full, sampler_kw, device=0, seed=1):
    """BF of the interaction model over the no-interaction model for one experiment (R/bayesynergy.R:332-337)."""
    b1 = _lib.Batch([sd], model="gp_grid", device=device)
    try:
        logml1, _ = bridge_logml(b1, udraws_full, seed=seed)
    finally:
        b1.close()
    b0 = _lib.Batch([sd], model="nointeraction", device=device)
    try:
        res0 = b0.sample(want_udraws=True, want_qdraws=False, want_sampler_params=False, **sampler_kw)
        logml0, _ = bridge_logml(b0, res0["udraws"][0], seed=seed + 1)
    finally:
        b0.close()
    return float(np.exp(logml1 - logml0))
