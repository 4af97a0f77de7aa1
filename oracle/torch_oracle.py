"""ORACLE (test infrastructure, not product code) -- PARITY UNPINNED.

torch-float64 restatement of the reference's `gp_grid` / `nointeraction` log-density,
statement by statement from the Stan programs, with the gradient taken by autograd
(so it is independent of the hand-derived adjoints in oracle/gp_oracle.c and in the
CUDA kernels).  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg
may import this module.

"Parity unpinned": the reference ships no tests, golden vectors or fixtures for this
path and rstan/Stan Math are un-vendored dependencies
(DESCRIPTION:30-44: rstan >= 2.18.1, StanHeaders >= 2.18.0; generated code stamped
Stan 2.21.0, src/stanExports_gp_grid.h:21) that cannot be built or imported here, so
this restatement follows the published Stan Math formulas and
is anchored only on the reference's own call sites.

Follows:
  inst/stan/gp_grid.stan:64-233          (parameters, transformed parameters, model)
  src/stanExports_gp_grid.h:1086-1215    (parameter read order + constraining transforms)
  src/stanExports_gp_grid.h:1497-1568    (which lpdfs are called; all keep their constants)
  inst/stan/nointeraction.stan:33-139
  inst/stan/include/{kron_mvprod,pc_prior,lptn}.stan
"""
import math

import numpy as np
import torch

LOG_2PI_HALF = 0.5 * math.log(2 * math.pi)
LN10 = math.log(10.0)


def _lub01(u, lp, jacobian):
    """stan::math::lub_constrain(u, 0, 1, lp)  [ext, Stan Math 2.21 lub_constrain.hpp]."""
    if u.item() > 0:
        e = torch.exp(-u)
        th = 1.0 / (1.0 + e)
        if jacobian:
            lp = lp - u - 2.0 * torch.log1p(e)
        if th.item() == 1.0:
            th = th * 0 + (1.0 - 1e-15)
    else:
        e = torch.exp(u)
        th = 1.0 - 1.0 / (1.0 + e)
        if jacobian:
            lp = lp + u - 2.0 * torch.log1p(e)
        if th.item() == 0.0:
            th = th * 0 + 1e-15
    return th, lp


def _lb0(u, lp, jacobian):
    """stan::math::lb_constrain(u, 0, lp): lp += u; return exp(u)."""
    if jacobian:
        lp = lp + u
    return torch.exp(u), lp


def _normal_lpdf(y, mu, sigma):
    return -LOG_2PI_HALF - torch.log(sigma) - 0.5 * ((y - mu) / sigma) ** 2


def _std_normal_lpdf(y):
    return -0.5 * y * y - LOG_2PI_HALF


def _inv_gamma_lpdf(y, a, b):
    return a * math.log(b) - math.lgamma(a) - (a + 1) * torch.log(y) - b / y


def _lptn(z, lam, tau):
    """inst/stan/include/lptn.stan:1-14."""
    if abs(z.item()) <= tau:
        return _std_normal_lpdf(z)
    logt = math.log(tau)
    logabsz = torch.log(torch.clamp(torch.abs(z), min=1.0001))
    return (-0.5 * tau * tau - LOG_2PI_HALF) + logt - logabsz + (lam + 1) * (math.log(logt) - torch.log(logabsz))


def lptn_constants(rho):
    """gp_grid.stan:37-38: tauLPTN = inv_Phi((1+rho)/2); lambdaLPTN = 2/(1-rho)*phi(tau)*tau*log(tau)."""
    from scipy.special import ndtri
    tau = float(ndtri((1 + rho) / 2))
    phi = math.exp(-0.5 * tau * tau - LOG_2PI_HALF)
    lam = 2.0 / (1.0 - rho) * phi * tau * math.log(tau)
    return tau, lam


def _kernel_mats(sd, ell, sigma_f, alpha):
    """gp_grid.stan:112-141."""
    x1 = torch.as_tensor(sd.x1, dtype=torch.float64)
    x2 = torch.as_tensor(sd.x2, dtype=torch.float64)
    d1 = torch.sqrt((x1[:, None] - x1[None, :]) ** 2)
    d2 = torch.sqrt((x2[:, None] - x2[None, :]) ** 2)
    d1s, d2s = d1 ** 2, d2 ** 2
    I1 = 1e-10 * torch.eye(sd.n1, dtype=torch.float64)
    I2 = 1e-10 * torch.eye(sd.n2, dtype=torch.float64)
    if sd.kernel == 1:
        cov1 = sigma_f * torch.exp(-d1s / (2 * ell ** 2)) + I1
        cov2 = torch.exp(-d2s / (2 * ell ** 2)) + I2
    elif sd.kernel == 2:
        if sd.nu_matern == 1:
            cov1 = sigma_f * torch.exp(-d1 / ell) + I1
            cov2 = torch.exp(-d2 / ell) + I2
        elif sd.nu_matern == 2:
            s3 = math.sqrt(3.0)
            cov1 = sigma_f * ((1 + s3 * (d1 / ell)) * torch.exp(-s3 * d1 / ell)) + I1
            cov2 = (1 + s3 * (d2 / ell)) * torch.exp(-s3 * d2 / ell) + I2
        elif sd.nu_matern == 3:
            s5 = math.sqrt(5.0)
            cov1 = sigma_f * ((1 + s5 * (d1 / ell) + (5.0 / 3.0) * (d1s / ell ** 2)) * torch.exp(-s5 * d1 / ell)) + I1
            cov2 = (1 + s5 * (d2 / ell) + (5.0 / 3.0) * (d2s / ell ** 2)) * torch.exp(-s5 * d2 / ell) + I2
        else:
            raise ValueError("nu_matern must be 1, 2 or 3")
    elif sd.kernel == 3:
        cov1 = sigma_f * torch.exp(-alpha * torch.log(1 + d1s / (2 * alpha * ell ** 2))) + I1
        cov2 = torch.exp(-alpha * torch.log(1 + d2s / (2 * alpha * ell ** 2))) + I2
    else:
        raise ValueError("kernel must be 1, 2 or 3")
    return cov1, cov2


class DomainError(Exception):
    """The reference throws std::domain_error (rejected proposal) -- h:1396-1457, cholesky_decompose."""


def log_prob(sd, u, model="gp_grid", jacobian=True, return_parts=False):
    """lp(u) as a torch scalar with autograd graph.  u: 1-D float64 tensor (requires_grad ok)."""
    n1, n2 = sd.n1, sd.n2
    lp = torch.zeros((), dtype=torch.float64)
    k = 0

    def nxt():
        nonlocal k
        v = u[k]
        k += 1
        return v

    la_1 = la_2 = None
    if sd.est_la:
        la_1, lp = _lub01(nxt(), lp, jacobian)
        la_2, lp = _lub01(nxt(), lp, jacobian)
    m1, m2, th1, th2 = nxt(), nxt(), nxt(), nxt()
    slope_1, lp = _lb0(nxt(), lp, jacobian)
    slope_2, lp = _lb0(nxt(), lp, jacobian)
    gp = model == "gp_grid"
    if gp:
        b1, lp = _lb0(nxt(), lp, jacobian)
        b2, lp = _lb0(nxt(), lp, jacobian)
        ell, lp = _lb0(nxt(), lp, jacobian)
        sigma_f, lp = _lb0(nxt(), lp, jacobian)
        alpha = None
        if sd.est_alpha:
            alpha, lp = _lb0(nxt(), lp, jacobian)
        z = u[k:k + n1 * n2].reshape(n1, n2).T   # column-major n2 x n1
        k += n1 * n2
    s, lp = _lb0(nxt(), lp, jacobian)
    s2_1, lp = _lb0(nxt(), lp, jacobian)
    s2_2, lp = _lb0(nxt(), lp, jacobian)
    assert k == u.shape[0], (k, u.shape)

    x1 = torch.as_tensor(sd.x1, dtype=torch.float64)
    x2 = torch.as_tensor(sd.x2, dtype=torch.float64)
    zero = torch.zeros((), dtype=torch.float64)
    la1p = la_1 if sd.est_la else zero
    la2p = la_2 if sd.est_la else zero

    if gp:
        cov1, cov2 = _kernel_mats(sd, ell, sigma_f, alpha)
        try:
            L1 = torch.linalg.cholesky(cov1)
            L2 = torch.linalg.cholesky(cov2)
        except Exception as e:  # not positive definite
            raise DomainError(str(e))
        GP = (L1 @ (L2 @ z).T).T               # include/kron_mvprod.stan:4
    p01 = la1p + (1 - la1p) / (1 + 10.0 ** (slope_1 * (x1 - m1)))
    p02 = la2p + (1 - la2p) / (1 + 10.0 ** (slope_2 * (x2 - m2)))
    p0 = p02[:, None] * p01[None, :]
    if gp:
        lg = torch.log(p0 / (1 - p0))
        Delta = -p0 / (1 + torch.exp(b1 * GP + lg)) + (1 - p0) / (1 + torch.exp(-b2 * GP - lg))
    # validation (h:1396-1457)
    chk = [p0, p01, p02] + ([Delta] if gp else [])
    for t in chk:
        if torch.isnan(t).any():
            raise DomainError("nan in transformed parameter")
    if (p0 < 0).any() or (p0 > 1).any() or (p01 < 0).any() or (p01 > 1).any() or (p02 < 0).any() or (p02 > 1).any():
        raise DomainError("p0/p01/p02 out of [0,1]")
    if gp and ((Delta < -1).any() or (Delta > 1).any()):
        raise DomainError("Delta out of [-1,1]")

    f = torch.empty((n2 + 1, n1 + 1), dtype=torch.float64)
    f = f.clone()
    rows = [torch.cat([torch.ones(1, dtype=torch.float64), p01])]
    interior = p0 + Delta if gp else p0
    body = torch.cat([p02[:, None], interior], dim=1)
    f = torch.cat([rows[0][None, :], body], dim=0)

    parts = {}
    jac = lp
    lp = torch.zeros((), dtype=torch.float64)
    # priors (gp_grid.stan:177-209)
    lp = lp + (-math.log(math.pi) - torch.log1p(s * s)) - math.log(0.5)
    lp = lp + _inv_gamma_lpdf(s2_1, 3.0, 2.0) + _inv_gamma_lpdf(s2_2, 3.0, 2.0)
    if sd.est_la:
        cb = math.lgamma(2.25) - math.lgamma(1.0) - math.lgamma(1.25)
        lp = lp + cb + 0.25 * torch.log1p(-la_1)
        lp = lp + cb + 0.25 * torch.log1p(-la_2)
    lp = lp - slope_1 - slope_2
    lp = lp + _std_normal_lpdf(th1) + _std_normal_lpdf(th2)
    lp = lp + _normal_lpdf(m1, th1, torch.sqrt(s2_1)) + _normal_lpdf(m2, th2, torch.sqrt(s2_2))
    if gp:
        lp = lp + _normal_lpdf(b1, 1.0, torch.tensor(0.1, dtype=torch.float64))
        lp = lp + _normal_lpdf(b2, 1.0, torch.tensor(0.1, dtype=torch.float64))
        if sd.pcprior:
            h = sd.pcprior_hypers
            lam1 = -math.log(h[1]) * h[0] ** 1.0
            lam2 = -math.log(h[3]) / h[2]
            lp = lp + (math.log(1.0) + math.log(lam1) + math.log(lam2) + (-2.0) * torch.log(ell)
                       - lam1 / ell - lam2 * torch.sqrt(sigma_f) - torch.log(torch.sqrt(sigma_f)))
        else:
            lp = lp + _inv_gamma_lpdf(ell, 5.0, 5.0)
            lp = lp + (-LOG_2PI_HALF - torch.log(sigma_f) - 0.5 * (torch.log(sigma_f) - 1.0) ** 2)
        if sd.est_alpha:
            lp = lp - alpha
        lp = lp + _std_normal_lpdf(z).sum()
    parts["prior"] = lp
    # likelihood (gp_grid.stan:212-232)
    ii = torch.as_tensor(np.asarray(sd.ii_obs, dtype=np.int64) - 1)
    fvec = f.T.reshape(-1)                       # to_vector: column-major
    fobs = fvec[ii]
    y = torch.as_tensor(sd.y, dtype=torch.float64)
    noise = s * torch.sqrt(fobs + sd.lambda_)
    lik = torch.zeros((), dtype=torch.float64)
    if sd.robust:
        tau, lam = lptn_constants(sd.rho)
        for n in range(y.shape[0]):
            sdn = noise[n] if sd.heteroscedastic else s
            lik = lik + _lptn((y[n] - fobs[n]) / sdn, lam, tau) - torch.log(sdn)
    else:
        if sd.heteroscedastic:
            lik = _normal_lpdf(y, fobs, noise).sum()
        else:
            lik = _normal_lpdf(y, fobs, s).sum()
    if torch.isnan(noise).any() and sd.heteroscedastic:
        raise DomainError("scale parameter is nan")
    total = lp + lik + jac
    if return_parts:
        parts["lik"] = lik
        parts["jac"] = jac
        return total, parts
    return total


def log_prob_grad(sd, u, model="gp_grid", jacobian=True):
    """(lp, grad) as (float, np.ndarray).  Raises DomainError where the reference would throw."""
    ut = torch.tensor(np.asarray(u, dtype=np.float64), requires_grad=True)
    lp = log_prob(sd, ut, model=model, jacobian=jacobian)
    (g,) = torch.autograd.grad(lp, ut)
    return float(lp.item()), g.numpy().copy()
