"""Bridge-sampling estimator (the Bayes-factor row, R/bayesynergy.R:332-337): CPU test on an analytic target,
GPU test of the batched log_prob path against the oracle on the same draws."""
import numpy as np
import pytest

from bayesynergy_b200.bridge import bridge_logml


class _AnalyticBatch:
    """log p~(u) = c - 1/2 (u-m)' P (u-m): log marginal likelihood = c + D/2 log(2 pi) - 1/2 log det P."""

    def __init__(self, D, c, seed=0):
        rng = np.random.default_rng(seed)
        A = rng.normal(size=(D, D))
        self.P = A @ A.T + D * np.eye(D)
        self.m = rng.normal(size=D)
        self.c, self.D = c, D

    def num_params(self, k=0):
        return self.D

    def logp_grad(self, u, prob_idx=None, jacobian=True, want_grad=True):
        d = np.asarray(u)[:, :self.D] - self.m
        lp = self.c - 0.5 * np.einsum("ni,ij,nj->n", d, self.P, d)
        return lp, None, np.zeros(len(lp), dtype=np.int32)

    def truth(self):
        return self.c + 0.5 * self.D * np.log(2 * np.pi) - 0.5 * np.linalg.slogdet(self.P)[1]

    def draws(self, chains, n, seed=1):
        rng = np.random.default_rng(seed)
        L = np.linalg.cholesky(np.linalg.inv(self.P))
        return self.m + rng.standard_normal((chains, n, self.D)) @ L.T


def test_bridge_logml_analytic():
    b = _AnalyticBatch(D=6, c=-3.7)
    est, it = bridge_logml(b, b.draws(4, 1000))
    assert abs(est - b.truth()) < 0.02 and it < 100


@pytest.mark.gpu
def test_bridge_gpu_matches_oracle(built, datasets):
    from bayesynergy_b200 import _lib
    from bayesynergy_b200.dataprep import prepare
    from oracle.c_oracle import Problem
    y, x = datasets["mathews:ispinesib + ibrutinib"]
    sd = prepare(y, x)
    for model in ("gp_grid", "nointeraction"):
        b = _lib.Batch([sd], model=model)
        res = b.sample(chains=4, iter=600, warmup=300, seed=5, want_udraws=True, want_qdraws=False)
        ud = res["udraws"][0]
        est, _ = bridge_logml(b, ud, seed=3)

        class _OracleBatch:
            P = Problem(sd, model)

            def num_params(self, k=0):
                return self.P.dim

            def logp_grad(self, u, prob_idx=None, jacobian=True, want_grad=True):
                st, lp, _ = self.P.logp_grad_batch(np.ascontiguousarray(np.asarray(u)[:, :self.P.dim]), jacobian=jacobian)
                return lp, None, st

        ref, _ = bridge_logml(_OracleBatch(), ud, seed=3)
        assert np.isfinite(est) and abs(est - ref) < 1e-7 * max(1.0, abs(ref)), (model, est, ref)
        b.close()


@pytest.mark.gpu
def test_bayesynergy_api_with_bayes_factor(built, datasets):
    """bayesynergy() end to end on the vignette's example (vignettes/Example_single.Rmd): result fields, LPML, BF."""
    from bayesynergy_b200.api import bayesynergy
    y, x = datasets["mathews:ispinesib + ibrutinib"]
    fit = bayesynergy(y, x, drug_names=["ispinesib", "ibrutinib"], bayes_factor=True, iter=1000, seed=7)
    pm = fit["posterior_mean"]
    assert 75 < pm["rVUS_f"] < 90 and 65 < pm["rVUS_p0"] < 80 and -16 < pm["VUS_syn"] < -5
    assert pm["f"].shape == (6, 6) and pm["f"][0, 0] == 1.0 and np.all(pm["p0"] <= 1.0)
    assert np.isfinite(fit["LPML"]) and fit["returnCode"] in (0, 1)
    assert fit["posterior"]["Delta"].shape == (2000, 5, 5) and fit["posterior"]["CPO"].shape == (2000, 36)
    assert fit["bayesfactor"] > 1.0          # strong interaction in this experiment (the vignette reports synergy)


def test_bulk_ess_rank_normalised():
    """Host-side rank-normalised bulk ESS: ~N for independent draws whatever the tails, and close to the classic
    estimator for a Gaussian AR(1) chain."""
    from bayesynergy_b200.diagnostics import bulk_ess_rhat, split_ess_rhat
    rng = np.random.default_rng(5)
    x = rng.standard_cauchy(size=(4, 1000))
    ess, rhat = bulk_ess_rhat(x)
    assert 3000 < ess < 5200 and abs(rhat - 1.0) < 0.01
    phi = 0.8
    e = rng.standard_normal((4, 2000))
    y = np.zeros_like(e)
    for t in range(1, 2000):
        y[:, t] = phi * y[:, t - 1] + e[:, t]
    eb, _ = bulk_ess_rhat(y)
    ec, _ = split_ess_rhat(y)
    assert abs(eb - ec) < 0.15 * ec and 500 < ec < 1400      # theory: 8000 (1 - phi) / (1 + phi) = 889
