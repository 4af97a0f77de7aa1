"""Bridge-sampling estimator (the Bayes-factor row, R/bayesynergy.R:332-337): CPU test on an analytic target,
GPU test of the batched log_prob path against the oracle on the same draws."""
import numpy as np
import pytest

from bayesynergy_b200.bridge import bridge_logml


class _AnalyticBatch:
    """log p~(u) = c - 1/2 (u-m)' P (u-m): log marginal likelihood = c + D/2 log(2 pi) - 1/2 log det P."""

    def __init__(self, D, c, seed=0):
        self.max_params = D
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
            max_params = P.dim

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
    # the stanfit-shaped view: rstan::extract / summary / get_sampler_params / get_divergent_iterations
    sf = fit["stanfit"]
    ex = sf.extract()
    assert ex["z"].shape == (2000, 5, 5) and ex["la_1"].shape == (2000, 1) and ex["alpha"].shape == (2000, 0)
    assert np.array_equal(ex["VUS_syn"], fit["posterior"]["VUS_syn"]) and np.array_equal(ex["Delta"], fit["posterior"]["Delta"])
    assert np.array_equal(ex["lp__"], fit["posterior"]["lp__"])
    sm = sf.summary(pars=["VUS_syn", "s", "lp__"])
    assert sm["rownames"] == ["s", "VUS_syn", "lp__"] and sm["colnames"][:3] == ["mean", "se_mean", "sd"]
    assert abs(sm["summary"][1, 0] - pm["VUS_syn"]) < 1e-12 and (sm["summary"][:, -1] < 1.05).all()
    assert (sm["summary"][:, -2] > 100).all()                          # n_eff
    sp = sf.get_sampler_params()
    assert len(sp) == 4 and sp[0]["n_leapfrog__"].shape == (500,) and (sp[0]["stepsize__"] > 0).all()
    assert sf.get_divergent_iterations().shape == (2000,)
    assert "# Step size = " in sf.get_adaptation_info()[0] and sf.get_elapsed_time().shape == (4, 2)
    assert sf.get_inits().shape == (4, 40) and np.abs(sf.get_inits()).max() <= 2.0


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


@pytest.mark.gpu
def test_device_surfaces_and_lpml_match_host(built, datasets):
    """posterior_mean$f / p0 / Delta and LPML (R/bayesynergy.R:244-283) are reduced on the device from the resident
    draws (bsyn_sample's `surfaces` / `lpml` sinks): compare with numpy mean / median on the returned draws, for the
    mean (non-robust) and the median (robust) response surface."""
    from bayesynergy_b200.api import bayesynergy
    y, x = datasets["mathews:ispinesib + ibrutinib"]
    for robust in (False, True):
        fit = bayesynergy(y, x, iter=600, seed=3, robust=robust)
        po, pm = fit["posterior"], fit["posterior_mean"]
        n_save = po["lp__"].shape[0]
        f = np.empty((n_save, 6, 6))
        f[:, 0, 0] = 1.0
        f[:, 1:, 0] = po["p02"]
        f[:, 0, 1:] = po["p01"]
        f[:, 1:, 1:] = po["p0"] + po["Delta"]
        p0 = f.copy()
        p0[:, 1:, 1:] = po["p0"]
        Delta = np.zeros_like(f)
        Delta[:, 1:, 1:] = po["Delta"]
        f_ref = np.median(f, axis=0) if robust else f.mean(axis=0)
        assert np.allclose(pm["f"], f_ref, rtol=1e-12, atol=1e-14), robust
        assert np.allclose(pm["p0"], np.median(p0, axis=0), rtol=1e-12, atol=1e-14)
        assert np.allclose(pm["Delta"], np.median(Delta, axis=0), rtol=1e-12, atol=1e-14)
        lpml = -np.sum(np.log(po["CPO"].sum(axis=0) / n_save))
        assert abs(fit["LPML"] - lpml) < 1e-9 * abs(lpml)


@pytest.mark.gpu
def test_screen_bayes_factor_and_samples(built):
    """cfg4: synergyscreen(bayes_factor = TRUE) -- one gp_grid launch, one nointeraction launch, one batched log_prob
    launch per model for all bridge points -- and return_samples = TRUE.  The Bayes factors are checked against the
    bridge estimator fed by the ORACLE's log_prob on the same draws of 8 combinations."""
    import warnings
    from bayesynergy_b200 import _lib
    from bayesynergy_b200.api import synergyscreen, SCREEN_COLUMNS
    from bayesynergy_b200.bridge import bridge_logml_many
    from bayesynergy_b200.dataprep import prepare, synthetic_experiment
    from oracle.c_oracle import Problem
    n = 8
    exps = []
    for k in range(n):
        yy, xx = synthetic_experiment(300 + k)
        exps.append(dict(y=yy, x=xx, drug_names=[f"A{k}", f"B{k}"], experiment_ID="syn"))
    launches0 = _lib.load().bsyn_launch_count()
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        out = synergyscreen(exps, return_samples=True, max_retries=0,
                            bayesynergy_params=dict(bayes_factor=True, iter=600, chains=4, seed=2))
    launches = _lib.load().bsyn_launch_count() - launches0
    rows = out["screenSummary"]
    assert len(rows) + len(out.get("failed", [])) == n and len(rows) >= n - 1
    assert list(rows[0].keys()) == SCREEN_COLUMNS + ["Bayes Factor"]
    bfs = np.array([r["Bayes Factor"] for r in rows])
    assert np.isfinite(bfs).all() and (bfs > 0).all()
    # batched: the launch count does not grow with the number of experiments (2 sampler launches + summaries / ESS +
    # 2 log_prob launches), far below one launch per experiment and model
    assert launches <= 32, launches      # first pass + one retry pass + two Bayes-factor groups (each: sampler, summaries, ESS, 3 bridge kernels), whatever n is
    samples = out["screenSamples"]
    assert len(samples) == len(rows)
    s0 = samples[0]
    assert s0["posterior"]["Delta"].shape == (1200, 10, 10) and s0["posterior_mean"]["f"].shape == (11, 11)
    assert abs(s0["posterior_mean"]["VUS_syn"] - rows[0]["Synergy (mean)"]) < 1e-9
    assert np.isfinite(s0["LPML"]) and s0["bayesfactor"] == rows[0]["Bayes Factor"]
    # the same estimator with the oracle's log_prob on freshly sampled draws of every experiment
    sds = [prepare(e["y"], e["x"]) for e in exps]

    class _OracleBatch:
        def __init__(self, model):
            self.P = [Problem(sd, model) for sd in sds]
            self.max_params = max(p.dim for p in self.P)

        def num_params(self, k=0):
            return self.P[k].dim

        def logp_grad(self, u, prob_idx=None, jacobian=True, want_grad=True):
            u = np.asarray(u)
            lp, st = np.empty(len(u)), np.empty(len(u), dtype=np.int32)
            for k in np.unique(prob_idx):
                m = prob_idx == k
                s_, l_, _ = self.P[k].logp_grad_batch(np.ascontiguousarray(u[m][:, :self.P[k].dim]), jacobian=jacobian)
                lp[m], st[m] = l_, s_
            return lp, None, st

    lml = {}
    for model in ("gp_grid", "nointeraction"):
        b = _lib.Batch(sds, model=model)
        res = b.sample(chains=4, iter=600, warmup=300, seed=2, want_udraws=True, want_qdraws=False, want_sampler_params=False)
        ud = [res["udraws"][k] for k in range(n)]
        dev, _ = bridge_logml_many(b, ud, seed=4)
        orc, _ = bridge_logml_many(_OracleBatch(model), ud, seed=4)
        b.close()
        assert np.allclose(dev, orc, rtol=1e-7, atol=1e-6), (model, dev, orc)
        lml[model] = dev
    # the screen's Bayes factors agree in log scale with this independent run (different draws: Monte Carlo error)
    ids = [r["listID"] - 1 for r in rows]
    ref = (lml["gp_grid"] - lml["nointeraction"])[ids]
    # (bridge sampling in 115 dimensions from 1200 draws is noisy: the strict check is the same-draws one above)
    d = np.abs(np.log(bfs) - ref)
    assert np.median(d) < 1.5 and d.max() < 8.0, (np.log(bfs), ref)


@pytest.mark.gpu
def test_device_bridge_matches_host_estimator(built, datasets):
    """The bridge-sampling estimator run entirely on the device (bsyn_sample's bridge_logml sink: proposal fit, batched
    log_prob, iterative scheme) against the host estimator of bridge.py on the SAME posterior draws.  The two differ only
    in their proposal draws (Philox on the device, numpy on the host), i.e. by Monte Carlo error of the estimator."""
    from bayesynergy_b200 import _lib
    from bayesynergy_b200.bridge import bridge_logml_many
    from bayesynergy_b200.dataprep import prepare, synthetic_experiment
    y, x = datasets["mathews:ispinesib + ibrutinib"]
    sds = [prepare(y, x), prepare(*datasets["mathews:canertinib + ibrutinib"])]
    for model, tol in (("gp_grid", 0.25), ("nointeraction", 0.05)):
        b = _lib.Batch(sds, model=model)
        res = b.sample(chains=4, iter=2000, warmup=1000, seed=5, want_udraws=True, want_bridge=True, want_qdraws=False)
        host, _ = bridge_logml_many(b, [res["udraws"][k] for k in range(2)], seed=3)
        host2, _ = bridge_logml_many(b, [res["udraws"][k] for k in range(2)], seed=4)      # the host's own proposal noise
        b.close()
        dev = res["bridge_logml"]
        assert np.isfinite(dev).all() and (res["bridge_niter"] > 0).all() and (res["bridge_niter"] < 1000).all()
        assert np.abs(dev - host).max() < max(tol, 4 * np.abs(host - host2).max()), (model, dev, host, host2)
    # 10 x 10 grid (D = 115): one batched launch over several experiments
    sds = [prepare(*synthetic_experiment(40 + k)) for k in range(4)]
    b = _lib.Batch(sds)
    res = b.sample(chains=4, iter=1000, warmup=500, seed=8, want_udraws=True, want_bridge=True, want_qdraws=False)
    host, _ = bridge_logml_many(b, [res["udraws"][k] for k in range(4)], seed=3)
    b.close()
    assert np.isfinite(res["bridge_logml"]).all()
    assert np.abs(res["bridge_logml"] - host).max() < 2.0, (res["bridge_logml"], host)     # 115 dimensions, 2000 draws: noisy
