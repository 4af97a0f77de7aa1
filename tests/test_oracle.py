"""CPU tests: the oracle (C restatement with hand adjoints) against the committed golden vectors
(torch float64 autograd), against SURVEY.md Appendix E, and its NUTS against the vignette's anchors."""
import numpy as np
import pytest

from bayesynergy_b200.dataprep import prepare, output_names, param_names
from oracle.c_oracle import Problem, nuts_screen

REL = 1e-10   # north_star: relative 1e-10 in FP64


def tol_for(cond):
    """1e-10 relative wherever the two kernel matrices have condition number <= 1e7; beyond that two correct
    FP64 implementations differ by ~eps*cond (jitter is only 1e-10, gp_grid.stan:113: SURVEY.md §7.4), so the
    bar scales with cond / 1e7.  Measured disagreement C-oracle vs autograd: <= 5e-13 below cond 1e6."""
    return max(REL, 1e-17 * cond)


def relerr(a, b):
    a, b = np.asarray(a, dtype=float), np.asarray(b, dtype=float)
    return np.abs(a - b).max() / max(1.0, np.abs(b).max())


def test_c_oracle_matches_golden(datasets, golden_cases):
    cache = {}
    worst = 0.0
    for c in golden_cases:
        key = (c["dataset"], tuple(sorted(c["variant"].items())), c["model"])
        if key not in cache:
            y, x = datasets[c["dataset"]]
            cache[key] = Problem(prepare(y, x, **c["variant"]), c["model"])
        P = cache[key]
        st, lp, g = P.logp_grad(np.array(c["u"]))
        assert st == 0
        e = max(abs(lp - c["lp"]) / max(1.0, abs(c["lp"])), relerr(g, c["grad"]))
        assert e <= tol_for(c["cond"]), (c["dataset"], c["variant"], c["model"], e, c["cond"])
        if c["cond"] <= 1e6:
            worst = max(worst, e)
    assert worst <= 1e-11


def test_appendix_e_probe(datasets):
    # SURVEY.md Appendix E (self-derived during the survey with an independent throwaway restatement + mpmath)
    y, x = datasets["mathews:ispinesib + ibrutinib"]
    u = 0.3 * np.sin(np.arange(40) + 1.0)
    st, lp, g = Problem(prepare(y, x)).logp_grad(u)
    assert st == 0 and abs(lp - (-61.13253597186737584778953)) < 1e-11
    assert abs(np.abs(g).max() - 32.6499056939939) < 1e-9
    st, lp, g = Problem(prepare(y, x, type=2)).logp_grad(u)
    assert abs(lp - (-61.13391566814387529627098)) < 1e-11
    st, lp, g = Problem(prepare(y, x, robust=True, pcprior=True)).logp_grad(u)
    assert abs(lp - (-62.2232201595374)) < 1e-10
    u2 = u.copy()
    u2[37] = -3.0
    st, lp, g = Problem(prepare(y, x, robust=True, pcprior=True)).logp_grad(u2)
    assert abs(lp - (-165.640708515838)) < 1e-9 and abs(g[37] - 61.97934236433) < 1e-8
    st, lp, g = Problem(prepare(y, x, robust=True, pcprior=True, heteroscedastic=False)).logp_grad(u2)
    assert abs(lp - (-143.007218995399)) < 1e-9 and abs(g[12] - (-11.13804035505)) < 1e-8


def test_mpmath_forward(datasets):
    # 50-digit forward evaluation of lp at one well- and one ill-conditioned point
    from oracle.mp_oracle import log_prob_mp
    y, x = datasets["mathews:ispinesib + ibrutinib"]
    sd = prepare(y, x)
    P = Problem(sd)
    u = 0.3 * np.sin(np.arange(40) + 1.0)
    st, lp, _ = P.logp_grad(u, want_grad=False)
    assert abs(lp - float(log_prob_mp(sd, u))) < 1e-11
    u2 = u.copy()
    u2[10] = np.log(4.0)    # ell = 4: kernel matrices with condition number ~1e6-1e8
    st, lp, _ = P.logp_grad(u2, want_grad=False)
    ref = float(log_prob_mp(sd, u2))
    assert abs(lp - ref) < 1e-6 * max(1.0, abs(ref))


def test_domain_error_is_reported(datasets):
    y, x = datasets["mathews:ispinesib + ibrutinib"]
    P = Problem(prepare(y, x, type=2))
    u = np.zeros(P.dim)
    u[11] = 800.0   # sigma_f = exp(800) = inf: cov1 is not finite -> cholesky_decompose throws in the reference
    st, lp, g = P.logp_grad(u)
    assert st == 1
    u[11] = 0.0
    u[10] = 6.0     # RBF with ell = e^6: near-singular, but the 1e-10 jitter keeps it positive definite
    st, lp, g = P.logp_grad(u)
    assert st == 0 and np.isfinite(lp)


def test_write_array_layout(datasets):
    y, x = datasets["synthetic:7"]
    sd = prepare(y, x)
    for model in ("gp_grid", "nointeraction"):
        P = Problem(sd, model)
        names = output_names(sd, model)
        assert len(names) == P.n_out and len(param_names(sd, model)) == P.dim
        u = np.random.default_rng(1).uniform(-1, 1, P.dim)
        st, out = P.write_array(u)
        assert st == 0
        v = dict(zip(names, out))
        assert abs(v["ec50_1"] - 10 ** v["log10_ec50_1"]) < 1e-12 * v["ec50_1"]
        assert 0 <= v["p0[3,4]"] <= 1 and abs(v["p0[3,4]"] - v["p01[4]"] * v["p02[3]"]) < 1e-15
        assert abs(v["s"] - np.exp(u[P.dim - 3])) < 1e-15
        if model == "gp_grid":
            assert abs(v["VUS_Delta"] - (v["VUS_syn"] + v["VUS_ant"])) < 1e-9
            assert abs(v["rVUS_f"] - (v["rVUS_p0"] - v["VUS_Delta"])) < 1e-9
        assert np.all(out[names.index("CPO[1]"):names.index("CPO[1]") + sd.N] > 0)


def test_oracle_nuts_reproduces_vignette_anchor(datasets):
    # vignettes/Example_single.Rmd:62-65: rVUS_f ~ 80, rVUS_p0 ~ 70, VUS_syn ~ 10 (in absolute value)
    y, x = datasets["mathews:ispinesib + ibrutinib"]
    sd = prepare(y, x)
    names = output_names(sd)
    ng, outs, sp = nuts_screen([Problem(sd)], n_chains=4, num_warmup=300, num_samples=300, seed=5)
    assert ng > 0
    f = outs[0, :, :, names.index("rVUS_f")].mean()
    p0 = outs[0, :, :, names.index("rVUS_p0")].mean()
    syn = outs[0, :, :, names.index("VUS_syn")].mean()
    assert 75 < f < 90 and 65 < p0 < 80 and -16 < syn < -5
    assert sp[..., 0].mean() > 0.7


def test_oracle_advi_reproduces_vignette_anchor(datasets):
    """Full-rank ADVI restatement (oracle/gp_oracle_advi.c): the vignette's example again (efficacy ~80, Bliss ~70,
    i.e. interaction ~ -10; vignettes/Example_single.Rmd:62-65), plus the bookkeeping of advi::run."""
    from oracle.c_oracle import advi_screen
    sd = prepare(*datasets["mathews:ispinesib + ibrutinib"])
    names = output_names(sd)
    mean, outs, info = advi_screen([Problem(sd)] * 6, seed=2, iter=2000)
    ok = info[:, 0] == 0
    assert ok.sum() >= 4 and set(info[~ok, 0]) <= {3.0, 4.0}
    assert set(info[ok, 1]) <= {100.0, 10.0, 1.0, 0.1, 0.01}
    assert (info[ok, 2] <= 2000).all() and (info[ok, 2] % 100 == 0).all()
    assert ((info[ok, 5] - info[ok, 2]) % 50 == 0).all()        # gradient evaluations: iterations + 50 per eta tried
    f = outs[ok][:, :, names.index("rVUS_f")].mean()
    p0 = outs[ok][:, :, names.index("rVUS_p0")].mean()
    assert 76 < f < 90 and 64 < p0 < 82 and -16 < (p0 - f) < -3
    assert np.isnan(outs[~ok]).all()
    # without adaptation the given eta is used and the run ends at an ELBO check or at the cap
    mean, outs, info = advi_screen([Problem(sd)], seed=4, iter=250, adapt_engaged=0, eta=0.1, output_samples=10)
    assert info[0, 0] in (0.0, 4.0) and (info[0, 0] != 0 or (info[0, 1] == 0.1 and info[0, 2] in (100.0, 200.0, 250.0)))


def test_oracle_dense_metric_agrees_with_diag(datasets):
    """hmc_nuts_dense_e_adapt restatement (what robust = TRUE selects, R/bayesynergy.R:112-115): same posterior as
    diag_e within 3 Monte Carlo standard errors; the adapted metric's diagonal is a positive variance estimate."""
    from bayesynergy_b200.diagnostics import mcse_mean
    sd = prepare(*datasets["mathews:ispinesib + ibrutinib"])
    names = output_names(sd)
    ng0, o0, sp0 = nuts_screen([Problem(sd)], 4, 600, 600, seed=3)
    ng1, o1, sp1 = nuts_screen([Problem(sd)], 4, 600, 600, seed=4, metric=1)
    for qn in ("rVUS_f", "rVUS_p0", "VUS_syn", "VUS_ant"):
        a, c = o0[0, :, :, names.index(qn)], o1[0, :, :, names.index(qn)]
        assert abs(a.mean() - c.mean()) <= 3.5 * np.hypot(mcse_mean(a), mcse_mean(c)), qn
    assert sp1[0, :, :, 4].mean() < 0.02 and 1 <= sp1[0, :, :, 2].min() and sp1[0, :, :, 2].max() <= 10
    assert 0.75 < sp1[0, :, :, 0].mean() < 0.99
