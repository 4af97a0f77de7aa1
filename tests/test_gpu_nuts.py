"""GPU tests of the device-resident NUTS sampler against the oracle's NUTS (statistical agreement:
posterior means within 3 Monte Carlo standard errors, the north_star's second correctness criterion),
plus the on-device summaries / ESS against numpy."""
import numpy as np
import pytest

from bayesynergy_b200 import _lib
from bayesynergy_b200.dataprep import prepare, output_names
from bayesynergy_b200.diagnostics import split_ess_rhat, mcse_mean, bulk_ess_rhat, tail_ess
from oracle.c_oracle import Problem, nuts_screen

pytestmark = pytest.mark.gpu
Q = _lib.Q_NAMES


def _compare(res_q, outs, names, tag):
    for qn in ("rVUS_f", "rVUS_p0", "VUS_Delta", "VUS_syn", "VUS_ant", "s", "dss_1", "dss_2"):
        if qn not in names:
            continue
        a = res_q[:, :, Q.index(qn)]
        c = outs[:, :, names.index(qn)]
        se = np.hypot(mcse_mean(a), mcse_mean(c))
        # north_star: synergy / antagonism (VUS) scores within 3 MCSE; the nuisance scalars get 4 (16 comparisons
        # per test at 3 sigma would fail spuriously every ~20 runs)
        nsig = 3.0 if "VUS" in qn else 4.0
        assert abs(a.mean() - c.mean()) <= nsig * se + 1e-6, (tag, qn, a.mean(), c.mean(), se)


def test_nuts_matches_oracle_mathews(built, datasets):
    for dname in ("mathews:ispinesib + ibrutinib", "mathews:canertinib + ibrutinib"):
        y, x = datasets[dname]
        sd = prepare(y, x)
        b = _lib.Batch([sd])
        res = b.sample(chains=4, iter=2000, warmup=1000, seed=11)
        cs = res["chain_stats"][0]
        assert (cs[:, 6] == 0).all() and (cs[:, 0] > 1e-3).all()            # initialised, sane step sizes
        assert 0.75 < cs[:, 5].mean() < 0.99                                # adapt_delta = 0.9
        sp = res["sampler_params"][0]
        assert sp[:, :, 3].min() >= 1 and sp[:, :, 2].max() <= 10
        ng, outs, spo = nuts_screen([Problem(sd)], n_chains=4, num_warmup=1000, num_samples=1000, seed=21)
        _compare(res["qdraws"][0], outs[0], output_names(sd), dname)
        # device summaries / ESS vs numpy on the returned draws
        q = res["qdraws"][0]
        for qi in (Q.index("VUS_syn"), Q.index("rVUS_f"), Q.index("s")):
            assert abs(res["summary"][0, qi, 0] - q[:, :, qi].mean()) < 1e-9 * max(1, abs(q[:, :, qi].mean()))
            assert abs(res["summary"][0, qi, 1] - q[:, :, qi].std(ddof=1)) < 1e-9 * max(1, q[:, :, qi].std())
            ess_d, rhat_d = b.ess(qi)
            ess_h, rhat_h = split_ess_rhat(q[:, :, qi])
            assert abs(ess_d[0] - ess_h) < 1e-6 * ess_h and abs(rhat_d[0] - rhat_h) < 1e-9
            assert rhat_h < 1.05
            # rank-normalised bulk ESS / R-hat and tail ESS (the estimator of SURVEY 8(d)): device sort vs scipy ranks
            eb_d, rb_d, et_d = b.bulk_ess(qi)
            eb_h, rb_h = bulk_ess_rhat(q[:, :, qi])
            assert abs(eb_d[0] - eb_h) < 1e-6 * eb_h and abs(rb_d[0] - rb_h) < 1e-9, (eb_d[0], eb_h)
            et_h = tail_ess(q[:, :, qi])
            assert abs(et_d[0] - et_h) < 1e-6 * et_h, (et_d[0], et_h)
        b.close()


def test_nuts_draw_sinks_are_consistent(built, datasets):
    y, x = datasets["mathews:ispinesib + ibrutinib"]
    sd = prepare(y, x)
    b = _lib.Batch([sd, sd])
    names = output_names(sd)
    res = b.sample(chains=2, iter=120, warmup=60, seed=4, want_draws=True, want_udraws=True, want_cpo=True,
                   want_inv_metric=True)
    assert res["draws"].shape == (2, 2, 60, b.max_outputs)
    d, ud, q = res["draws"], res["udraws"], res["qdraws"]
    for qn in ("ec50_1", "dss_2", "s", "rVUS_f", "VUS_syn", "VUS_ant"):
        assert np.array_equal(d[..., names.index(qn)], q[..., Q.index(qn)])
    # constrained draws are write_array of the unconstrained ones
    out, st = b.write_array(ud[0, 1, :5, :b.num_params(0)], seed=1)
    cols = [i for i, nm in enumerate(names) if nm not in ("VUS_syn", "VUS_ant", "rVUS_f", "rVUS_p0")]
    assert np.allclose(out[:, cols], d[0, 1, :5][:, cols], rtol=1e-12, atol=1e-14)
    # lp__ equals log_prob at the draw
    lp, _, st = b.logp_grad(ud[1, 0, :5, :b.num_params(0)], prob_idx=np.ones(5, dtype=np.int32), want_grad=False)
    assert np.allclose(lp, res["sampler_params"][1, 0, :5, 6], rtol=1e-12)
    # LPML ingredients: mean CPO over draws
    k0 = names.index("CPO[1]")
    cpo = d[0][:, :, k0:k0 + sd.N].reshape(-1, sd.N).mean(axis=0)
    assert np.allclose(res["cpo_mean"][0, :sd.N], cpo, rtol=1e-10)
    assert (res["inv_metric"][0, :, :b.num_params(0)] > 0).all()
    # the two identical problems use different RNG streams
    assert not np.array_equal(q[0], q[1])
    b.close()


def test_nuts_options_run_and_agree(built, datasets):
    """robust / RBF / nointeraction variants and a larger grid: sampler runs, no init failures, agrees with oracle."""
    y, x = datasets["synthetic:7"]
    for v, model in ((dict(), "gp_grid"), (dict(), "nointeraction"), (dict(type=2, robust=True), "gp_grid")):
        sd = prepare(y, x, **v)
        b = _lib.Batch([sd], model=model)
        res = b.sample(chains=4, iter=600, warmup=300, seed=2)
        assert (res["chain_stats"][0, :, 6] == 0).all()
        ng, outs, spo = nuts_screen([Problem(sd, model)], n_chains=4, num_warmup=300, num_samples=300, seed=9)
        _compare(res["qdraws"][0], outs[0], output_names(sd, model), (v, model))
        b.close()


def test_tree_statistics_match_oracle(built, datasets):
    """The iterative tree builder against the oracle's recursive base_nuts over many chains: adapted step size,
    tree depth and leapfrogs per iteration must agree in distribution (chain-to-chain spread averaged out)."""
    y, x = datasets["mathews:ispinesib + ibrutinib"]
    sd = prepare(y, x)
    nrep = 6                                  # 6 copies x 4 chains = 24 chains each side
    b = _lib.Batch([sd] * nrep)
    res = b.sample(chains=4, iter=800, warmup=400, seed=17)
    sp = res["sampler_params"]
    ng, outs, spo = nuts_screen([Problem(sd)] * nrep, n_chains=4, num_warmup=400, num_samples=400, seed=29)
    eps_g, eps_o = np.log(sp[..., 1].mean(axis=2)).ravel(), np.log(spo[..., 1].mean(axis=2)).ravel()
    dep_g, dep_o = sp[..., 2].mean(axis=2).ravel(), spo[..., 2].mean(axis=2).ravel()
    se_eps = np.hypot(eps_g.std(ddof=1), eps_o.std(ddof=1)) / np.sqrt(len(eps_g))
    se_dep = np.hypot(dep_g.std(ddof=1), dep_o.std(ddof=1)) / np.sqrt(len(dep_g))
    assert abs(eps_g.mean() - eps_o.mean()) < 4 * se_eps + 0.05, (eps_g.mean(), eps_o.mean(), se_eps)
    assert abs(dep_g.mean() - dep_o.mean()) < 4 * se_dep + 0.1, (dep_g.mean(), dep_o.mean(), se_dep)
    assert abs(sp[..., 0].mean() - spo[..., 0].mean()) < 0.03          # accept_stat ~ adapt_delta on both sides
    b.close()


def test_dense_metric_agrees_with_diag(built, datasets):
    """dense_e (what bayesynergy() selects for robust = TRUE, R/bayesynergy.R:112-115, with stepsize_jitter 0.5):
    same posterior as diag_e within Monte Carlo error, sane adaptation, and a non-trivial adapted metric."""
    y, x = datasets["mathews:ispinesib + ibrutinib"]
    for v in (dict(), dict(robust=True)):
        sd = prepare(y, x, **v)
        b = _lib.Batch([sd, sd])
        jit = 0.5 if v else 0.0
        rd = b.sample(chains=4, iter=1200, warmup=600, seed=31, metric=1, stepsize_jitter=jit)
        r0 = b.sample(chains=4, iter=1200, warmup=600, seed=32, metric=0, stepsize_jitter=jit)
        assert (rd["chain_stats"][:, :, 6] == 0).all()
        assert 0.7 < rd["chain_stats"][:, :, 5].mean() < 0.99
        for qn in ("rVUS_f", "rVUS_p0", "VUS_syn", "VUS_ant", "s", "dss_1"):
            a = rd["qdraws"][:, :, :, Q.index(qn)].reshape(8, -1)
            c = r0["qdraws"][:, :, :, Q.index(qn)].reshape(8, -1)
            se = np.hypot(mcse_mean(a), mcse_mean(c))
            assert abs(a.mean() - c.mean()) <= 4.0 * se + 1e-6, (v, qn, a.mean(), c.mean(), se)
        # the dense sampler should not need more leapfrogs per iteration than the diagonal one
        assert rd["sampler_params"][..., 3].mean() < 1.5 * r0["sampler_params"][..., 3].mean()
        b.close()


def test_dense_tree_statistics_match_oracle(built, datasets):
    """dense_e against the oracle's hmc_nuts_dense_e_adapt restatement (Welford covariance windows, regularisation,
    p = L^-T z): adapted step size, tree depth and posterior means agree over 24 chains each."""
    y, x = datasets["mathews:ispinesib + ibrutinib"]
    sd = prepare(y, x)
    nrep = 6
    b = _lib.Batch([sd] * nrep)
    res = b.sample(chains=4, iter=1000, warmup=500, seed=41, metric=1)
    b.close()
    sp = res["sampler_params"]
    ng, outs, spo = nuts_screen([Problem(sd)] * nrep, n_chains=4, num_warmup=500, num_samples=500, seed=43, metric=1)
    eps_g, eps_o = np.log(sp[..., 1].mean(axis=2)).ravel(), np.log(spo[..., 1].mean(axis=2)).ravel()
    dep_g, dep_o = sp[..., 2].mean(axis=2).ravel(), spo[..., 2].mean(axis=2).ravel()
    se_eps = np.hypot(eps_g.std(ddof=1), eps_o.std(ddof=1)) / np.sqrt(len(eps_g))
    se_dep = np.hypot(dep_g.std(ddof=1), dep_o.std(ddof=1)) / np.sqrt(len(dep_g))
    assert abs(eps_g.mean() - eps_o.mean()) < 4 * se_eps + 0.05, (eps_g.mean(), eps_o.mean(), se_eps)
    assert abs(dep_g.mean() - dep_o.mean()) < 4 * se_dep + 0.1, (dep_g.mean(), dep_o.mean(), se_dep)
    names = output_names(sd)
    for qn in ("rVUS_f", "rVUS_p0", "VUS_syn", "VUS_ant"):
        a = res["qdraws"][..., Q.index(qn)].reshape(4 * nrep, -1)
        c = outs[..., names.index(qn)].reshape(4 * nrep, -1)
        se = np.hypot(mcse_mean(a), mcse_mean(c))
        assert abs(a.mean() - c.mean()) <= 3.0 * se + 1e-6, (qn, a.mean(), c.mean(), se)


def test_oneil_shaped_screen(built, datasets):
    """cfg2-shaped problems (11x11 grid, 32 of 144 cells observed, 9-12 replicates; the two raw experiments shipped in
    ONeil_A375$failed) through synergyscreen(): two-pass strip products, unobserved cells, > 4 replicates per cell."""
    from bayesynergy_b200.api import synergyscreen, SCREEN_COLUMNS
    exps = []
    for k in (0, 1):
        y, x = datasets[f"oneil_failed:{k}"]
        exps.append(dict(y=y, x=x, drug_names=[f"A{k}", f"B{k}"], experiment_ID="A375"))
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        out = synergyscreen(exps, bayesynergy_params=dict(iter=600, chains=4, seed=3), max_retries=1)
    rows = out["screenSummary"] + [None] * len(out.get("failed", []))
    assert len(rows) == 2
    for r in out["screenSummary"]:
        assert list(r.keys()) == SCREEN_COLUMNS
        assert 0 < r["Efficacy (mean)"] < 100 and r["Synergy (mean)"] <= 1e-3 and r["Antagonism (mean)"] >= 0
        assert np.isfinite(r["Synergy Score"]) and r["s"] > 0
    # statistical agreement with the oracle on the first experiment
    sd = prepare(*datasets["oneil_failed:0"])
    b = _lib.Batch([sd])
    res = b.sample(chains=4, iter=1000, warmup=500, seed=8)
    ng, outs, spo = nuts_screen([Problem(sd)], n_chains=4, num_warmup=500, num_samples=500, seed=12)
    _compare(res["qdraws"][0], outs[0], output_names(sd), "oneil_failed:0")
    b.close()


def test_cfg3_shape_dense_robust_pcprior_matches_oracle(built):
    """cfg3 of BASELINE.json on its real shape: O'Neil-shaped 11 x 11 grids (D = 136, 32 observed cells), robust
    likelihood + PC prior, which makes bayesynergy() select dense_e + stepsize_jitter 0.5 (R/bayesynergy.R:112-115).
    Device sampler vs the oracle's hmc_nuts_dense_e_adapt restatement on 6 combinations x 4 chains: VUS means within 3
    MCSE, adapted step size / tree depth distributions agree."""
    from bayesynergy_b200.dataprep import synthetic_oneil_experiment
    ncomb, it, wu = 6, 1000, 500
    sds = [prepare(*synthetic_oneil_experiment(k), robust=True, pcprior=True) for k in range(ncomb)]
    b = _lib.Batch(sds)
    res = b.sample(chains=4, iter=it, warmup=wu, seed=51, metric=1, stepsize_jitter=0.5)
    b.close()
    assert (res["chain_stats"][:, :, 6] == 0).all()
    ng, outs, spo = nuts_screen([Problem(sd) for sd in sds], n_chains=4, num_warmup=wu, num_samples=it - wu, seed=53,
                                metric=1, stepsize_jitter=0.5)
    names = output_names(sds[0])
    zs = []
    for k in range(ncomb):
        for qn in ("VUS_syn", "VUS_ant", "rVUS_f", "rVUS_p0"):
            a = res["qdraws"][k][:, :, Q.index(qn)]
            c = outs[k][:, :, names.index(qn)]
            se = np.hypot(mcse_mean(a), mcse_mean(c))
            zs.append((abs(a.mean() - c.mean()) - 1e-6) / se)
    zs = np.array(zs)
    # 24 comparisons: every one within 4 MCSE, at most one beyond 3 (P(|z| > 3) = 0.27 % each under agreement)
    assert zs.max() <= 4.0 and (zs > 3.0).sum() <= 1, zs
    sp = res["sampler_params"]
    eps_g, eps_o = np.log(sp[..., 1].mean(axis=2)).ravel(), np.log(spo[..., 1].mean(axis=2)).ravel()
    dep_g, dep_o = sp[..., 2].mean(axis=2).ravel(), spo[..., 2].mean(axis=2).ravel()
    se_eps = np.hypot(eps_g.std(ddof=1), eps_o.std(ddof=1)) / np.sqrt(len(eps_g))
    se_dep = np.hypot(dep_g.std(ddof=1), dep_o.std(ddof=1)) / np.sqrt(len(dep_g))
    assert abs(eps_g.mean() - eps_o.mean()) < 4 * se_eps + 0.05, (eps_g.mean(), eps_o.mean(), se_eps)
    assert abs(dep_g.mean() - dep_o.mean()) < 4 * se_dep + 0.15, (dep_g.mean(), dep_o.mean(), se_dep)
    assert abs(sp[..., 0].mean() - spo[..., 0].mean()) < 0.04


def test_sample_to_ess_extends_only_what_is_below_target(built):
    """ESS-target control (north_star: every combination sampled to an ESS of 400): warm-up + 150 draws, then extension
    launches over the combinations still below the target, chains resuming from their adapted state."""
    from bayesynergy_b200.dataprep import synthetic_experiment
    from bayesynergy_b200.diagnostics import bulk_ess
    n = 48
    sds = [prepare(*synthetic_experiment(1000 + k)) for k in range(n)]
    b = _lib.Batch(sds)
    block, rounds_max = 150, 6
    r = b.sample_to_ess(target=400.0, block=block, max_rounds=rounds_max, retry_rounds=0, chains=4, warmup=300, seed=5,
                        want_summary=True, want_qdraws=True)
    assert r["n_retried"] == 0 and abs(r["phase1_ms"] - r["kernel_ms"]) < 1e-6 * r["kernel_ms"]
    nd, ess = r["draws_per_chain"], r["ess"]
    assert r["rounds"] >= 1 and (nd % block == 0).all() and nd.min() >= block and nd.max() <= block * rounds_max
    assert nd.max() == block * r["rounds"]
    assert ((ess >= 400.0) | (nd == block * rounds_max)).all()
    assert (nd > block).any() and (nd == block).any()          # some combinations were extended, some were not
    q = r["qdraws"]
    for k in range(n):
        # the device ESS that stopped the combination = the host estimator on exactly its draws
        e = min(bulk_ess(q[k, :, :nd[k], Q.index("VUS_syn")]), bulk_ess(q[k, :, :nd[k], Q.index("VUS_ant")]))
        assert abs(e - ess[k]) < 1e-6 * max(1.0, e), (k, e, ess[k])
        assert (q[k, :, nd[k]:, :] == 0).all()                  # nothing written behind the last block
        m = q[k, :, :nd[k], Q.index("rVUS_f")].mean()
        assert abs(r["summary"][k, Q.index("rVUS_f"), 0] - m) < 1e-9 * max(1.0, abs(m))
    # rank-normalised ESS after the run uses each combination's own draw count
    eb, _, _ = b.bulk_ess("VUS_syn")
    k = int(np.argmax(nd))
    assert abs(eb[k] - bulk_ess(q[k, :, :nd[k], Q.index("VUS_syn")])) < 1e-6 * eb[k]
    # an extended run samples the same posterior as one long run: VUS means within 4 MCSE on the extended combinations
    long = b.sample(chains=4, iter=300 + block * rounds_max, warmup=300, seed=77)
    b.close()
    zs = []
    for k in np.where(nd > block)[0][:12]:
        for qn in ("VUS_syn", "VUS_ant", "rVUS_f"):
            a = q[k, :, :nd[k], Q.index(qn)]
            c = long["qdraws"][k][:, :, Q.index(qn)]
            zs.append(abs(a.mean() - c.mean()) / (np.hypot(mcse_mean(a), mcse_mean(c)) + 1e-9))
    zs = np.array(zs)
    assert zs.max() < 4.5 and (zs > 3.0).sum() <= 2, zs


def test_screen_to_ess_two_phases(built):
    """api.screen_to_ess: phase 1 (extension launches), then the reference's remedy for what is left: the same chains
    resume, re-adapt their step size to adapt_delta 0.99 and go on.  A deliberately small phase-1 budget forces phase 2."""
    from bayesynergy_b200.api import screen_to_ess
    from bayesynergy_b200.dataprep import synthetic_experiment
    from bayesynergy_b200.diagnostics import bulk_ess
    n, block = 32, 100
    sds = [prepare(*synthetic_experiment(2000 + k)) for k in range(n)]
    b = _lib.Batch(sds)
    r = screen_to_ess(sds, target=900.0, batch=b, chains=4, warmup=300, block=block, max_rounds=2, seed=9, retry="resume",
                      retry_warmup=100, retry_rounds=8)
    nd, ess = r["draws_per_chain"], r["ess"]
    assert ess.shape == (n,) and r["summary"].shape == (n, 12, 2) and r["chain_stats"].shape == (n, 4, 8)
    assert 0 < r["n_retry"] <= n and r["n_retry"] == int((nd > 2 * block).sum())      # phase 2 ran on exactly what was left
    assert r["rounds"] > 2 and r["reached"] >= n - 4
    assert r["phase2_ms"] > 0 and abs(r["kernel_ms"] - r["phase1_ms"] - r["phase2_ms"]) < 1e-6 * r["kernel_ms"]
    assert (nd >= block).all() and (nd % block == 0).all() and nd.max() == block * r["rounds"]
    f = r["summary"][:, Q.index("rVUS_f"), 0]
    assert np.isfinite(f).all() and ((f > 0) & (f < 100)).all()
    # the chains of a retried experiment ended with a smaller step size than those that stopped in phase 1 (delta 0.99 vs 0.9)
    eps = r["chain_stats"][:, :, 0]
    retried = nd > 2 * block
    if retried.any() and (~retried).any():
        assert np.median(eps[retried]) < np.median(eps[~retried])
    # the pooled draws (phase 1 + phase 2) are what the ESS on the device was computed from
    e_syn, _, _ = b.bulk_ess("VUS_syn")
    e_ant, _, _ = b.bulk_ess("VUS_ant")
    assert np.allclose(np.fmin(e_syn, e_ant), ess, rtol=1e-6)
    b.close()
    # retry="refit": what phase 1 leaves is fitted again from scratch in a small batch; the better of the two fits is kept
    r1 = screen_to_ess(sds, target=900.0, chains=4, warmup=300, block=block, max_rounds=2, seed=9, retry=None)
    r2 = screen_to_ess(sds, target=900.0, chains=4, warmup=300, block=block, max_rounds=2, seed=9, retry="refit",
                       retry_block=250, retry_rounds=6)
    assert r1["n_retry"] == 0 and r1["phase2_ms"] == 0 and r1["reached"] < n
    assert r2["n_retry"] == n - r1["reached"] and r2["phase2_ms"] > 0 and abs(r2["phase1_ms"] - r1["kernel_ms"]) < 0.5 * r1["kernel_ms"]
    assert (r2["ess"] >= r1["ess"] - 1e-9).all() and r2["reached"] >= n - 4
    with pytest.raises(ValueError):
        screen_to_ess(sds, retry="again")
