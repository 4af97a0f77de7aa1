"""CPU tests of the host-side views that sit behind the C ABI: the stanfit-shaped object (stanfit.py, the Python counterpart
of include/bsyn_stanfit.hpp) and rstan::extract-style assembly of a posterior from write_array rows (api.py)."""
import json
import os

import numpy as np
import pytest

from bayesynergy_b200.stanfit import StanFit, SAMPLER_PARAMS
from bayesynergy_b200.dataprep import prepare, output_names, param_names
from bayesynergy_b200 import api

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _fit(chains=4, n=500, seed=0):
    rng = np.random.default_rng(seed)
    flat = ["a", "z[1,1]", "z[2,1]", "z[1,2]", "z[2,2]", "z[1,3]", "z[2,3]", "v[1]", "v[2]", "VUS_syn"]
    blocks = [("a", [], 0), ("z", [2, 3], 1), ("v", [2], 7), ("empty", [0], 9), ("VUS_syn", [], 9), ("lp__", [], 10)]
    draws = rng.normal(size=(chains, n, len(flat) + 3))          # wider than the flat names, as the device rows are
    draws[:, :, 1:7] = np.arange(1, 7)[None, None, :] + 0.01 * draws[:, :, 1:7]
    sp = np.zeros((chains, n, 7))
    sp[:, :, 0] = 0.9
    sp[:, :, 1] = 0.05
    sp[:, :, 2] = rng.integers(3, 11, size=(chains, n))
    sp[:, :, 3] = 2 ** sp[:, :, 2] - 1
    sp[:, :, 4] = rng.random((chains, n)) < 0.02
    sp[:, :, 5] = 10.0
    sp[:, :, 6] = -50.0 + rng.normal(size=(chains, n))
    cs = np.zeros((chains, 8))
    cs[:, 0] = 0.05
    cs[:, 3] = 3000.0      # leapfrogs in total
    cs[:, 4] = 1000.0      # ... of which sampling
    return StanFit(flat, blocks, draws, sp, cs, inv_metric=np.ones((chains, 12)), inits=np.zeros((chains, 12)),
                   kernel_ms=2000.0, warmup=500), draws, sp


def test_stanfit_extract_shapes_and_column_major_order():
    fit, draws, sp = _fit()
    assert fit.flat_names[-1] == "lp__" and len(fit.flat_names) == 11
    e = fit.extract()
    assert e["a"].shape == (2000,) and e["z"].shape == (2000, 2, 3) and e["v"].shape == (2000, 2)
    assert e["empty"].shape == (2000, 0) and e["lp__"].shape == (2000,)
    # z[i,j] sits at 1 + (i-1) + 2 (j-1) in the flat vector: z[2,3] is the sixth entry (~ 6)
    assert abs(e["z"][:, 1, 2].mean() - 6.0) < 0.01 and abs(e["z"][:, 0, 1].mean() - 3.0) < 0.01
    assert np.array_equal(e["lp__"], sp[:, :, 6].reshape(-1))
    assert set(fit.extract(pars=["a", "z"])) == {"a", "z"}
    arr, names = fit.extract(permuted=False)
    assert arr.shape == (500, 4, 11) and names == fit.flat_names
    assert np.array_equal(arr[:, 2, 0], draws[2, :, 0])


def test_stanfit_summary_and_diagnostics():
    fit, draws, sp = _fit()
    s = fit.summary()
    assert s["rownames"] == fit.flat_names and s["colnames"][:3] == ["mean", "se_mean", "sd"]
    assert s["colnames"][-2:] == ["n_eff", "Rhat"] and s["summary"].shape == (11, 10)
    row = s["summary"][0]                                   # iid N(0,1) draws: n_eff ~ 2000, Rhat ~ 1
    assert abs(row[0]) < 0.1 and abs(row[2] - 1.0) < 0.1 and 1200 < row[-2] < 2800 and abs(row[-1] - 1.0) < 0.01
    assert abs(row[1] - row[2] / np.sqrt(row[-2])) < 1e-12
    one = fit.summary(pars=["z"])
    assert one["rownames"] == [f"z[{i},{j}]" for j in (1, 2, 3) for i in (1, 2)]
    gp = fit.get_sampler_params()
    assert len(gp) == 4 and list(gp[0]) == SAMPLER_PARAMS and np.array_equal(gp[1]["treedepth__"], sp[1, :, 2])
    assert fit.get_num_divergent() == int(sp[:, :, 4].sum()) == int(fit.get_divergent_iterations().sum())
    assert fit.get_max_treedepth_iterations().sum() == int((sp[:, :, 2] >= 10).sum())
    t = fit.get_elapsed_time()
    assert t.shape == (4, 2) and np.allclose(t.sum(axis=1), 2.0) and np.allclose(t[:, 0], 2.0 * 2000 / 3000)
    info = fit.get_adaptation_info()
    assert len(info) == 4 and "Step size = 0.05" in info[0] and "Diagonal elements of inverse mass matrix" in info[0]
    assert fit.get_inits().shape == (4, 12) and "4 chains x 500 draws" in repr(fit)


def test_posterior_from_draws_matches_the_flat_layout():
    """api._posterior_from_draws = rstan::extract on write_array rows: grids come back as [draw, n2, n1] (column-major in the
    flat vector, src/stanExports_gp_grid.h:2256-2402), scalars by name, chains pooled."""
    m = json.load(open(os.path.join(ROOT, "tests", "golden", "mathews_DLBCL.json")))
    e = m["ispinesib + ibrutinib"]
    sd = prepare(np.array(e["y"]), np.array(e["x"]).reshape(2, -1).T)
    names = output_names(sd, "gp_grid")
    chains, n = 2, 5
    draws = np.tile(np.arange(len(names), dtype=float), (chains, n, 1)) + 1000.0 * np.arange(chains)[:, None, None]
    lp = np.arange(chains * n, dtype=float)
    post = api._posterior_from_draws(sd, draws, lp)
    col = {nm: i for i, nm in enumerate(names)}
    assert post["p0"].shape == (chains * n, sd.n2, sd.n1) and post["Delta"].shape == post["p0"].shape
    assert post["p0"][0, 1, 2] == col["p0[2,3]"] and post["GP"][0, 0, 1] == col["GP[1,2]"] and post["z"][0, 2, 0] == col["z[3,1]"]
    assert post["CPO"].shape == (chains * n, sd.N) and post["CPO"][0, 0] == col["CPO[1]"]
    assert post["VUS_syn"][0] == col["VUS_syn"] and post["ell"][n] == 1000.0 + col["ell"] and np.array_equal(post["lp__"], lp)
    for nm in param_names(sd, "gp_grid"):
        base = nm.split("[")[0]
        assert base in post
    post0 = api._posterior_from_draws(sd, draws[:, :, :len(output_names(sd, "nointeraction"))], lp, model="nointeraction")
    assert "z" not in post0 and "Delta" not in post0 and post0["p0"].shape == (chains * n, sd.n2, sd.n1)


def test_surface_grids_are_column_major():
    class SD:
        n1, n2 = 2, 3
    ncf = 3 * 4
    s = np.stack([np.arange(ncf + 4, dtype=float) + 100 * k for k in range(3)])
    f, p0, d = api._surface_grids(SD, s)
    assert f.shape == (4, 3) and f[1, 0] == 1 and f[0, 1] == 4 and p0[3, 2] == 100 + 11 and d[0, 0] == 200


def _mathews_sd():
    m = json.load(open(os.path.join(ROOT, "tests", "golden", "mathews_DLBCL.json")))
    e = m["ispinesib + ibrutinib"]
    return prepare(np.array(e["y"]), np.array(e["x"]).reshape(2, -1).T), e


def test_diagnose_maps_sampler_warnings_to_messages():
    """R/bayesynergy.R:210-214: any rstan warning makes returnCode 1; api._diagnose restates the warnings rstan raises."""
    from bayesynergy_b200._lib import Q_NAMES, BSYN_NQ
    rng = np.random.default_rng(1)
    chains, n = 4, 400
    q = rng.normal(size=(chains, n, BSYN_NQ))
    sp = np.zeros((chains, n, 7))
    sp[:, :, 2] = 5
    cs = np.zeros((chains, 8))
    msgs, ndiv = api._diagnose(sp, cs, q, 10)
    assert msgs == [] and ndiv == 0
    sp[0, :3, 4] = 1
    sp[1, :7, 2] = 10
    cs[2, 6] = 1
    msgs, ndiv = api._diagnose(sp, cs, q, 10)
    assert ndiv == 3 and msgs[0] == "There were 3 divergent transitions after warmup."
    assert msgs[1].startswith("There were 7 transitions after warmup that exceeded the maximum treedepth")
    assert msgs[2].startswith("Initialization failed") and len(msgs) == 3
    q2 = q.copy()
    q2[0, :, Q_NAMES.index("VUS_ant")] += 5.0                       # one chain elsewhere: R-hat
    msgs, _ = api._diagnose(np.zeros((chains, n, 7)), np.zeros((chains, 8)), q2, 10)
    assert len(msgs) == 1 and "R-hat" in msgs[0]
    q3 = q.copy()
    q3[:, :, Q_NAMES.index("s")] = np.cumsum(q3[:, :, Q_NAMES.index("s")], axis=1) * 0.01 + q3[:, :1, Q_NAMES.index("s")] * 0  # random walk
    msgs, _ = api._diagnose(np.zeros((chains, n, 7)), np.zeros((chains, 8)), q3, 10)
    assert len(msgs) == 1 and ("ESS" in msgs[0] or "R-hat" in msgs[0])


def test_outlier_rule_and_object_fields():
    """R/bayesynergy.R:305-314 (residual > 3 x observation noise => warning, returnCode 1) and the fields of the object."""
    import warnings
    from bayesynergy_b200._lib import BSYN_NQ
    sd, e = _mathews_sd()
    ncf = (sd.n1 + 1) * (sd.n2 + 1)
    f_cells = np.zeros(ncf)
    f_cells[sd.ii_obs - 1] = sd.y                    # a surface through the data: no outlier
    assert api._outlier_message(sd, f_cells, 0.1, 0.005) is None
    f_bad = f_cells.copy()
    f_bad[sd.ii_obs[3] - 1] += 0.9
    msg = api._outlier_message(sd, f_bad, 0.1, 0.005)
    assert msg and "more than three times the observation noise" in msg
    rng = np.random.default_rng(2)
    chains, n = 4, 200
    names = output_names(sd, "gp_grid")
    draws = rng.random((chains, n, len(names)))
    post = api._posterior_from_draws(sd, draws, rng.normal(size=chains * n))
    post["s"] = 0.1 * post["s"]                      # observation noise scale ~ 0.05
    surf = np.zeros((3, ncf + 3))
    surf[0, :ncf] = f_cells
    options = dict(type=3, lower_asymptotes=True, method="sampling", bayes_factor=False, heteroscedastic=True, robust=False,
                   pcprior=False, lambda_=0.005, nu=1.5, pcprior_hypers=(1, 0.1, 1, 0.2))
    meta = dict(y=sd.y, x=np.zeros((sd.N, 2)), drug_names=["A", "B"], experiment_ID="e", units=["uM", "uM"])
    q = rng.normal(size=(chains, n, BSYN_NQ))
    with warnings.catch_warnings():
        warnings.simplefilter("error")               # a clean fit raises no warning
        obj = api._assemble_object(sd, meta, post, surf, -12.5, np.zeros((chains, n, 7)), np.zeros((chains, 8)), q, options, 10)
    assert obj["returnCode"] == 0 and "messages" not in obj and obj["LPML"] == -12.5 and np.isnan(obj["divergent"])
    assert set(obj) >= {"posterior", "posterior_mean", "data", "model", "returnCode", "LPML", "divergent"}
    assert obj["posterior_mean"]["f"].shape == (sd.n2 + 1, sd.n1 + 1) and obj["model"]["nu"] == 1.5
    assert "z" not in obj["posterior_mean"] and isinstance(obj["posterior_mean"]["ell"], float)
    assert np.array_equal(obj["data"]["indices"], sd.ii_obs)
    surf[0, :ncf] = f_bad
    sp = np.zeros((chains, n, 7))
    sp[0, 0, 4] = 1
    with pytest.warns(UserWarning):
        obj = api._assemble_object(sd, meta, post, surf, -12.5, sp, np.zeros((chains, 8)), q, options, 10)
    assert obj["returnCode"] == 1 and obj["divergent"] == 1 and len(obj["messages"]) == 2
    options["method"] = "vb"                          # no sampler diagnostics for a variational fit
    surf[0, :ncf] = f_cells
    obj = api._assemble_object(sd, meta, post, surf, -1.0, None, None, None, options, 10)
    assert obj["returnCode"] == 0 and obj["model"]["method"] == "vb"
