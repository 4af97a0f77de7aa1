#!/usr/bin/env python3
"""Measured rates for every BASELINE.json config (the bench line is cfg5; these are documentation).

cfg1  bayesynergy() on mathews_DLBCL[[1]], 4 chains x 2000 iter: a batch of identical copies (different seeds)
cfg1' same with the RBF kernel (type = 2)
cfg2  O'Neil-shaped screen (11x11 grid, 32 observed cells, N = 256), default options
cfg3  robust + PC-prior Matern + lower asymptotes, dense_e + stepsize_jitter 0.5 (what bayesynergy() selects)
cfg4  heteroscedastic gp_grid + the nointeraction model (Bayes-factor pair) on the cfg5 data
cfg5  synthetic 10x10x3 screen (the bench line)
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bayesynergy_b200 import _lib  # noqa: E402
from bayesynergy_b200.dataprep import prepare, synthetic_experiment, synthetic_oneil_experiment  # noqa: E402


def run(name, sds, model="gp_grid", iters=int(os.environ.get("CB_ITER", "2000")), **kw):
    b = _lib.Batch(sds, model=model)
    b.sample_device(chains=4, iter=200, warmup=100, seed=9, **kw)     # warm-up launch
    ng, ms = b.sample_device(chains=4, iter=iters, warmup=iters // 2, seed=1, **kw)
    q1, q2 = ("VUS_syn", "VUS_ant") if model == "gp_grid" else ("rVUS_f", "rVUS_p0")
    ess_syn, rhat1, _ = b.bulk_ess(q1)          # rank-normalised bulk ESS / R-hat
    ess_ant, rhat2, _ = b.bulk_ess(q2)
    ess = np.fmin(ess_syn, ess_ant)
    rhat = np.fmax(rhat1, rhat2)
    out = dict(config=name, model=model, combos=len(sds), D=b.max_params, n_grad=ng, kernel_s=ms * 1e-3,
               grad_evals_per_s=ng / ms * 1e3, leapfrogs_per_iter=ng / (len(sds) * 4 * iters),
               ess_min_mean=float(np.nanmean(ess)), total_ess_per_s=float(np.nansum(ess) / (ms * 1e-3)),
               ess_ge_400_frac=float(np.mean(ess >= 400)), rhat_max=float(np.nanmax(rhat)),
               rhat_gt_1_05=int((rhat > 1.05).sum()), sampler=kw, iter=iters, chains=4,
               ess_estimator="rank-normalised bulk ESS, min over " + q1 + " / " + q2)
    b.close()
    print(json.dumps(out), flush=True)
    return out


def main():
    nb = int(os.environ.get("CB_COMBOS", "888"))
    m = json.load(open(os.path.join(ROOT, "tests/golden/mathews_DLBCL.json")))
    e = m["ispinesib + ibrutinib"]
    x = np.array(e["x"]).reshape(2, -1).T
    y = np.array(e["y"])
    res = []
    only = os.environ.get("CB_ONLY", "")          # e.g. CB_ONLY=cfg1 : just the configurations whose name starts with it

    def go(name, *a, **k):
        if only and not name.startswith(only):
            return dict(config=name, skipped=True)
        return run(name, *a, **k)
    res.append(go("cfg1 mathews_DLBCL[[1]] Matern 3/2 (package default)", [prepare(y, x)] * nb))
    res.append(go("cfg1' mathews_DLBCL[[1]] RBF (type=2)", [prepare(y, x, type=2)] * nb))
    on = [prepare(*synthetic_oneil_experiment(k)) for k in range(min(nb, 583))]
    res.append(go("cfg2 O'Neil-shaped screen, default options", on))
    onr = [prepare(*synthetic_oneil_experiment(k), robust=True, pcprior=True) for k in range(min(nb, 583))]
    res.append(go("cfg3 O'Neil-shaped, robust + PC prior, dense_e + jitter 0.5", onr, metric=1, stepsize_jitter=0.5))
    res.append(go("cfg3-diag same model with diag_e (for comparison)", onr, metric=0, stepsize_jitter=0.5))
    syn = [prepare(*synthetic_experiment(k)) for k in range(nb)]
    res.append(go("cfg4 nointeraction (H0 of the Bayes factor) on cfg5 data", syn, model="nointeraction"))
    res.append(go("cfg5 synthetic 10x10x3 (bench line)", syn))
    json.dump(res, open(os.path.join(ROOT, "gpurun_out", "r02_config_bench.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
