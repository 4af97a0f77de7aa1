#!/usr/bin/env python3
"""cfg4 of BASELINE.json: heteroscedastic gp_grid + nointeraction fits and the bridge-sampling Bayes factor per combination
("two models per combo"), as synergyscreen(bayes_factor = TRUE) runs it: one gp_grid launch + one nointeraction launch, the
bridge estimator on the device.  Reports wall time per stage.  BF_COMBOS = combinations (default 444)."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bayesynergy_b200 import _lib  # noqa: E402
from bayesynergy_b200.dataprep import prepare, synthetic_experiment  # noqa: E402

n = int(os.environ.get("BF_COMBOS", "444"))
it = int(os.environ.get("BF_ITER", "2000"))
sds = [prepare(*synthetic_experiment(k)) for k in range(n)]
out = dict(combos=n, iter=it, chains=4)
lml = {}
for model in ("gp_grid", "nointeraction"):
    b = _lib.Batch(sds, model=model)
    b.sample_device(chains=4, iter=100, warmup=50, seed=9)           # warm the kernels
    t = time.time()
    r0 = b.sample(chains=4, iter=it, warmup=it // 2, seed=1, want_qdraws=False, want_sampler_params=False)
    t_plain = time.time() - t
    t = time.time()
    r = b.sample(chains=4, iter=it, warmup=it // 2, seed=1, want_bridge=True, want_qdraws=False, want_sampler_params=False)
    t_bridge = time.time() - t
    b.close()
    lml[model] = r["bridge_logml"]
    out[model] = dict(sample_s=t_plain, sample_plus_bridge_s=t_bridge, bridge_s=t_bridge - t_plain, n_grad=r["n_grad"],
                      bridge_iterations_median=float(np.median(r["bridge_niter"])), logml_finite=int(np.isfinite(r["bridge_logml"]).sum()))
bf = np.exp(lml["gp_grid"] - lml["nointeraction"])
out["log_bf_quantiles"] = np.nanquantile(np.log(bf), [0.05, 0.5, 0.95]).tolist()
out["total_s"] = out["gp_grid"]["sample_plus_bridge_s"] + out["nointeraction"]["sample_plus_bridge_s"]
out["bayes_factors_per_s"] = n / out["total_s"]
print(json.dumps(out))
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "r02_bf_bench.json"), "w"), indent=1)
