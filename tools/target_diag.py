#!/usr/bin/env python3
"""Which combinations do not reach ESS 400, and why?  Runs bsyn_sample_to_ess on N cfg5 combinations and lists the
stragglers with their rank-normalised bulk ESS / R-hat per scalar, divergences and step sizes.
Output: gpurun_out/r02_target_diag.json"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bayesynergy_b200 import _lib  # noqa: E402
from bayesynergy_b200.dataprep import prepare, synthetic_experiment  # noqa: E402

n = int(os.environ.get("N_COMBOS", "1332"))
rounds = int(os.environ.get("MAX_ROUNDS", "13"))
sds = [prepare(*synthetic_experiment(k)) for k in range(n)]
b = _lib.Batch(sds)
r = b.sample_to_ess(target=400.0, block=150, max_rounds=rounds, retry_rounds=0, chains=4, warmup=1000, seed=1000,
                    want_summary=True)
nd, ess = r["draws_per_chain"], r["ess"]
out = dict(n=n, rounds=r["rounds"], kernel_s=r["kernel_ms"] * 1e-3, n_grad=r["n_grad"],
           reached=int((ess >= 400).sum()), draws_hist={int(k): int(v) for k, v in zip(*np.unique(nd, return_counts=True))})
eb_s, rh_s, et_s = b.bulk_ess("VUS_syn")
eb_a, rh_a, et_a = b.bulk_ess("VUS_ant")
cs = r["chain_stats"]
out["stragglers"] = [dict(combo=int(k), ess=float(ess[k]), ess_syn=float(eb_s[k]), rhat_syn=float(rh_s[k]), ess_ant=float(eb_a[k]),
                          rhat_ant=float(rh_a[k]), divergent=float(cs[k, :, 1].sum()), stepsizes=cs[k, :, 0].tolist(),
                          leapfrogs_sampling=cs[k, :, 4].tolist())
                     for k in np.where(~(ess >= 400))[0]]
b.close()
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "r02_target_diag.json"), "w"), indent=1)
print(json.dumps({k: v for k, v in out.items() if k != "stragglers"}))
for s in out["stragglers"]:
    print(s["combo"], round(s["ess"], 1), "syn", round(s["ess_syn"], 1), round(s["rhat_syn"], 3), "ant", round(s["ess_ant"], 1),
          round(s["rhat_ant"], 3), "div", s["divergent"], "eps", np.round(s["stepsizes"], 4))
