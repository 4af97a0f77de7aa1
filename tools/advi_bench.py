#!/usr/bin/env python3
"""Device timing of the full-rank ADVI screen (synergyscreen's default method) on cfg5-shaped synthetic combos.
AB_COMBOS = number of combinations (default 1776), AB_ITER = iteration cap (default 10000)."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bayesynergy_b200 import _lib  # noqa: E402
from bayesynergy_b200.dataprep import prepare, synthetic_experiment  # noqa: E402

n = int(os.environ.get("AB_COMBOS", "1776"))
it = int(os.environ.get("AB_ITER", "10000"))
sds = [prepare(*synthetic_experiment(k)) for k in range(n)]
b = _lib.Batch(sds)
t = time.time()
res = b.advi(seed=1, iter=it)
wall = time.time() - t
info = res["info"]
out = {"combos": n, "iter_cap": it, "kernel_ms": res["kernel_ms"], "wall_s": wall, "n_evals": res["n_evals"],
       "evals_per_s": res["n_evals"] / res["kernel_ms"] * 1e3, "combos_per_s": n / res["kernel_ms"] * 1e3,
       "status_counts": {str(int(s)): int((info[:, 0] == s).sum()) for s in np.unique(info[:, 0])},
       "iterations_median": float(np.median(info[:, 2])), "eta_counts": {str(e): int((info[:, 1] == e).sum()) for e in np.unique(info[:, 1])},
       "converged_bits": {str(int(c)): int((info[:, 4] == c).sum()) for c in np.unique(info[:, 4])}}
print(json.dumps(out))
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "r02_advi_bench.json"), "w"), indent=1)
