#!/usr/bin/env python3
"""Tiny driver for ncu: a few launches of the stand-alone lp+grad kernel (cfg5-shaped problems) and,
optionally, one short NUTS launch.  PROF_B = points per launch, PROF_NUTS=1 adds the sampler."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bayesynergy_b200 import _lib  # noqa: E402
from bayesynergy_b200.dataprep import prepare, synthetic_experiment  # noqa: E402

B = int(os.environ.get("PROF_B", "37888"))
nprob = int(os.environ.get("PROF_NPROB", "148"))
sds = [prepare(*synthetic_experiment(k)) for k in range(nprob)]
b = _lib.Batch(sds)
rng = np.random.default_rng(0)
U = rng.uniform(-1, 1, size=(B, b.max_params))
idx = (np.arange(B) * nprob // B).astype(np.int32)
for rep in range(3):
    lp, g, st = b.logp_grad(U, prob_idx=idx)
print("logp ok", int((st != 0).sum()), float(lp[:4].sum()))
if os.environ.get("PROF_NUTS"):
    it = int(os.environ.get("PROF_ITER", "60"))
    ng, ms = b.sample_device(chains=4, iter=it, warmup=it // 2, seed=1)
    print("nuts", ng, ms, "evals/s %.3e" % (ng / ms * 1e3))
