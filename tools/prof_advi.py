#!/usr/bin/env python3
"""Tiny driver for ncu: one full-rank ADVI launch on cfg5-shaped problems (PROF_NPROB experiments, PROF_ITER cap)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bayesynergy_b200 import _lib  # noqa: E402
from bayesynergy_b200.dataprep import prepare, synthetic_experiment  # noqa: E402

n = int(os.environ.get("PROF_NPROB", "1776"))
it = int(os.environ.get("PROF_ITER", "400"))
b = _lib.Batch([prepare(*synthetic_experiment(k)) for k in range(n)])
res = b.advi(seed=1, iter=it, output_samples=100)
print("advi", n, it, res["n_evals"], res["kernel_ms"], "evals/s %.3e" % (res["n_evals"] / res["kernel_ms"] * 1e3))
