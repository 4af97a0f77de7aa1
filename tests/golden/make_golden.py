#!/usr/bin/env python3
"""Generate tests/golden/logp_golden.json: lp and gradient of the gp_grid / nointeraction log-density at
seeded points, from the torch-float64 AUTOGRAD restatement (oracle/torch_oracle.py), for the shipped
datasets and a synthetic 10x10x3 experiment, over kernels x likelihoods x options.

These are SELF-DERIVED golden vectors (the reference has none: SURVEY.md §4/§8c; parity unpinned).  They pin
(i) the C oracle's hand adjoints against autograd and (ii) the CUDA path against both.

    python tests/golden/make_golden.py
"""
import itertools
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from bayesynergy_b200.dataprep import prepare, synthetic_experiment  # noqa: E402
from oracle.torch_oracle import log_prob_grad, DomainError, _kernel_mats  # noqa: E402
import torch  # noqa: E402


def kernel_cond(sd, u, model):
    """max condition number of the two kernel matrices at u (1 for nointeraction)"""
    if model != "gp_grid":
        return 1.0
    k0 = 2 * sd.est_la + 6
    ell, sigf = np.exp(u[k0 + 2]), np.exp(u[k0 + 3])
    alpha = torch.tensor(np.exp(u[k0 + 4])) if sd.est_alpha else None
    c1, c2 = _kernel_mats(sd, torch.tensor(ell), torch.tensor(sigf), alpha)
    return float(max(np.linalg.cond(c1.numpy()), np.linalg.cond(c2.numpy())))


def datasets():
    m = json.load(open(os.path.join(ROOT, "tests/golden/mathews_DLBCL.json")))
    out = {}
    for k, e in m.items():
        out["mathews:" + k] = (np.array(e["y"]), np.array(e["x"]).reshape(2, -1).T)
    o = json.load(open(os.path.join(ROOT, "tests/golden/ONeil_A375_failed.json")))
    for i, e in enumerate(o):
        out[f"oneil_failed:{i}"] = (np.array(e["y"]), np.array(e["x"]).reshape(2, -1).T)
    ys, xs = synthetic_experiment(7)
    out["synthetic:7"] = (ys, xs)
    return out


def main():
    rng = np.random.default_rng(20261019)
    cases = []
    variants = []
    for typ, nu in [(2, 1.5), (3, 0.5), (3, 1.5), (3, 2.5), (4, 1.5)]:
        for het, rob in itertools.product([1, 0], [0, 1]):
            variants.append(dict(type=typ, nu=nu, heteroscedastic=het, robust=rob, lower_asymptotes=1, pcprior=0))
    variants.append(dict(type=3, nu=1.5, heteroscedastic=1, robust=1, lower_asymptotes=1, pcprior=1))
    variants.append(dict(type=3, nu=1.5, heteroscedastic=1, robust=0, lower_asymptotes=0, pcprior=0))
    variants.append(dict(type=3, nu=2.5, heteroscedastic=0, robust=0, lower_asymptotes=0, pcprior=1))
    for dname, (y, x) in datasets().items():
        big = dname.startswith("oneil") or dname.startswith("synthetic")
        for vi, v in enumerate(variants):
            if big and vi % 4 not in (0, 1) and vi < 20:
                continue   # keep the fixture small: the large grids get a subset of the variants
            sd = prepare(y, x, **v)
            for model in ("gp_grid", "nointeraction"):
                if model == "nointeraction" and (v["type"] != 3 or v["nu"] != 1.5):
                    continue   # nointeraction has no kernel: one copy per likelihood is enough
                D = sd.dim_gp_grid if model == "gp_grid" else sd.dim_nointeraction
                for t in range(2):
                    u = rng.uniform(-1.0, 1.0, D)
                    if t == 1:
                        u[D - 3] = -2.5   # small noise scale: large residuals, LPTN tails
                    try:
                        lp, g = log_prob_grad(sd, u, model=model)
                    except DomainError:
                        continue
                    cases.append(dict(dataset=dname, variant=v, model=model, u=u.tolist(), lp=lp, grad=g.tolist(),
                                      cond=kernel_cond(sd, u, model)))
    # SURVEY.md Appendix E probe point
    y, x = datasets()["mathews:ispinesib + ibrutinib"]
    u = (0.3 * np.sin(np.arange(40) + 1.0))
    sd = prepare(y, x)
    lp, g = log_prob_grad(sd, u)
    cases.append(dict(dataset="mathews:ispinesib + ibrutinib", variant=dict(type=3, nu=1.5, heteroscedastic=1, robust=0,
                      lower_asymptotes=1, pcprior=0), model="gp_grid", u=u.tolist(), lp=lp, grad=g.tolist(),
                      cond=kernel_cond(sd, u, "gp_grid"),
                      note="SURVEY.md Appendix E: lp = -61.1325359718674"))
    with open(os.path.join(ROOT, "tests/golden/logp_golden.json"), "w") as f:
        json.dump(dict(generator="tests/golden/make_golden.py (torch float64 autograd)", cases=cases), f)
    print(len(cases), "cases")


if __name__ == "__main__":
    main()
