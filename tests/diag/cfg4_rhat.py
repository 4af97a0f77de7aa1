#!/usr/bin/env python3
"""Root cause of the cfg4 R-hat outliers: the nointeraction model (H0 of the Bayes factor)
fitted to cfg5 data showed max R-hat 1.95.  Is it the sampler with model = 1, or the posterior?

The device sampler runs the nointeraction model on N cfg5 combinations; the 10 combinations with the largest
R-hat (max over rVUS_p0, s, dss_1, dss_2, ec50_1, ec50_2) are then refitted by the ORACLE's NUTS (oracle/gp_oracle.c,
recursive Stan-style sampler, independent code) with two seeds, and by the device with a second seed.  Per
combination the script records R-hat from each, the per-chain means of `s` and `rVUS_p0` (modes), the data-generating
interaction strength, and the lp__ level of each chain.  Output: gpurun_out/r02_cfg4_rhat.json.
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from bayesynergy_b200 import _lib  # noqa: E402
from bayesynergy_b200.dataprep import prepare, synthetic_experiment, output_names  # noqa: E402
from bayesynergy_b200.diagnostics import split_ess_rhat  # noqa: E402
from oracle.c_oracle import Problem, nuts_screen  # noqa: E402

QS = ("rVUS_p0", "s", "dss_1", "dss_2", "ec50_1", "ec50_2")


def strength(cid):
    rng = np.random.default_rng(20261019 + cid)
    rng.beta(1.0, 1.25, size=2); rng.gamma(2.0, 1.0, size=2); rng.normal(-1.0, 1.0, size=2)
    return float(rng.choice([0.0, 0.0, 0.5, 1.5]))


def rhat_q(q):      # q: [chains, draws, NQ] (device layout)
    return {n: split_ess_rhat(q[:, :, _lib.Q_NAMES.index(n)])[1] for n in QS}


def main():
    n = int(os.environ.get("N_COMBOS", "1776"))
    sds = [prepare(*synthetic_experiment(k)) for k in range(n)]
    b = _lib.Batch(sds, model="nointeraction")
    res = b.sample(chains=4, iter=2000, warmup=1000, seed=1)
    b.close()
    rh = np.array([max(rhat_q(res["qdraws"][k]).values()) for k in range(n)])
    worst = [int(k) for k in np.argsort(-np.nan_to_num(rh))[:10]]
    out = dict(n_combos=n, device_rhat_max=float(np.nanmax(rh)), device_n_rhat_gt_1_05=int((rh > 1.05).sum()),
               device_n_rhat_gt_1_5=int((rh > 1.5).sum()), worst=[])
    sub = [sds[k] for k in worst]
    b2 = _lib.Batch(sub, model="nointeraction")
    res2 = b2.sample(chains=4, iter=2000, warmup=1000, seed=77)
    b2.close()
    probs = [Problem(sd, "nointeraction") for sd in sub]
    names = output_names(sub[0], "nointeraction")
    orc = []
    for seed in (5, 6):
        ng, outs, sp = nuts_screen(probs, n_chains=4, num_warmup=1000, num_samples=1000, seed=seed)
        orc.append((outs, sp))
    for j, k in enumerate(worst):
        row = dict(combo=k, interaction_strength_of_the_data=strength(k),
                   device_seed1=dict(rhat=rhat_q(res["qdraws"][k]),
                                     chain_means_s=res["qdraws"][k][:, :, _lib.Q_NAMES.index("s")].mean(axis=1).tolist(),
                                     chain_means_rVUS_p0=res["qdraws"][k][:, :, _lib.Q_NAMES.index("rVUS_p0")].mean(axis=1).tolist(),
                                     chain_means_lp=res["sampler_params"][k][:, :, 6].mean(axis=1).tolist(),
                                     divergent=int(res["sampler_params"][k][:, :, 4].sum())),
                   device_seed77=dict(rhat=rhat_q(res2["qdraws"][j]),
                                      chain_means_s=res2["qdraws"][j][:, :, _lib.Q_NAMES.index("s")].mean(axis=1).tolist(),
                                      chain_means_lp=res2["sampler_params"][j][:, :, 6].mean(axis=1).tolist()))
        for (outs, sp), seed in zip(orc, (5, 6)):
            row[f"oracle_seed{seed}"] = dict(
                rhat={q: split_ess_rhat(outs[j][:, :, names.index(q)])[1] for q in QS},
                chain_means_s=outs[j][:, :, names.index("s")].mean(axis=1).tolist(),
                chain_means_rVUS_p0=outs[j][:, :, names.index("rVUS_p0")].mean(axis=1).tolist(),
                chain_means_lp=sp[j][:, :, 6].mean(axis=1).tolist(), divergent=int(sp[j][:, :, 4].sum()))
        out["worst"].append(row)
    # the same statistic for the oracle on the first 240 combinations (what fraction of H0 posteriors is hard at all)
    m = min(n, int(os.environ.get("N_ORACLE", "240")))
    ng, outs, sp = nuts_screen([Problem(sd, "nointeraction") for sd in sds[:m]], n_chains=4, num_warmup=1000,
                               num_samples=1000, seed=5)
    rho = np.array([max(split_ess_rhat(outs[k][:, :, names.index(q)])[1] for q in QS) for k in range(m)])
    out["oracle_first"] = dict(n=m, rhat_max=float(np.nanmax(rho)), n_rhat_gt_1_05=int((rho > 1.05).sum()),
                               n_rhat_gt_1_5=int((rho > 1.5).sum()))
    out["device_first"] = dict(n=m, rhat_max=float(np.nanmax(rh[:m])), n_rhat_gt_1_05=int((rh[:m] > 1.05).sum()),
                               n_rhat_gt_1_5=int((rh[:m] > 1.5).sum()))
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", "r02_cfg4_rhat.json"), "w"), indent=1)
    print(json.dumps({k: v for k, v in out.items() if k != "worst"}))
    for r in out["worst"]:
        print(r["combo"], r["interaction_strength_of_the_data"], "device", max(r["device_seed1"]["rhat"].values()),
              max(r["device_seed77"]["rhat"].values()), "oracle", max(r["oracle_seed5"]["rhat"].values()),
              max(r["oracle_seed6"]["rhat"].values()))


if __name__ == "__main__":
    main()
