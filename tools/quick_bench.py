#!/usr/bin/env python3
"""Development micro-bench (not the contract bench.py): FP64 peak, lp+grad kernel rate, short NUTS screen."""
import ctypes as C
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bayesynergy_b200 import _lib  # noqa: E402
from bayesynergy_b200.dataprep import prepare, synthetic_experiment  # noqa: E402


def main():
    L = _lib.load()
    pk = C.c_double()
    rc = L.bsyn_measure_fp64_peak(0, C.byref(pk))
    print("fp64 peak: rc", rc, "%.3e FMA/s = %.2f TFLOP/s" % (pk.value, 2 * pk.value / 1e12), flush=True)
    n_prob = int(os.environ.get("QB_NPROB", "1184"))
    sds = [prepare(*synthetic_experiment(k)) for k in range(n_prob)]
    b = _lib.Batch(sds)
    D = b.max_params
    if os.environ.get("QB_NUTS_ONLY"):
        for (it, wu) in ((int(os.environ.get("QB_ITER", "200")), int(os.environ.get("QB_ITER", "200")) // 2),):
            ng, ms = b.sample_device(chains=4, iter=it, warmup=wu, seed=1)
            print(f"NUTS {n_prob} combos x 4 chains, iter={it} warmup={wu}: {ng} grad evals, kernel {ms:.1f} ms, "
                  f"{ng/ms*1e3:.3e} evals/s ({ng/ms*1e3*27824/pk.value:.3f} of FP64 peak)", flush=True)
        return
    rng = np.random.default_rng(0)
    B = 400_000
    U = rng.uniform(-1, 1, size=(B, D))
    idx = (np.arange(B) // (B // n_prob + 1)).astype(np.int32)
    for rep in range(3):
        t = time.time()
        lp, g, st = b.logp_grad(U, prob_idx=idx)
        dt = time.time() - t
        print(f"logp_grad host path: B={B} {dt*1e3:.1f} ms  {B/dt:.3e} evals/s (incl. H2D/D2H)  bad={int((st!=0).sum())}", flush=True)
    # device-resident timing via torch tensors
    import torch
    dU = torch.from_numpy(U).cuda()
    didx = torch.from_numpy(idx).cuda()
    dlp = torch.empty(B, dtype=torch.float64, device="cuda")
    dg = torch.empty((B, D), dtype=torch.float64, device="cuda")
    dst = torch.empty(B, dtype=torch.int32, device="cuda")
    st_ = torch.cuda.current_stream().cuda_stream
    for rep in range(4):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        rc = L.bsyn_logp_grad_device(b.h, B, didx.data_ptr(), dU.data_ptr(), D, 1, dlp.data_ptr(), dg.data_ptr(),
                                     dst.data_ptr(), st_)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        print(f"logp_grad device: rc={rc} {ms:.2f} ms  {B/ms*1e3:.3e} evals/s  -> {B/ms*1e3*26904/ (pk.value):.3f} of measured FP64 issue peak (26904 slots/eval)", flush=True)
    for (it, wu) in ((200, 100), (400, 200)):
        t = time.time()
        ng, ms = b.sample_device(chains=4, iter=it, warmup=wu, seed=1)
        dt = time.time() - t
        print(f"NUTS {n_prob} combos x 4 chains, iter={it} warmup={wu}: {ng} grad evals, kernel {ms:.1f} ms, "
              f"{ng/ms*1e3:.3e} evals/s ({ng/ms*1e3*27824/pk.value:.3f} of FP64 peak), wall {dt:.2f}s", flush=True)
        ess, rhat = b.ess("VUS_syn")
        print("   ESS(VUS_syn) median %.1f  rhat median %.3f  nan %d" % (np.nanmedian(ess), np.nanmedian(rhat), int(np.isnan(ess).sum())), flush=True)
    res = b.sample(chains=4, iter=400, warmup=200, seed=1, want_qdraws=False, want_sampler_params=True)
    sp = res["sampler_params"]
    cs = res["chain_stats"]
    print("treedepth mean %.2f  n_leapfrog mean %.1f  accept %.3f  divergent %d  init_failed %d  stepsize median %.4f" % (
        sp[..., 2].mean(), sp[..., 3].mean(), sp[..., 0].mean(), int(sp[..., 4].sum()), int(cs[..., 6].sum()), np.median(cs[..., 0])), flush=True)
    print("summary VUS_syn mean of means %.3f ; rVUS_f %.3f" % (res["summary"][:, 8, 0].mean(), res["summary"][:, 5, 0].mean()))
    print("launches:", L.bsyn_launch_count())


if __name__ == "__main__":
    main()
