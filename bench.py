#!/usr/bin/env python3
"""bench.py -- the measurement contract for the gp_grid NUTS hot path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--combos C]

Metric (BASELINE.json): grad evals/sec (value) and ESS/sec per combo (extra keys) of a synergyscreen batch.
Workload (config.workload): cfg5 of SURVEY.md §8d -- synthetic 10x10-dose-grid x 3-replicate combinations,
gp_grid with the package-default Matern 3/2 kernel and priors, 4 chains x 2000 iterations (1000 warmup) of
adaptive NUTS per combination.  A "step" = one full screen (all warmup + sampling iterations of every chain)
over C combinations per GPU; weak scaling: every rank gets its own C combinations (combination ids are
disjoint), no collective in the data path, one all_reduce of the timings/counters at the end.

  value : whole-job gradient evaluations per second, problems resident in HBM, device time (CUDA events on the
          launch stream inside the library), max over ranks.
  e2e   : the same through the reference-facing call with HOST buffers: bsyn_batch_create (packing + H2D) +
          bsyn_sample with the screenSummary sinks (D2H) + destroy, wall clock around the calls.
  roofline : FP64 issue roofline: achieved = evals/s x FLOP per eval, peak = DFMA microbenchmark
          measured live on the same GPU (bsyn_measure_fp64_peak); HBM figures reported alongside.
  cpu_baseline : the oracle's NUTS ("port": the CPU restatement, NOT rstan) on a bounded sample, all host cores.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# Algorithmic work per gradient evaluation at cfg5 (SURVEY.md §8d / BASELINE.md §3): 26 904 FP64 issue slots
# per lp+gradient + 920 for the leapfrog / tree bookkeeping (8 D), FMA counted as 2 FLOP.
SLOTS_PER_EVAL = 26904 + 920
FLOP_PER_EVAL = 2.0 * SLOTS_PER_EVAL
# Algorithmic HBM bytes per evaluation for the resident design: state lives on chip; per saved draw the kernel
# writes 12 generated scalars + 7 sampler params (152 B) and per combination reads ~5.9 KB of inputs once.
METRIC = "grad_evals_per_sec"


def dist_env():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))


class ClockSampler(threading.Thread):
    def __init__(self, dev):
        super().__init__(daemon=True)
        self.dev, self.samples, self.reasons, self.stop_flag, self.max_mhz = dev, [], set(), False, None

    def run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.dev), f"--query-gpu={q}", "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip().split(",")
                self.samples.append(float(out[0]))
                self.max_mhz = float(out[1])
                for nm, v in zip(names, out[2:]):
                    if "Active" in v and "Not" not in v:
                        self.reasons.add(nm)
            except Exception:
                pass
            time.sleep(0.25)

    def summary(self):
        return {"sm_mhz": float(np.median(self.samples)) if self.samples else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


def make_experiments(first_id, n):
    from bayesynergy_b200.dataprep import prepare, synthetic_experiment
    return [prepare(*synthetic_experiment(first_id + k)) for k in range(n)]


def h2d_bytes(sds):
    tot = 0
    for sd in sds:
        n1, n2, N = sd.n1, sd.n2, sd.N
        ncf = (n1 + 1) * (n2 + 1)
        pp = n1 * (n1 - 1) // 2 + n2 * (n2 - 1) // 2
        tot += 8 * (2 * (n1 + n2) + pp + 3 * ncf + 2 * N) + 4 * (pp + ncf + 1 + N) + 48
    return tot


def cpu_leg(n_combos, chains, iters, warmup, threads, seed=1, first_id=10_000_000):
    """Oracle NUTS (CPU restatement) on a bounded sample; returns dict(value, seconds, n_grad, ess_per_sec)."""
    from oracle.c_oracle import Problem, nuts_screen, lib
    from bayesynergy_b200.dataprep import output_names
    from bayesynergy_b200.diagnostics import split_ess_rhat
    sds = make_experiments(first_id, n_combos)
    probs = [Problem(sd) for sd in sds]
    t = time.time()
    ng, outs, sp = nuts_screen(probs, n_chains=chains, num_warmup=warmup, num_samples=iters - warmup, seed=seed,
                               nthreads=threads)
    dt = time.time() - t
    names = output_names(sds[0])
    ess = []
    for k in range(n_combos):
        e1, _ = split_ess_rhat(outs[k, :, :, names.index("VUS_syn")])
        e2, _ = split_ess_rhat(outs[k, :, :, names.index("VUS_ant")])
        ess.append(min(e1, e2))
    return dict(value=ng / dt, seconds=dt, n_grad=int(ng), ess_per_sec_per_combo=float(np.nanmean(ess) / dt),
                total_ess_per_sec=float(np.nansum(ess) / dt),
                cores=threads if threads > 0 else lib().orc_num_threads())


def run_reference(args):
    rank, _, world = dist_env()
    if rank != 0:
        return
    from oracle.c_oracle import lib
    cores = lib().orc_num_threads()
    n_combos = max(1, cores // args.chains)
    it, wu = args.ref_iter, args.ref_iter // 2
    vals, secs = [], []
    for s in range(args.warmup + args.steps):
        r = cpu_leg(n_combos, args.chains, it, wu, cores, seed=1 + s)
        if s >= args.warmup:
            vals.append(r["value"]); secs.append(r["seconds"])
    v = float(np.mean(vals))
    sample = f"{n_combos} cfg5 combinations x {args.chains} chains x {it} iterations ({wu} warmup) per step"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": "grad evals/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * float(np.mean(secs)),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args, n_combos, it, wu),
        "cpu_baseline": {"value": v, "unit": "grad evals/s", "cores": cores, "kind": "port", "sample": sample,
                         "note": "oracle/gp_oracle.c NUTS: the CPU restatement, NOT rstan (R/rstan are not in the image)",
                         "ess_per_sec_per_combo": r["ess_per_sec_per_combo"], "total_ess_per_sec": r["total_ess_per_sec"]},
        "e2e": {"value": v, "unit": "grad evals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def workload_config(args, combos, it, wu):
    return {"workload": "cfg5: synthetic gp_grid screen, 10x10 dose grid x 3 replicates (N=363, D=115), Matern 3/2, "
                        "default priors, heteroscedastic Normal likelihood, adaptive NUTS (diag_e, adapt_delta 0.9, "
                        "max_treedepth 10)",
            "combos_per_gpu_per_step": combos, "chains": args.chains, "iter": it, "warmup_iter": wu,
            "sharding": "combinations across GPUs, no data-path collective",
            "l2": "256 MiB buffer written between timed steps (L2 flush)"}


def run_ours(args):
    import torch
    rank, local_rank, world = dist_env()
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch with torch.distributed.run --nproc-per-node N for --gpus N")
    dev = local_rank
    torch.cuda.set_device(dev)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", dev))
    from bayesynergy_b200 import _lib
    L = _lib.load()
    it = args.iter
    wu = args.warmup_iter if args.warmup_iter is not None else args.iter // 2
    sds = make_experiments(rank * args.combos, args.combos)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")
    pk = C.c_double()
    L.bsyn_measure_fp64_peak(dev, C.byref(pk))
    b = _lib.Batch(sds, device=dev)
    launches0 = L.bsyn_launch_count()
    # ---- device-resident leg -------------------------------------------------------------------
    for s in range(args.warmup):
        b.sample_device(chains=args.chains, iter=it, warmup=wu, seed=100 + s)
        flush.fill_(s & 255)
    clocks = ClockSampler(dev)
    clocks.start()
    barrier()
    launches1 = L.bsyn_launch_count()
    tot_ms, tot_grad = 0.0, 0
    t0 = time.time()
    for s in range(args.steps):
        ng, ms = b.sample_device(chains=args.chains, iter=it, warmup=wu, seed=1000 + s)
        tot_ms += ms
        tot_grad += ng
        flush.fill_(s & 255)
    barrier()
    wall = time.time() - t0
    clocks.stop_flag = True
    launches = L.bsyn_launch_count() - launches1
    ess_syn, _ = b.ess("VUS_syn")
    ess_ant, rhat = b.ess("VUS_ant")
    ess = np.fmin(ess_syn, ess_ant)
    step_s = tot_ms / args.steps * 1e-3
    # ---- end-to-end leg: host buffers in, summaries out, through the public call --------------------
    b.close()
    e2e_t, e2e_grad = [], []
    for s in range(1):
        barrier()
        t = time.time()
        bb = _lib.Batch(sds, device=dev)
        res = bb.sample(chains=args.chains, iter=it, warmup=wu, seed=2000 + s, want_qdraws=False, want_sampler_params=False)
        bb.close()
        torch.cuda.synchronize()
        e2e_t.append(time.time() - t)
        e2e_grad.append(res["n_grad"])
    d2h = 8 * (args.combos * 12 * 2 + args.combos * args.chains * 8)
    # ---- reduce over ranks -------------------------------------------------------------------------
    stats = torch.tensor([tot_ms, float(tot_grad), wall, float(np.nansum(ess)), float(np.sum(e2e_grad)),
                          float(np.sum(e2e_t)), float(np.isnan(ess).sum()), float(launches),
                          float((ess >= 400).sum())], dtype=torch.float64, device="cuda")
    mx = stats.clone()
    sm = stats.clone()
    if world > 1:
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        dist.all_reduce(sm, op=dist.ReduceOp.SUM)
    mx, sm = mx.cpu().numpy(), sm.cpu().numpy()
    if rank == 0:
        dev_s = mx[0] * 1e-3
        value = sm[1] / dev_s
        e2e_value = sm[4] / mx[5]
        n_combo_total = args.combos * world
        step_time = dev_s / args.steps
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        achieved_tflops = value / world * FLOP_PER_EVAL / 1e12
        peak_tflops = 2.0 * pk.value / 1e12
        out = {
            "metric": METRIC, "value": value, "unit": "grad evals/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * step_time, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(args, args.combos, it, wu),
            "ess_per_sec_per_combo": float(sm[3] / n_combo_total / step_time),
            "total_ess_per_sec": float(sm[3] / step_time),
            "ess_min_syn_ant_mean": float(sm[3] / n_combo_total), "ess_nan_combos": int(sm[6]),
            # time until every combination of the step has ESS >= 400 -- the step itself when all do
            "ess_ge_400_frac": float(sm[8] / n_combo_total),
            "time_to_ess400_all_s": float(step_time) if int(sm[8]) == n_combo_total else None,
            "grad_evals_per_step": sm[1] / args.steps,
            "wall_s_timed_region": float(mx[2]),
            "e2e": {"value": e2e_value, "unit": "grad evals/s", "h2d_bytes_per_step": h2d_bytes(sds),
                    "d2h_bytes_per_step": d2h, "seconds_per_step": float(mx[5] / len(e2e_t))},
            "gpu_launches": int(mx[7]),
            "clocks": clocks.summary(),
            # the bound of this path is FP64 issue throughput (neither HBM nor tensor cores)
            "roofline": {"bound": "fp64", "achieved": achieved_tflops, "peak": peak_tflops, "unit": "TFLOP/s",
                         "frac": achieved_tflops / peak_tflops if peak_tflops > 0 else None,
                         "peak_source": "bsyn_measure_fp64_peak: register-resident DFMA microbenchmark, measured in this run "
                                        "(MEASURED_PEAKS.json has no FP64 figure)",
                         "flop_per_eval": FLOP_PER_EVAL, "per_gpu": True,
                         "hbm": {"peak_gbs": peaks.get("hbm_gbs"), "note": "state is on-chip; HBM is not the bound"},
                         # DRAM bytes of one sampler launch: 13.3 B per gradient evaluation (dram__bytes_read.sum +
                         # dram__bytes_write.sum of the ncu --set full capture in profiles/r01_nuts_v6_summary.txt)
                         # x the evaluations of one launch; the algorithmic figure is ~0 (state is resident)
                         "traffic": 13.3 * sm[1] / args.steps / world,
                         "traffic_unit": "bytes per launch (extrapolated from a 444-combination ncu capture)"},
        }
        if world == 1 and not args.no_cpu:
            from oracle.c_oracle import lib
            cores = lib().orc_num_threads()
            nc = max(1, cores // args.chains)
            r = cpu_leg(nc, args.chains, args.ref_iter, args.ref_iter // 2, cores)
            out["cpu_baseline"] = {"value": r["value"], "unit": "grad evals/s", "cores": cores, "kind": "port",
                                   "sample": f"{nc} cfg5 combinations x {args.chains} chains x {args.ref_iter} iterations "
                                             f"({args.ref_iter // 2} warmup), {r['seconds']:.1f} s",
                                   "note": "oracle/gp_oracle.c NUTS: CPU restatement (not rstan)",
                                   "ess_per_sec_per_combo": r["ess_per_sec_per_combo"],
                                   "total_ess_per_sec": r["total_ess_per_sec"]}
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--combos", type=int, default=int(os.environ.get("BENCH_COMBOS", "1776")),
                    help="combinations per GPU per step (the full cfg5 screen is 12500 per GPU on 8 GPUs; 1776 = 4 chains "
                         "per resident warp slot keeps a step at ~30 s so that W + K steps finish within minutes)")
    ap.add_argument("--chains", type=int, default=4)
    ap.add_argument("--iter", type=int, default=2000)
    ap.add_argument("--warmup-iter", type=int, default=None,
                    help="warm-up iterations per chain (default iter / 2 as in rstan); e.g. --iter 1150 --warmup-iter 1000 "
                         "samples just long enough for ESS >= 400 per combination")
    ap.add_argument("--ref-iter", type=int, default=600, help="iterations per chain in the bounded CPU sample")
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
