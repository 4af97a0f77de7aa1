#!/usr/bin/env python3
"""bench.py -- the measurement contract for the gp_grid NUTS hot path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--combos C] [--workload screen|target]

Metric (BASELINE.json): grad evals/sec (value) and ESS/sec per combo (extra keys) of a synergyscreen batch.
Workload (config.workload): cfg5 of SURVEY.md §8d -- synthetic 10x10-dose-grid x 3-replicate combinations,
gp_grid with the package-default Matern 3/2 kernel and priors, 4 chains of adaptive NUTS per combination with
rstan's 1000 warm-up iterations + 150 sampling iterations (the budget that takes a combination to north_star's
ESS of 400: `--iter 1150 --warmup-iter 1000`; `--iter 2000` is rstan's default budget).  A "step" = one full
screen (all warm-up + sampling iterations of every chain) over C combinations per GPU; weak scaling: every rank
gets its own C combinations (combination ids are disjoint), no collective in the data path, one all_reduce of the
timings/counters at the end.  The W warm-up steps run the same screen at 1/5 of the iterations (declared in
config.warmup_steps): they warm the instruction caches, the allocator and the clocks, not a cache of results.
Both arms (`--impl ours|reference`) print the SAME config; the reference arm's bounded sample is in cpu_baseline.

  value : whole-job gradient evaluations per second, problems resident in HBM, device time (CUDA events on the
          launch stream inside the library), max over ranks.
  e2e   : the same through the reference-facing call with HOST buffers: bsyn_batch_create (packing + H2D) +
          bsyn_sample with the screenSummary sinks (D2H) + destroy, wall clock around the calls.
  roofline : FP64 issue roofline: achieved = evals/s x FLOP per eval, peak = DFMA microbenchmark
          measured live on the same GPU (bsyn_measure_fp64_peak); HBM figures reported alongside.
  cpu_baseline : the oracle's NUTS ("port": the CPU restatement, NOT rstan) on a bounded sample, all host cores.

`--workload target` (north_star's target run): every combination is sampled until its rank-normalised bulk ESS of
VUS_syn and VUS_ant reaches 400 -- one warm-up + sampling launch, then extension launches over the combinations
still below the target (chains resume from their adapted state); reports time_to_ess400_all_s.
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
METRIC = "grad_evals_per_sec"
ESS_TARGET = 400.0


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


def iters(args):
    it = args.iter
    wu = args.warmup_iter if args.warmup_iter is not None else args.iter // 2
    return it, wu


def workload_config(args):
    """The SAME dict in both arms: it names the workload, not how large a sample of it an arm timed."""
    it, wu = iters(args)
    cfg = {"workload": "cfg5: synthetic gp_grid screen, 10x10 dose grid x 3 replicates (N=363, D=115), Matern 3/2, "
                       "default priors, heteroscedastic Normal likelihood, adaptive NUTS (diag_e, adapt_delta 0.9, "
                       "max_treedepth 10)"
                       + ("; every combination sampled until min(bulk-ESS(VUS_syn), bulk-ESS(VUS_ant)) >= 400"
                          if args.workload == "target" else ""),
           "mode": args.workload,
           "combos_per_gpu_per_step": args.combos, "chains": args.chains, "iter": it, "warmup_iter": wu,
           "warmup_steps": "same screen at 1/5 of the iterations",
           "ess_estimator": "rank-normalised split-chain bulk ESS (Vehtari et al. 2021), min over VUS_syn / VUS_ant",
           "sharding": "combinations across GPUs, no data-path collective",
           "l2": "256 MiB buffer written between timed steps (L2 flush)"}
    if args.workload == "target":
        cfg["phase2"] = getattr(args, "retry", "refit")
    return cfg


def cpu_leg(n_combos, chains, it, wu, threads, seed=1, first_id=10_000_000):
    """Oracle NUTS (CPU restatement) on a bounded sample; returns dict(value, seconds, n_grad, ess_per_sec)."""
    from oracle.c_oracle import Problem, nuts_screen, lib
    from bayesynergy_b200.dataprep import output_names
    from bayesynergy_b200.diagnostics import bulk_ess
    sds = make_experiments(first_id, n_combos)
    probs = [Problem(sd) for sd in sds]
    t = time.time()
    ng, outs, sp = nuts_screen(probs, n_chains=chains, num_warmup=wu, num_samples=it - wu, seed=seed,
                               nthreads=threads)
    dt = time.time() - t
    names = output_names(sds[0])
    ess = []
    for k in range(n_combos):
        e1 = bulk_ess(outs[k, :, :, names.index("VUS_syn")])
        e2 = bulk_ess(outs[k, :, :, names.index("VUS_ant")])
        ess.append(min(e1, e2))
    return dict(value=ng / dt, seconds=dt, n_grad=int(ng), ess_per_sec_per_combo=float(np.nanmean(ess) / dt),
                total_ess_per_sec=float(np.nansum(ess) / dt), ess_ge_400_frac=float(np.mean(np.array(ess) >= ESS_TARGET)),
                cores=threads if threads > 0 else lib().orc_num_threads())


def run_reference(args):
    rank, _, world = dist_env()
    if rank != 0:
        return
    from oracle.c_oracle import lib
    cores = lib().orc_num_threads()
    n_combos = max(1, cores // args.chains)      # one chain per host thread
    it, wu = iters(args)
    vals, secs, last = [], [], None
    for s in range(args.warmup + args.steps):
        warm = s < args.warmup
        r = cpu_leg(n_combos, args.chains, max(10, it // 5) if warm else it, max(5, wu // 5) if warm else wu, cores,
                    seed=1 + s, first_id=10_000_000 + s * n_combos)
        if not warm:
            vals.append(r["value"]); secs.append(r["seconds"]); last = r
    v = float(np.sum([x * y for x, y in zip(vals, secs)]) / np.sum(secs))     # total evals / total seconds
    sample = (f"{n_combos} cfg5 combinations x {args.chains} chains x {it} iterations ({wu} warm-up) per step = one chain "
              f"per host thread; {args.steps} steps, {np.sum(secs):.1f} s")
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": "grad evals/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * float(np.mean(secs)),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args),
        "cpu_baseline": {"value": v, "unit": "grad evals/s", "cores": cores, "kind": "port", "sample": sample,
                         "note": "oracle/gp_oracle.c NUTS: the CPU restatement, NOT rstan (R/rstan are not in the image)"
                                 + ("; --workload target: the CPU arm runs the first block only (warm-up + 150 draws per chain, "
                                    "no extension launches): its ess_ge_400_frac says how far that gets" if args.workload == "target" else ""),
                         "ess_per_sec_per_combo": last["ess_per_sec_per_combo"], "total_ess_per_sec": last["total_ess_per_sec"],
                         "ess_ge_400_frac": last["ess_ge_400_frac"]},
        "e2e": {"value": v, "unit": "grad evals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def measured_traffic(args):
    """DRAM bytes of ONE sampler launch from the committed ncu capture of this command (profiles/r02_traffic.json,
    written by tools/ncu_traffic.py from `ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum`): used only when
    the capture is of the same workload (combos, iter, warm-up), otherwise null -- never extrapolated."""
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", "r02_traffic.json")))
    except Exception:
        return None, None
    it, wu = iters(args)
    if (t.get("combos"), t.get("iter"), t.get("warmup_iter"), t.get("chains")) != (args.combos, it, wu, args.chains):
        return None, None
    return float(t["dram_bytes_per_launch"]), t.get("source")


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
    from bayesynergy_b200.api import screen_to_ess
    L = _lib.load()
    it, wu = iters(args)
    t_setup = time.time()
    sds = make_experiments(rank * args.combos, args.combos)
    t_setup = time.time() - t_setup

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")
    pk = C.c_double()
    L.bsyn_measure_fp64_peak(dev, C.byref(pk))
    b = _lib.Batch(sds, device=dev)
    target = args.workload == "target"

    def one_step(seed, n_it, n_wu):
        """One pass of the hot path over the resident batch; returns (grad evals, device ms, extra info)."""
        if target:
            r = screen_to_ess(sds, target=ESS_TARGET, batch=b, device=dev, chains=args.chains, warmup=n_wu, block=n_it - n_wu,
                              seed=seed, max_rounds=args.max_rounds, retry=args.retry, want_summary=False)
            return r["n_grad"], r["kernel_ms"], r
        ng, ms = b.sample_device(chains=args.chains, iter=n_it, warmup=n_wu, seed=seed)
        return ng, ms, None

    # ---- device-resident leg -------------------------------------------------------------------
    for s in range(args.warmup):
        one_step(100 + s, max(10, it // 5), max(5, wu // 5))
        flush.fill_(s & 255)
    clocks = ClockSampler(dev)
    if rank == 0:          # one nvidia-smi poller per job: rank 0's GPU stands for the box (all ranks run the same load)
        clocks.start()
    barrier()
    launches1 = L.bsyn_launch_count()
    tot_ms, tot_grad, info = 0.0, 0, None
    t0 = time.time()
    for s in range(args.steps):
        ng, ms, info = one_step(1000 + s, it, wu)
        tot_ms += ms
        tot_grad += ng
        flush.fill_(s & 255)
    barrier()
    wall = time.time() - t0
    clocks.stop_flag = True
    launches = L.bsyn_launch_count() - launches1
    if target:
        ess, rounds = info["ess"], info["rounds"]
    else:
        ess = np.fmin(b.bulk_ess("VUS_syn")[0], b.bulk_ess("VUS_ant")[0])
        rounds = 1
    # ---- end-to-end leg: host buffers in, summaries out, through the public call --------------------
    b.close()
    e2e_t, e2e_grad = [], []
    for s in range(args.e2e_reps):
        barrier()
        t = time.time()
        bb = _lib.Batch(sds, device=dev)
        if target:
            res = screen_to_ess(sds, target=ESS_TARGET, batch=bb, device=dev, chains=args.chains, warmup=wu, block=it - wu,
                                seed=2000 + s, max_rounds=args.max_rounds, retry=args.retry, want_summary=True)
        else:
            res = bb.sample(chains=args.chains, iter=it, warmup=wu, seed=2000 + s, want_qdraws=False, want_sampler_params=False)
        bb.close()
        torch.cuda.synchronize()
        e2e_t.append(time.time() - t)
        e2e_grad.append(res["n_grad"])
    d2h = 8 * (args.combos * 12 * 2 + args.combos * args.chains * 8)
    # ---- reduce over ranks -------------------------------------------------------------------------
    stats = torch.tensor([tot_ms, float(tot_grad), wall, float(np.nansum(ess)), float(np.sum(e2e_grad)),
                          float(np.sum(e2e_t)), float(np.isnan(ess).sum()), float(launches),
                          float((ess >= ESS_TARGET).sum()), float(rounds)], dtype=torch.float64, device="cuda")
    mx = stats.clone()
    sm = stats.clone()
    if world > 1:
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        dist.all_reduce(sm, op=dist.ReduceOp.SUM)
    mx, sm = mx.cpu().numpy(), sm.cpu().numpy()
    if rank == 0:
        dev_s = mx[0] * 1e-3
        value = sm[1] / dev_s
        e2e_value = sm[4] / mx[5] if mx[5] > 0 else None
        n_combo_total = args.combos * world
        step_time = dev_s / args.steps
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        achieved_tflops = value / world * FLOP_PER_EVAL / 1e12
        peak_tflops = 2.0 * pk.value / 1e12
        traffic, traffic_src = measured_traffic(args)
        out = {
            "metric": METRIC, "value": value, "unit": "grad evals/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * step_time, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(args),
            "ess_per_sec_per_combo": float(sm[3] / n_combo_total / step_time),
            "total_ess_per_sec": float(sm[3] / step_time),
            "ess_min_syn_ant_mean": float(sm[3] / n_combo_total), "ess_nan_combos": int(sm[6]),
            # time until every combination of the step has ESS >= 400 -- the step itself when all do
            "ess_ge_400_frac": float(sm[8] / n_combo_total),
            "time_to_ess400_all_s": float(step_time) if int(sm[8]) == n_combo_total else None,
            "sampler_launch_rounds_last_step": int(mx[9]),
            "target_mode": ({"phase1_max_rounds": args.max_rounds, "phase1_s_last_step_rank0": info["phase1_ms"] * 1e-3,
                             "phase2_s_last_step_rank0": info["phase2_ms"] * 1e-3,
                             "phase2": args.retry, "combos_in_phase2_at_adapt_delta_0.99_last_step_rank0": int(info["n_retry"]),
                             "combos_below_target_after_both_phases_rank0": int(args.combos - info["reached"]),
                             "note": "phase 2 = refit from scratch (refit) or resumed chains with the step size re-adapted (resume), both at adapt_delta 0.99; what stays below the target after it has R-hat > 1.05 (divergent / "
                                     "multimodal posteriors: profiles/r02_target_diag.json); the reference reports such fits with "
                                     "returnCode 1"} if target else None),
            "grad_evals_per_step": sm[1] / args.steps,
            "wall_s_timed_region": float(mx[2]),
            "setup_s_rank0": {"synthetic data + prepare (host, before any timed region)": t_setup},
            "e2e": {"value": e2e_value, "unit": "grad evals/s", "h2d_bytes_per_step": h2d_bytes(sds),
                    "d2h_bytes_per_step": d2h, "seconds_per_step": float(mx[5] / len(e2e_t)) if e2e_t else None,
                    "reps": len(e2e_t)},
            "gpu_launches": int(mx[7]),
            "clocks": clocks.summary(),
            # the bound of this path is FP64 issue throughput (neither HBM nor tensor cores)
            "roofline": {"bound": "fp64", "achieved": achieved_tflops, "peak": peak_tflops, "unit": "TFLOP/s",
                         "frac": achieved_tflops / peak_tflops if peak_tflops > 0 else None,
                         "peak_source": "bsyn_measure_fp64_peak: register-resident DFMA microbenchmark, measured in this run "
                                        "(MEASURED_PEAKS.json has no FP64 figure)",
                         "flop_per_eval": FLOP_PER_EVAL, "per_gpu": True,
                         "hbm": {"peak_gbs": peaks.get("hbm_gbs"), "note": "state is on-chip; HBM is not the bound"},
                         "traffic": traffic, "traffic_source": traffic_src},
        }
        if world == 1 and not args.no_cpu:
            from oracle.c_oracle import lib
            cores = lib().orc_num_threads()
            nc = max(1, 2 * cores // args.chains)
            r = cpu_leg(nc, args.chains, it, wu, cores)
            out["cpu_baseline"] = {"value": r["value"], "unit": "grad evals/s", "cores": cores, "kind": "port",
                                   "sample": f"{nc} cfg5 combinations x {args.chains} chains x {it} iterations "
                                             f"({wu} warm-up), {r['seconds']:.1f} s",
                                   "note": "oracle/gp_oracle.c NUTS: CPU restatement (not rstan)",
                                   "ess_per_sec_per_combo": r["ess_per_sec_per_combo"],
                                   "total_ess_per_sec": r["total_ess_per_sec"]}
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="screen", choices=["screen", "target"])
    ap.add_argument("--combos", type=int, default=int(os.environ.get("BENCH_COMBOS", "1332")),
                    help="combinations per GPU per step (the full cfg5 screen is 12500 per GPU on 8 GPUs; 1332 = 3 chains "
                         "per resident warp slot (148 SMs x 12 warps) keeps the launch tail amortised and a step near 12 s, "
                         "so that --steps 20 --warmup 5 ends within ~5 minutes)")
    ap.add_argument("--chains", type=int, default=4)
    ap.add_argument("--iter", type=int, default=1150)
    ap.add_argument("--warmup-iter", type=int, default=1000,
                    help="warm-up iterations per chain (rstan: 1000); --iter 2000 --warmup-iter 1000 is rstan's default budget")
    ap.add_argument("--retry", choices=["refit", "resume"], default="refit",
                    help="--workload target: phase 2 for what is below the target after phase 1 (api.screen_to_ess)")
    ap.add_argument("--max-rounds", type=int, default=6, help="--workload target: cap on the launches of phase 1 (warm-up + "
                    "extensions); what is still below the target is refitted at adapt_delta 0.99 (api.screen_to_ess)")
    ap.add_argument("--e2e-reps", type=int, default=2)
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
