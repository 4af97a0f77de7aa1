"""CPU tests of the multi-GPU host logic with gloo, world_size 2: sharding by combination and the final gather
(the path's only collective).  The GPU work itself is covered by the `-m gpu` tests."""
import os
import subprocess
import sys
import textwrap

import warnings

import numpy as np
import pytest

from bayesynergy_b200.api import shard_slices, SCREEN_COLUMNS

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_slices_cover_and_balance():
    for n in (0, 1, 7, 8, 100, 100003):
        for w in (1, 2, 3, 8):
            sl = shard_slices(n, w)
            assert len(sl) == w and sl[0][0] == 0 and sl[-1][1] == n
            assert all(sl[r][1] == sl[r + 1][0] for r in range(w - 1))
            sizes = [hi - lo for lo, hi in sl]
            assert max(sizes) - min(sizes) <= 1


def test_screen_columns_match_reference():
    # R/synergyscreen.R:185-218 (23 columns, in order)
    assert len(SCREEN_COLUMNS) == 23
    assert SCREEN_COLUMNS[:4] == ["listID", "Experiment ID", "Drug A", "Drug B"]
    assert SCREEN_COLUMNS[-3:] == ["s", "returnCode", "divergentTransitions"]


def test_gather_screen_gloo_world2(tmp_path):
    script = tmp_path / "w.py"
    script.write_text(textwrap.dedent(f"""
        import os, sys, json
        sys.path.insert(0, {ROOT!r})
        import torch.distributed as dist
        from bayesynergy_b200.api import shard_slices, gather_screen
        dist.init_process_group("gloo")
        rank, world = dist.get_rank(), dist.get_world_size()
        n = 11
        lo, hi = shard_slices(n, world)[rank]
        local = dict(screenSummary=[dict(listID=k + 1, s=0.1 * k, returnCode=0) for k in range(lo, hi)])
        if rank == 1:
            local["failed"] = [dict(experiment_ID="bad")]
        out = gather_screen(local, world)
        if rank == 0:
            json.dump(out, open({str(tmp_path / 'out.json')!r}, "w"))
        dist.destroy_process_group()
    """))
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29511")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", "29511", str(script)],
                       env=env, capture_output=True, text=True, timeout=240)
    assert r.returncode == 0, r.stderr[-2000:]
    import json
    out = json.load(open(tmp_path / "out.json"))
    assert [row["listID"] for row in out["screenSummary"]] == list(range(1, 12))
    assert out["failed"] == [dict(experiment_ID="bad")]


def test_method_argument_is_validated():
    from bayesynergy_b200.api import bayesynergy, synergyscreen
    y = np.linspace(1, 0.2, 9)
    x = np.array([[a, b] for a in (0.0, 1.0, 2.0) for b in (0.0, 1.0, 2.0)])
    with pytest.raises(ValueError, match="Method must be one of"):
        bayesynergy(y, x, method="opt")
    with pytest.raises(ValueError, match="Method must be one of"):
        synergyscreen([dict(y=y, x=x)], bayesynergy_params=dict(method="opt"))


def _fake_res(n, s_vals, syn=5.0):
    from bayesynergy_b200._lib import Q_NAMES, BSYN_NQ
    S = np.ones((n, BSYN_NQ, 2))
    S[:, Q_NAMES.index("s"), 0] = s_vals
    S[:, Q_NAMES.index("VUS_syn"), 0] = syn
    return dict(summary=S, chain_stats=np.zeros((n, 4, 8)))


def _toy_experiments(n):
    y = np.linspace(1, 0.2, 9)
    x = np.array([[a, b] for a in (0.0, 1.0, 2.0) for b in (0.0, 1.0, 2.0)])
    return [dict(y=y, x=x, drug_names=[f"A{k}", f"B{k}"], experiment_ID=f"e{k}") for k in range(n)]


def test_synergyscreen_vb_retry_rule(monkeypatch, built):
    """method="vb": whatever errored (returnCode 2) or ended with s > 1 is refitted, up to max_retries more times
    (R/synergyscreen.R:159-167); what still fails is dropped from the summary and returned in `failed`."""
    from bayesynergy_b200 import api
    calls = []

    def fake_vb(sds, vbkw, seed, device):
        calls.append((len(sds), seed))
        n = len(sds)
        if len(calls) == 1:      # first pass over 5 experiments: #1 errors, #3 has s > 1
            rc = np.array([0, 2, 0, 0, 0])
            return _fake_res(n, [0.1, 0.0, 0.2, 1.7, 0.3]), rc
        if len(calls) == 2:      # retry of (#1, #3): #1 fine now, #3 still bad
            return _fake_res(n, [0.15, 2.5]), np.array([0, 0])
        return _fake_res(n, [3.0]), np.array([0])           # #3 never recovers

    monkeypatch.setattr(api, "_fit_batch_vb", fake_vb)
    monkeypatch.setattr(api, "prepare", lambda y, x, **kw: ("sd", len(y)))
    out = api.synergyscreen(_toy_experiments(5), max_retries=2, bayesynergy_params=dict(method="vb", seed=10))
    assert [c[0] for c in calls] == [5, 2, 1, 1]            # 1 + (max_retries + 1) passes, shrinking
    assert len({c[1] for c in calls}) == len(calls)         # a fresh seed per pass
    rows = out["screenSummary"]
    assert [r["Experiment ID"] for r in rows] == ["e0", "e1", "e2", "e4"]
    assert rows[1]["s"] == 0.15 and all(r["returnCode"] == 0 for r in rows)
    assert [e["experiment_ID"] for e in out["failed"]] == ["e3"]


def test_synergyscreen_sampling_retry_raises_adapt_delta(monkeypatch, built):
    """method="sampling": experiments with warnings are rerun with adapt_delta = 0.99 (R/synergyscreen.R:149-157)."""
    from bayesynergy_b200 import api
    seen = []

    def fake_fit(sds, kw, device, max_treedepth, **sinks):
        seen.append((len(sds), kw["adapt_delta"]))
        n = len(sds)
        rc = np.array([0, 1, 0]) if len(seen) == 1 else np.zeros(n, dtype=int)
        return _fake_res(n, [0.1] * n), rc

    monkeypatch.setattr(api, "_fit_batch", fake_fit)
    monkeypatch.setattr(api, "prepare", lambda y, x, **kw: ("sd", len(y)))
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        out = api.synergyscreen(_toy_experiments(3))
    assert seen == [(3, 0.9), (1, 0.99)]
    assert len(out["screenSummary"]) == 3 and "failed" not in out


class _FakeEssBatch:
    """Stands in for _lib.Batch in the host logic of api.screen_to_ess (no GPU): records the calls, returns scripted results."""
    log = []

    def __init__(self, sds, device=0, **kw):
        self.sds = list(sds)
        self.n = len(self.sds)

    def sample_to_ess(self, target=400.0, block=150, max_rounds=6, retry_rounds=6, want_summary=False, **kw):
        from bayesynergy_b200._lib import BSYN_NQ
        first = not any(c[0] == "fit" for c in _FakeEssBatch.log)
        _FakeEssBatch.log.append(("fit", self.n, retry_rounds, kw.get("adapt_delta"), kw.get("seed")))
        if first:      # the whole screen: experiments 1 and 3 stay below the target
            ess = np.array([500.0, 120.0, 800.0, np.nan, 450.0])[:self.n]
            rounds, ms, ng = 3, 100.0, 1000
        else:          # the refit of (1, 3): 1 recovers, 3 gets a worse-than-nothing 50
            ess = np.array([410.0, 50.0])[:self.n]
            rounds, ms, ng = 2, 40.0, 300
        res = dict(ess=ess, draws_per_chain=np.full(self.n, block * rounds, dtype=np.int32), n_grad=ng, kernel_ms=ms,
                   phase1_ms=ms if retry_rounds == 0 else 0.75 * ms, rounds=rounds,
                   n_retried=0 if retry_rounds == 0 else int((~(ess >= target)).sum()))
        if want_summary:
            res["summary"] = np.full((self.n, BSYN_NQ, 2), 1.0 if first else 2.0)
            res["chain_stats"] = np.full((self.n, 4, 8), 1.0 if first else 2.0)
        return res

    def close(self):
        _FakeEssBatch.log.append(("close", self.n))


def test_screen_to_ess_host_logic(monkeypatch):
    """api.screen_to_ess without a GPU: retry="refit" refits exactly what phase 1 left, at adapt_delta 0.99 with another seed,
    keeps the better of the two fits per experiment and adds the counters up; retry="resume" / None make one library call."""
    from bayesynergy_b200 import api, _lib
    monkeypatch.setattr(_lib, "Batch", _FakeEssBatch)
    sds = [f"sd{k}" for k in range(5)]
    _FakeEssBatch.log = []
    r = api.screen_to_ess(sds, target=400.0, seed=3, retry="refit")
    fits = [c for c in _FakeEssBatch.log if c[0] == "fit"]
    assert [(c[1], c[2], c[3]) for c in fits] == [(5, 0, 0.9), (2, 0, 0.99)] and fits[0][4] != fits[1][4]
    assert [c for c in _FakeEssBatch.log if c[0] == "close"] == [("close", 5), ("close", 2)]     # both batches closed
    assert r["n_retry"] == 2 and r["reached"] == 4 and r["rounds"] == 5 and r["n_grad"] == 1300
    assert r["kernel_ms"] == 140.0 and r["phase1_ms"] == 100.0 and r["phase2_ms"] == 40.0
    assert r["ess"][1] == 410.0 and r["ess"][3] == 50.0             # NaN counts as worse than any number
    assert r["summary"][1, 0, 0] == 2.0 and r["summary"][3, 0, 0] == 2.0 and r["summary"][0, 0, 0] == 1.0
    assert r["draws_per_chain"][1] == 600 and r["draws_per_chain"][0] == 450   # the refit runs blocks of retry_block = 300
    for mode, rr in (("resume", 6), (None, 0)):
        _FakeEssBatch.log = []
        r = api.screen_to_ess(sds, target=400.0, retry=mode)
        fits = [c for c in _FakeEssBatch.log if c[0] == "fit"]
        assert [(c[1], c[2]) for c in fits] == [(5, rr)]
        assert r["reached"] == 3 and r["n_retry"] == (2 if mode else 0)
        assert abs(r["phase1_ms"] + r["phase2_ms"] - r["kernel_ms"]) < 1e-9
    with pytest.raises(ValueError):
        api.screen_to_ess(sds, retry="again")
    # a batch handed in by the caller is left open
    _FakeEssBatch.log = []
    b = _FakeEssBatch(sds)
    api.screen_to_ess(sds, batch=b, retry=None)
    assert ("close", 5) not in _FakeEssBatch.log


def test_synergyscreen_sharded_world2_gloo(tmp_path):
    """synergyscreen(rank, world_size) + gather_screen under torchrun, world_size 2, gloo: each rank fits only its shard
    (the device work is faked -- this is the host logic of SURVEY 8e), listIDs stay global, rank 0 holds all rows in order."""
    script = tmp_path / "w.py"
    script.write_text(textwrap.dedent(f"""
        import os, sys, json
        sys.path.insert(0, {ROOT!r})
        import numpy as np
        import torch.distributed as dist
        from bayesynergy_b200 import api
        from bayesynergy_b200._lib import Q_NAMES, BSYN_NQ
        dist.init_process_group("gloo")
        rank, world = dist.get_rank(), dist.get_world_size()
        seen = []
        def fake_fit(sds, kw, device, max_treedepth, **sinks):
            seen.append(len(sds))
            S = np.ones((len(sds), BSYN_NQ, 2))
            S[:, Q_NAMES.index("s"), 0] = 0.1
            return dict(summary=S, chain_stats=np.zeros((len(sds), 4, 8))), np.zeros(len(sds), dtype=int)
        api._fit_batch = fake_fit
        api.prepare = lambda y, x, **kw: ("sd", len(y))
        y = np.linspace(1, 0.2, 9)
        x = np.array([[a, b] for a in (0.0, 1.0, 2.0) for b in (0.0, 1.0, 2.0)])
        exps = [dict(y=y, x=x, drug_names=[f"A{{k}}", f"B{{k}}"], experiment_ID=f"e{{k}}") for k in range(7)]
        local = api.synergyscreen(exps, rank=rank, world_size=world, device=0)
        out = api.gather_screen(local, world)
        json.dump(dict(seen=seen, n_local=len(local["screenSummary"])), open({str(tmp_path)!r} + f"/r{{rank}}.json", "w"))
        if rank == 0:
            json.dump([[r["listID"], r["Experiment ID"]] for r in out["screenSummary"]], open({str(tmp_path / 'out.json')!r}, "w"))
        dist.destroy_process_group()
    """))
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29517")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", "29517", str(script)],
                       env=env, capture_output=True, text=True, timeout=240)
    assert r.returncode == 0, r.stderr[-3000:]
    import json
    r0, r1 = json.load(open(tmp_path / "r0.json")), json.load(open(tmp_path / "r1.json"))
    assert r0["seen"] == [4] and r1["seen"] == [3] and r0["n_local"] == 4 and r1["n_local"] == 3
    rows = json.load(open(tmp_path / "out.json"))
    assert rows == [[k + 1, f"e{k}"] for k in range(7)]
