"""CPU tests of the multi-GPU host logic with gloo, world_size 2: sharding by combination and the final gather
(the path's only collective).  The GPU work itself is covered by the `-m gpu` tests."""
import os
import subprocess
import sys
import textwrap

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
