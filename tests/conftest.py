import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def _datasets():
    from bayesynergy_b200.dataprep import synthetic_experiment
    m = json.load(open(os.path.join(GOLDEN, "mathews_DLBCL.json")))
    out = {}
    for k, e in m.items():
        out["mathews:" + k] = (np.array(e["y"]), np.array(e["x"]).reshape(2, -1).T)
    o = json.load(open(os.path.join(GOLDEN, "ONeil_A375_failed.json")))
    for i, e in enumerate(o):
        out[f"oneil_failed:{i}"] = (np.array(e["y"]), np.array(e["x"]).reshape(2, -1).T)
    for cid in (7,):
        out[f"synthetic:{cid}"] = synthetic_experiment(cid)
    return out


@pytest.fixture(scope="session")
def datasets():
    return _datasets()


@pytest.fixture(scope="session")
def golden_cases():
    return json.load(open(os.path.join(GOLDEN, "logp_golden.json")))["cases"]


@pytest.fixture(scope="session")
def built():
    """Build the native pieces once per session (nvcc cross-compiles without a GPU)."""
    import __graft_entry__ as ge
    ge.build()
    return True


def kernel_cond(sd, u, model="gp_grid"):
    """max condition number of the two kernel matrices at u (1 for nointeraction)."""
    if model != "gp_grid":
        return 1.0
    import torch
    from oracle.torch_oracle import _kernel_mats
    k0 = 2 * sd.est_la + 6
    ell, sigf = np.exp(u[k0 + 2]), np.exp(u[k0 + 3])
    alpha = torch.tensor(np.exp(u[k0 + 4])) if sd.est_alpha else None
    c1, c2 = _kernel_mats(sd, torch.tensor(ell), torch.tensor(sigf), alpha)
    return float(max(np.linalg.cond(c1.numpy()), np.linalg.cond(c2.numpy())))


@pytest.fixture(scope="session")
def cond_fn():
    return kernel_cond
