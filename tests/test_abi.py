"""CPU tests of the boundary: the C-ABI library builds, loads and exports every symbol include/bsyn.h
declares; the host-side mirror of bayesynergy()'s data preparation; loud failure without a GPU."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from bayesynergy_b200 import _lib
from bayesynergy_b200.dataprep import prepare, param_names, output_names, synthetic_experiment

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol(built):
    hdr = open(os.path.join(ROOT, "include", "bsyn.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(bsyn_[a-z0-9_]+)\s*\(", hdr))
    assert {"bsyn_batch_create", "bsyn_logp_grad", "bsyn_sample", "bsyn_write_array"} <= declared
    L = _lib.load()
    for name in sorted(declared):
        assert hasattr(L, name), f"{name} declared in include/bsyn.h but not exported by libbsyn.so"
    for name in _lib.EXPORTS:
        assert name in declared
    assert L.bsyn_version() == 201


def test_no_cpu_fallback(built, datasets):
    """Without a CUDA device the product path must fail loudly (never fall back to the oracle)."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    y, x = datasets["mathews:ispinesib + ibrutinib"]
    with pytest.raises(_lib.BsynError) as ei:
        _lib.Batch([prepare(y, x)])
    assert "no CUDA device" in str(ei.value) or "cuda" in str(ei.value).lower()


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "bayesynergy_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src.lower() or f == "diagnostics.py" and "oracle" not in src, (dirpath, f)


def test_sampler_defaults_match_rstan_and_bayesynergy(built):
    a = _lib.Batch.sampler_args()
    assert (a.chains, a.iter, a.warmup, a.thin, a.max_treedepth) == (4, 2000, 1000, 1, 10)
    assert a.adapt_delta == 0.9                      # R/bayesynergy.R:107-110
    assert (a.adapt_gamma, a.adapt_kappa, a.adapt_t0) == (0.05, 0.75, 10.0)
    assert (a.adapt_init_buffer, a.adapt_term_buffer, a.adapt_window) == (75, 50, 25)
    assert a.init_r == 2.0 and a.stepsize == 1.0 and a.stepsize_jitter == 0.0


# ---- data preparation (R/bayesynergy.R:63-200) ------------------------------------------------------

def test_prepare_mathews_matches_survey_appendix_f(datasets):
    y, x = datasets["mathews:ispinesib + ibrutinib"]
    sd = prepare(y, x)
    assert (sd.n1, sd.n2, sd.N, sd.nrep, sd.nmissing) == (5, 5, 36, 1, 0)
    # row k: x1 level a = k mod 6, x2 level b = k div 6  =>  ii_obs = a*6 + b + 1
    k = np.arange(36)
    assert np.array_equal(sd.ii_obs, (k % 6) * 6 + k // 6 + 1)
    assert np.allclose(sd.x1, np.log10([0.1954, 0.7812, 3.125, 12.5, 50]))
    assert (sd.kernel, sd.nu_matern, sd.est_alpha) == (2, 2, 0)      # package default: Matern 3/2
    assert sd.dim_gp_grid == 40 and sd.dim_nointeraction == 11
    assert len(param_names(sd)) == 40 and param_names(sd)[12] == "z[1,1]" and param_names(sd)[13] == "z[2,1]"
    assert len(output_names(sd)) == 40 + 3 * 25 + 10 + 2 + 36 + 7


def test_prepare_oneil_shapes(datasets):
    # n1=n2=11, N=256 / 208, nrep=12 / 9, nmissing=1472 / 1088, 32 distinct observed cells
    for key, (N, nrep, nmiss) in {"oneil_failed:0": (256, 12, 1472), "oneil_failed:1": (208, 9, 1088)}.items():
        y, x = datasets[key]
        sd = prepare(y, x)
        assert (sd.n1, sd.n2, sd.N, sd.nrep, sd.nmissing) == (11, 11, N, nrep, nmiss)
        assert len(set(sd.ii_obs.tolist())) == 32 and 1 not in sd.ii_obs


def test_prepare_replicate_matrix_and_missing():
    y, x = synthetic_experiment(3)
    y = y.copy()
    y[5, 1] = np.nan
    y[17, 2] = np.nan
    sd = prepare(y, x)
    assert (sd.n1, sd.n2, sd.nrep, sd.nmissing, sd.N) == (10, 10, 3, 2, 361)
    assert sd.ii_obs.min() == 1 and sd.ii_obs.max() == 121


def test_prepare_kernel_types_and_errors(datasets):
    y, x = datasets["mathews:ispinesib + ibrutinib"]
    assert (prepare(y, x, type=2).kernel, prepare(y, x, type=2).est_alpha) == (1, 0)
    assert (prepare(y, x, type=4).kernel, prepare(y, x, type=4).est_alpha) == (3, 1)
    assert prepare(y, x, type=3, nu=2.5).nu_matern == 3
    with pytest.raises(ValueError, match="type"):
        prepare(y, x, type=5)
    with pytest.raises(ValueError, match="PC prior"):
        prepare(y, x, type=2, pcprior=True)
    with pytest.raises(ValueError, match="rho"):
        prepare(y, x, robust=True, rho=0.5)
    with pytest.raises(ValueError, match="monotherapy"):
        prepare(y[7:], x[7:] + 1.0)
    with pytest.raises(ValueError, match="equally sized"):
        prepare(y[:-1], x)
    with pytest.raises(ValueError, match="Method"):
        prepare(y, x, method="opt")


def test_split_ess_reference():
    from bayesynergy_b200.diagnostics import split_ess_rhat
    rng = np.random.default_rng(0)
    x = rng.normal(size=(4, 1000))
    ess, rhat = split_ess_rhat(x)
    assert 3000 < ess < 5200 and abs(rhat - 1) < 0.01
    # AR(1) with phi = 0.9: ESS ~ N (1-phi)/(1+phi)
    z = np.zeros((4, 4000))
    e = rng.normal(size=z.shape)
    for t in range(1, z.shape[1]):
        z[:, t] = 0.9 * z[:, t - 1] + e[:, t]
    ess, rhat = split_ess_rhat(z)
    assert 500 < ess < 1300


def test_native_data_prep_matches_numpy(built, datasets):
    """bsyn_prepare (C, no GPU needed) against the numpy statement of R/bayesynergy.R:129-154 on every fixture:
    replicate matrices with missing wells, replicates as repeated rows, shuffled rows, non-aligned O'Neil grids."""
    import numpy as np
    from bayesynergy_b200.dataprep import prepare, synthetic_experiment, synthetic_oneil_experiment
    cases = dict(datasets)
    cases["cfg5"] = synthetic_experiment(3)
    cases["oneil_synth"] = synthetic_oneil_experiment(5)
    rng = np.random.default_rng(0)
    y, x = synthetic_experiment(4)
    y = y.copy()
    y[rng.integers(0, y.shape[0], 9), rng.integers(0, y.shape[1], 9)] = np.nan
    perm = rng.permutation(y.shape[0])
    cases["cfg5_missing_shuffled"] = (y[perm], x[perm])
    for name, (y, x) in cases.items():
        a, b = prepare(y, x), prepare(y, x, native=True)
        assert (a.n1, a.n2, a.nrep, a.nmissing, a.N) == (b.n1, b.n2, b.nrep, b.nmissing, b.N), name
        assert np.array_equal(a.ii_obs, b.ii_obs) and np.array_equal(a.y, b.y), name
        assert np.allclose(a.x1, b.x1, rtol=0, atol=1e-15) and np.allclose(a.x2, b.x2, rtol=0, atol=1e-15), name
    with pytest.raises(ValueError, match="Missing monotherapy data"):
        prepare(np.ones(4), np.array([[1.0, 0.0], [2.0, 0.0], [1.0, 1.0], [2.0, 1.0]]), native=True)
