"""A plain-C client of the drop-in boundary: tests/c/abi_smoke.c is compiled with gcc against include/bsyn.h, linked
with libbsyn.so and calls every export (no ctypes, no Python between the caller and the library)."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIBDIR = os.path.join(ROOT, "bayesynergy_b200", "lib")
EXE = os.path.join(ROOT, "tests", "c", "abi_smoke")


def _build():
    cmd = ["gcc", "-std=c99", "-Wall", "-Werror", "-O1", "-I", os.path.join(ROOT, "include"),
           os.path.join(ROOT, "tests", "c", "abi_smoke.c"), "-o", EXE, "-L", LIBDIR, "-lbsyn", "-lm",
           "-Wl,-rpath," + LIBDIR]
    out = subprocess.run(cmd, capture_output=True, text=True)
    assert out.returncode == 0, out.stderr


def test_c_client_without_gpu_fails_loudly(built):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present: see the gpu-marked variant")
    _build()
    out = subprocess.run([EXE], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "abi_smoke ok" in out.stdout and "no gpu" in out.stdout


def test_c_client_calls_every_export(built):
    """Every function include/bsyn.h declares is called by the C client (so a signature drift breaks its build)."""
    import re
    hdr = re.sub(r"/\*.*?\*/", "", open(os.path.join(ROOT, "include", "bsyn.h")).read(), flags=re.S)
    declared = set(re.findall(r"\b(bsyn_[a-z0-9_]+)\s*\(", hdr))
    src = open(os.path.join(ROOT, "tests", "c", "abi_smoke.c")).read()
    missing = sorted(n for n in declared if n + "(" not in src)
    assert not missing, missing


@pytest.mark.gpu
def test_c_client_on_gpu(built):
    _build()
    out = subprocess.run([EXE, "--gpu"], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "abi_smoke ok" in out.stdout


def _build_stanfit(werror=False):
    exe = os.path.join(ROOT, "tests", "c", "stanfit_smoke")
    cmd = ["g++", "-std=c++14", "-Wall", "-O1", "-I", os.path.join(ROOT, "include"),
           os.path.join(ROOT, "tests", "c", "stanfit_smoke.cpp"), "-o", exe, "-L", LIBDIR, "-lbsyn",
           "-Wl,-rpath," + LIBDIR] + (["-Werror"] if werror else [])
    out = subprocess.run(cmd, capture_output=True, text=True)
    assert out.returncode == 0, out.stderr
    return exe


def test_stanfit_header_compiles(built):
    """No GPU: include/bsyn_stanfit.hpp and its client compile and link against libbsyn.so."""
    _build_stanfit(werror=True)


@pytest.mark.gpu
def test_stanfit_contract_in_cpp(built):
    """include/bsyn_stanfit.hpp: call_sampler's return contract and the param_* bookkeeping of rstan::stan_fit,
    compiled as plain C++ against the C ABI (what the Rcpp module of INTEGRATION.md wraps)."""
    exe = _build_stanfit()
    out = subprocess.run([exe], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0 and "stanfit_smoke ok" in out.stdout, out.stdout + out.stderr
