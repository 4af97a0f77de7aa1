"""GPU parity tests (run on the B200 box): the CUDA path, called through the C ABI, against the oracle
and the committed golden vectors.  Tolerance: relative 1e-10 in FP64 (north_star) wherever the kernel
matrices have condition number <= 1e7, scaled by cond/1e7 beyond (see tests/test_oracle.py::tol_for)."""
import itertools

import numpy as np
import pytest

from bayesynergy_b200 import _lib
from bayesynergy_b200.dataprep import prepare, output_names
from oracle.c_oracle import Problem

pytestmark = pytest.mark.gpu
REL = 1e-10


def tol_for(cond):
    return max(REL, 1e-17 * cond)


def relerr(a, b):
    a, b = np.asarray(a, dtype=float), np.asarray(b, dtype=float)
    return np.abs(a - b).max() / max(1.0, np.abs(b).max())


def test_golden_vectors_through_c_abi(built, datasets, golden_cases):
    groups = {}
    for c in golden_cases:
        key = (c["dataset"], tuple(sorted(c["variant"].items())), c["model"])
        groups.setdefault(key, []).append(c)
    worst_wc = 0.0
    n = 0
    for (dname, var, model), cs in groups.items():
        y, x = datasets[dname]
        sd = prepare(y, x, **dict(var))
        b = _lib.Batch([sd], model=model)
        U = np.array([c["u"] for c in cs])
        lp, g, st = b.logp_grad(U)
        D = b.num_params(0)
        for k, c in enumerate(cs):
            assert st[k] == 0
            e = max(abs(lp[k] - c["lp"]) / max(1.0, abs(c["lp"])), relerr(g[k, :D], c["grad"]))
            assert e <= tol_for(c["cond"]), (dname, dict(var), model, e, c["cond"])
            if c["cond"] <= 1e6:
                worst_wc = max(worst_wc, e)
            n += 1
        b.close()
    assert n == len(golden_cases)
    assert worst_wc <= 1e-11, worst_wc


@pytest.mark.parametrize("dname", ["mathews:ispinesib + ibrutinib", "mathews:canertinib + ibrutinib",
                                   "oneil_failed:0", "oneil_failed:1", "synthetic:7"])
def test_random_points_against_oracle(built, datasets, dname, cond_fn):
    y, x = datasets[dname]
    rng = np.random.default_rng(abs(hash(dname)) % (2 ** 31))
    variants = [dict(), dict(type=2), dict(type=4), dict(type=3, nu=0.5), dict(type=3, nu=2.5),
                dict(robust=True), dict(heteroscedastic=False), dict(robust=True, heteroscedastic=False),
                dict(pcprior=True, robust=True), dict(lower_asymptotes=False)]
    for v in variants:
        sd = prepare(y, x, **v)
        for model in ("gp_grid", "nointeraction"):
            P = Problem(sd, model)
            b = _lib.Batch([sd], model=model)
            assert b.num_params(0) == P.dim and b.num_outputs(0) == P.n_out
            U = rng.uniform(-0.7, 0.7, size=(24, P.dim))      # moderate ell: well-conditioned kernel matrices
            U[12:, P.dim - 3] = rng.uniform(-3.0, -1.5, size=12)    # small s: large residuals / LPTN tails
            for jac in (True, False):
                lp, g, st = b.logp_grad(U, jacobian=jac)
                st_o, lp_o, g_o = P.logp_grad_batch(U, jacobian=jac)
                assert np.array_equal(st, st_o) and (st == 0).all()
                for k in range(U.shape[0]):
                    tol = tol_for(cond_fn(sd, U[k], model))
                    assert abs(lp[k] - lp_o[k]) <= tol * max(1.0, abs(lp_o[k])), (dname, v, model, k, lp[k], lp_o[k])
                    assert relerr(g[k, :P.dim], g_o[k]) <= tol, (dname, v, model, k)
            b.close()


def test_intermediates_cholesky_and_kron(built, datasets):
    """Kernel-internal checks: L1, L2 (two Choleskys) and GP = L2 z L1^T against numpy."""
    import torch
    from oracle.torch_oracle import _kernel_mats
    for dname in ("mathews:ispinesib + ibrutinib", "synthetic:7", "oneil_failed:0"):
        y, x = datasets[dname]
        sd = prepare(y, x)
        b = _lib.Batch([sd])
        D = b.num_params(0)
        U = np.random.default_rng(4).uniform(-0.5, 0.5, size=(3, D))
        lp, g, st, dbg = b.logp_grad(U, debug=True)
        NM, NN, GM = dbg["NM"], dbg["NN"], dbg["GM"]
        for k in range(3):
            raw = dbg["raw"][k]
            k0 = 2 * sd.est_la + 6
            ell, sigf = np.exp(U[k, k0 + 2]), np.exp(U[k, k0 + 3])
            c1, c2 = _kernel_mats(sd, torch.tensor(ell), torch.tensor(sigf), None)
            L1 = np.linalg.cholesky(c1.numpy())
            L2 = np.linalg.cholesky(c2.numpy())
            L1g = raw[:NN].reshape(NM, NM)[:sd.n1, :sd.n1]
            L2g = raw[NN:2 * NN].reshape(NM, NM)[:sd.n2, :sd.n2]
            assert relerr(L1g, L1) < 1e-10 and relerr(L2g, L2) < 1e-10
            z = U[k, k0 + 4:k0 + 4 + sd.n1 * sd.n2].reshape(sd.n1, sd.n2).T
            GP = L2 @ z @ L1.T
            GPg = raw[2 * NN:2 * NN + sd.n1 * sd.n2].reshape(sd.n1, sd.n2).T
            assert relerr(GPg, GP) < 1e-10
        b.close()


def test_heterogeneous_batch_and_status(built, datasets):
    """Several experiments of different size in one handle, points addressed by prob_idx; domain errors."""
    sds = [prepare(*datasets[k]) for k in ("mathews:ispinesib + ibrutinib", "oneil_failed:1", "synthetic:7",
                                          "mathews:canertinib + ibrutinib")]
    b = _lib.Batch(sds)
    Ps = [Problem(sd) for sd in sds]
    rng = np.random.default_rng(8)
    idx = rng.integers(0, len(sds), size=40).astype(np.int32)
    U = np.zeros((40, b.max_params))
    for k in range(40):
        U[k, :Ps[idx[k]].dim] = rng.uniform(-0.8, 0.8, Ps[idx[k]].dim)
    bad = [3, 17]
    for k in bad:
        U[k, 2 * sds[idx[k]].est_la + 6 + 3] = 800.0     # sigma_f = inf -> not positive definite -> reject
    lp, g, st = b.logp_grad(U, prob_idx=idx)
    for k in range(40):
        st_o, lp_o, g_o = Ps[idx[k]].logp_grad(U[k, :Ps[idx[k]].dim])
        assert st[k] == st_o
        if k in bad:
            assert st[k] == 1 and lp[k] == -np.inf and not g[k].any()
        else:
            assert abs(lp[k] - lp_o) <= 5e-10 * max(1.0, abs(lp_o))
            assert relerr(g[k, :Ps[idx[k]].dim], g_o) <= 5e-10
            assert not g[k, Ps[idx[k]].dim:].any()
    # empty call and argument errors
    lp0, g0, st0 = b.logp_grad(np.zeros((0, b.max_params)))
    assert lp0.shape == (0,)
    with pytest.raises(_lib.BsynError):
        b.logp_grad(U, prob_idx=np.full(40, 99, dtype=np.int32))
    b.close()


def test_nonfinite_gradient_where_reference_autodiff_breaks(built, datasets):
    """Where a monotherapy curve saturates so that the Bliss product rounds to exactly 1 (or 10^(...) overflows),
    the reference's log(p0 / (1 - p0)) is infinite: lp stays finite but the reverse sweep is 0 * inf = NaN (Stan then
    rejects the initial point / counts the leapfrog as divergent / aborts ADVI).  Same verdict on the GPU."""
    sd = prepare(*datasets["mathews:ispinesib + ibrutinib"])
    b = _lib.Batch([sd])
    P = Problem(sd)
    rng = np.random.default_rng(3)
    U = rng.uniform(-0.5, 0.5, size=(6, P.dim))
    off = 2 * sd.est_la
    U[0, off + 4] = U[0, off + 5] = 3.5          # both slopes ~ 33: p01 = p02 = 1 at the lowest doses -> p0 == 1
    U[0, off + 0], U[0, off + 1] = 1.0, 3.5
    U[1, off + 4] = 3.5                          # only drug 1 saturates: p0 = p02 < 1 everywhere -> finite
    U[1, off + 0] = 1.0
    U[2, off + 4] = 6.0; U[2, off + 0] = -2.0    # 10^(slope (x - m)) overflows at the top dose
    lp, g, st = b.logp_grad(U)
    for k in range(6):
        st_o, lp_o, g_o = P.logp_grad(U[k])
        assert st[k] == st_o == 0
        assert abs(lp[k] - lp_o) <= 5e-10 * max(1.0, abs(lp_o))
        assert np.isfinite(g_o).all() == np.isfinite(g[k, :P.dim]).all(), k
        if np.isfinite(g_o).all():
            assert relerr(g[k, :P.dim], g_o) <= 5e-10
    assert not np.isfinite(g[0, :P.dim]).any() and np.isfinite(g[1, :P.dim]).all() and not np.isfinite(g[2, :P.dim]).any()
    b.close()


def test_value_only_path_matches_full_evaluation(built, datasets):
    """log_prob(upar, jacobian, gradient = FALSE) skips the reverse sweep (bridge sampling, ADVI's ELBO samples):
    same lp and status as the full evaluation, for every likelihood / kernel family and both models."""
    rng = np.random.default_rng(12)
    for dname, kw, model in (("mathews:ispinesib + ibrutinib", {}, "gp_grid"),
                             ("synthetic:7", dict(robust=True, pcprior=True), "gp_grid"),
                             ("oneil_failed:1", dict(type=4, heteroscedastic=False), "gp_grid"),
                             ("mathews:canertinib + ibrutinib", dict(type=2), "nointeraction")):
        sd = prepare(*datasets[dname], **kw)
        b = _lib.Batch([sd], model=model)
        U = rng.uniform(-1.0, 1.0, size=(24, b.num_params(0)))
        U[5, 2 * sd.est_la + 6 + (3 if model == "gp_grid" else 0)] = 800.0      # a rejected point
        for jac in (True, False):
            lp_f, g, st_f = b.logp_grad(U, jacobian=jac)
            lp_v, none, st_v = b.logp_grad(U, jacobian=jac, want_grad=False)
            assert none is None and np.array_equal(st_f, st_v)
            okm = st_f == 0
            assert np.allclose(lp_v[okm], lp_f[okm], rtol=1e-13, atol=1e-12), (dname, jac)
            assert np.array_equal(np.isinf(lp_v), np.isinf(lp_f))
        b.close()


def test_output_names_from_the_library(built, datasets):
    """bsyn_output_name (what the Rcpp shim needs for param_names / param_fnames_oi, cc:16-18) against the Python
    statement of the declaration order (src/stanExports_gp_grid.h:2256-2402), for every layout variant."""
    for dname, kw, model in (("mathews:ispinesib + ibrutinib", {}, "gp_grid"),
                             ("mathews:ispinesib + ibrutinib", dict(lower_asymptotes=False), "gp_grid"),
                             ("oneil_failed:1", dict(type=4), "gp_grid"),
                             ("synthetic:7", {}, "nointeraction"),
                             ("synthetic:7", dict(lower_asymptotes=False), "nointeraction")):
        sd = prepare(*datasets[dname], **kw)
        b = _lib.Batch([sd], model=model)
        assert b.output_names(0) == output_names(sd, model), (dname, kw, model)
        b.close()


def test_write_array_against_oracle(built, datasets, cond_fn):
    for dname, v in itertools.product(("mathews:ispinesib + ibrutinib", "synthetic:7", "oneil_failed:0"),
                                      (dict(), dict(type=2, heteroscedastic=False), dict(robust=True, lower_asymptotes=False))):
        y, x = datasets[dname]
        sd = prepare(y, x, **v)
        for model in ("gp_grid", "nointeraction"):
            P = Problem(sd, model)
            b = _lib.Batch([sd], model=model)
            names = output_names(sd, model)
            U = np.random.default_rng(2).uniform(-0.7, 0.7, size=(6, P.dim))
            out, st = b.write_array(U)
            for k in range(6):
                st_o, out_o = P.write_array(U[k])
                assert st[k] == st_o == 0
                err = np.abs(out[k, :P.n_out] - out_o) / np.maximum(1.0, np.abs(out_o))
                for nm in ("rVUS_f", "rVUS_p0", "VUS_syn", "VUS_ant"):      # exact zeros are replaced by random draws
                    if nm in names and 1e-6 <= out_o[names.index(nm)] <= 1e-4:
                        assert 1e-6 <= out[k, names.index(nm)] <= 1e-4
                        err[names.index(nm)] = 0.0
                worst = int(np.argmax(err))
                assert err.max() <= 10 * tol_for(cond_fn(sd, U[k], model)), (dname, v, model, names[worst], out[k, worst], out_o[worst])
            b.close()


def test_write_array_zero_patching(built, datasets):
    """Exact-zero VUS_syn / VUS_ant are replaced by U(1e-6, 1e-4) draws (gp_grid.stan:294-297)."""
    y, x = datasets["mathews:ispinesib + ibrutinib"]
    sd = prepare(y, x)
    b = _lib.Batch([sd])
    names = output_names(sd)
    u = np.zeros(b.num_params(0))
    u[12:12 + 25] = 1.0                 # z > 0 and L1, L2 > 0  =>  GP > 0  =>  Delta > 0 everywhere  =>  VUS_syn == 0
    out, st = b.write_array(u[None, :], seed=5)
    assert 1e-6 <= out[0, names.index("VUS_syn")] <= 1e-4
    assert out[0, names.index("VUS_ant")] > 1e-3 and out[0, names.index("VUS_Delta")] > 1e-3
    st_o, out_o = Problem(sd).write_array(u)
    assert 1e-6 <= out_o[names.index("VUS_syn")] <= 1e-4          # the oracle patches the same entry
    u[12:12 + 25] = -1.0
    out, st = b.write_array(u[None, :], seed=5)
    assert 1e-6 <= out[0, names.index("VUS_ant")] <= 1e-4 and out[0, names.index("VUS_syn")] < -1e-3
    b.close()


def test_unconstrain_roundtrip(built, datasets):
    y, x = datasets["synthetic:7"]
    for model, v in (("gp_grid", dict(type=4)), ("nointeraction", dict()), ("gp_grid", dict(lower_asymptotes=False))):
        sd = prepare(y, x, **v)
        b = _lib.Batch([sd], model=model)
        D = b.num_params(0)
        u = np.random.default_rng(3).uniform(-1, 1, D)
        out, st = b.write_array(u[None, :])
        u2 = b.unconstrain(out[0, :D])
        assert np.abs(u2 - u).max() < 1e-12
        b.close()


def test_specialised_and_fallback_kernels_agree(built, datasets, monkeypatch):
    """The compile-time specialised kernels (SpecCT: package-default options; fixed 10 x 10 grid or run-time grid) and the
    run-time-flags fallback (BSYN_NO_FAST=1) against the oracle and against each other, for every data set whose default
    fit qualifies for a specialised variant, both models; plus a short sampler run through both (same seeds: the
    variants differ only in summation order, so the first draws agree closely and the posterior means within MC error)."""
    for dname in ("synthetic:7", "mathews:ispinesib + ibrutinib", "oneil_failed:0"):
        y, x = datasets[dname]
        sd = prepare(y, x)
        rng = np.random.default_rng(5)
        for model in ("gp_grid", "nointeraction"):
            P = Problem(sd, model)
            U = rng.uniform(-0.7, 0.7, size=(32, P.dim))
            st_o, lp_o, g_o = P.logp_grad_batch(U)
            res = {}
            for nofast in ("0", "1"):
                monkeypatch.setenv("BSYN_NO_FAST", nofast)
                b = _lib.Batch([sd], model=model)
                lp, g, st = b.logp_grad(U)
                lpv, _, _ = b.logp_grad(U, want_grad=False)           # value-only path
                res[nofast] = (lp, g[:, :P.dim])
                assert (st == 0).all()
                for k in range(U.shape[0]):
                    tol = tol_for(kernel_cond_safe(sd, U[k], model))
                    assert abs(lp[k] - lp_o[k]) <= tol * max(1.0, abs(lp_o[k])), (dname, model, nofast, k)
                    assert relerr(g[k, :P.dim], g_o[k]) <= tol, (dname, model, nofast, k)
                    assert abs(lpv[k] - lp[k]) <= 1e-12 * max(1.0, abs(lp[k]))
                b.close()
            assert relerr(res["0"][0], res["1"][0]) < 1e-11
    # sampler through both kernel families
    y, x = datasets["synthetic:7"]
    sd = prepare(y, x)
    means = {}
    for nofast in ("0", "1"):
        monkeypatch.setenv("BSYN_NO_FAST", nofast)
        b = _lib.Batch([sd, sd])
        r = b.sample(chains=4, iter=500, warmup=250, seed=12)
        b.close()
        assert (r["chain_stats"][:, :, 6] == 0).all()
        means[nofast] = r["summary"][:, _lib.Q_NAMES.index("rVUS_f"), :]
    assert np.abs(means["0"][:, 0] - means["1"][:, 0]).max() < 4 * means["0"][:, 1].max() / np.sqrt(200)


def kernel_cond_safe(sd, u, model):
    from conftest import kernel_cond
    return kernel_cond(sd, u, model)


def test_tiny_lambda_does_not_underflow_the_variance_product(built, datasets, cond_fn):
    """lambda is a user option (R/bayesynergy.R:57, default 0.005).  The kernels fold the heteroscedastic -n/2 log(f +
    lambda) terms of a lane into one logarithm of a product; with lambda = 1e-10 and f -> 0 at high doses that product
    would underflow, so factors below 1e-8 take their own logarithm.  lp and gradient against the oracle at such points."""
    y, x = datasets["synthetic:7"]
    sd = prepare(y, x, lambda_=1e-10)
    rng = np.random.default_rng(3)
    for model in ("gp_grid", "nointeraction"):
        P = Problem(sd, model)
        b = _lib.Batch([sd], model=model)
        U = rng.uniform(-0.5, 0.5, size=(32, P.dim))
        k0 = 2 * sd.est_la
        U[:, 0:2] = rng.uniform(-14.0, -9.0, size=(32, 2))            # lower asymptotes ~ 1e-5: f can reach ~ 1e-10
        U[:, k0 + 0:k0 + 2] = rng.uniform(-2.5, -1.5, size=(32, 2))   # log10 EC50 far below the top doses
        U[:, k0 + 4:k0 + 6] = rng.uniform(1.0, 1.6, size=(32, 2))     # steep slopes (log scale)
        lp, g, st = b.logp_grad(U)
        st_o, lp_o, g_o = P.logp_grad_batch(U)
        assert np.array_equal(st, st_o)
        checked = 0
        for k in range(U.shape[0]):
            if st[k] != 0 or not np.isfinite(g_o[k]).all():
                continue
            tol = max(1e-9, tol_for(cond_fn(sd, U[k], model)))       # 1/(f + lambda) ~ 1e10 amplifies rounding in f
            assert abs(lp[k] - lp_o[k]) <= tol * max(1.0, abs(lp_o[k])), (model, k, lp[k], lp_o[k])
            assert relerr(g[k, :P.dim], g_o[k]) <= 1e-6, (model, k)
            checked += 1
        assert checked >= 8
        b.close()


def test_gradient_is_the_derivative_of_lp_by_finite_differences(built, datasets):
    """A property that needs no oracle: the analytic gradient returned by the device kernels is the derivative of the
    lp they return.  Directional central differences (Richardson-extrapolated, O(h^4)) along random directions, for
    every kernel / likelihood / prior variant, both models, specialised and fallback kernels, on every data shape
    (5 x 5, 11 x 11 with unobserved cells, 10 x 10 x 3)."""
    rng = np.random.default_rng(11)
    variants = [dict(), dict(type=2), dict(type=4), dict(type=3, nu=2.5), dict(robust=True, pcprior=True),
                dict(heteroscedastic=False), dict(lower_asymptotes=False)]
    worst = 0.0
    for dname in ("mathews:ispinesib + ibrutinib", "oneil_failed:0", "synthetic:7"):
        y, x = datasets[dname]
        for v in variants:
            sd = prepare(y, x, **v)
            for model in ("gp_grid", "nointeraction"):
                b = _lib.Batch([sd], model=model)
                D = b.num_params(0)
                U = rng.uniform(-0.6, 0.6, size=(6, D))
                V = rng.standard_normal((6, D))
                V /= np.linalg.norm(V, axis=1, keepdims=True)
                lp0, g, st = b.logp_grad(U)
                assert (st == 0).all()
                h = 1e-3
                pts = np.concatenate([U + h * V, U - h * V, U + 0.5 * h * V, U - 0.5 * h * V])
                lp, _, st2 = b.logp_grad(pts, want_grad=False)
                assert (st2 == 0).all()
                d1 = (lp[0:6] - lp[6:12]) / (2 * h)
                d2 = (lp[12:18] - lp[18:24]) / h
                fd = (4 * d2 - d1) / 3                          # Richardson: error O(h^4)
                an = (g[:, :D] * V).sum(axis=1)
                err = np.abs(fd - an) / np.maximum(1.0, np.abs(an))
                worst = max(worst, err.max())
                assert err.max() < 2e-6, (dname, v, model, err, fd, an)
                b.close()
    assert worst < 2e-6
