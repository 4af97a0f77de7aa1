"""ORACLE (test infrastructure, not product code) -- PARITY UNPINNED.

ctypes binding of oracle/libgp_oracle.so (the plain-C restatement, see gp_oracle.h).
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
may import this.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


class OrcProblem(C.Structure):
    _fields_ = [("model", C.c_int), ("n1", C.c_int), ("n2", C.c_int), ("N", C.c_int),
                ("est_la", C.c_int), ("robust", C.c_int), ("pcprior", C.c_int), ("heteroscedastic", C.c_int),
                ("kernel", C.c_int), ("nu_matern", C.c_int), ("est_alpha", C.c_int),
                ("lambda_", C.c_double), ("rho", C.c_double), ("pcprior_hypers", C.c_double * 4),
                ("y", C.POINTER(C.c_double)), ("ii_obs", C.POINTER(C.c_int)),
                ("x1", C.POINTER(C.c_double)), ("x2", C.POINTER(C.c_double))]


class OrcNutsArgs(C.Structure):
    _fields_ = [("num_warmup", C.c_int), ("num_samples", C.c_int), ("max_treedepth", C.c_int),
                ("adapt_delta", C.c_double), ("stepsize", C.c_double), ("stepsize_jitter", C.c_double),
                ("gamma", C.c_double), ("kappa", C.c_double), ("t0", C.c_double),
                ("init_buffer", C.c_int), ("term_buffer", C.c_int), ("window", C.c_int),
                ("init_radius", C.c_double), ("seed", C.c_uint64), ("chain_id", C.c_int),
                ("save_warmup", C.c_int), ("metric", C.c_int)]


class OrcAdviArgs(C.Structure):
    _fields_ = [("iter", C.c_int), ("grad_samples", C.c_int), ("elbo_samples", C.c_int), ("eval_elbo", C.c_int),
                ("eta", C.c_double), ("adapt_engaged", C.c_int), ("adapt_iter", C.c_int), ("tol_rel_obj", C.c_double),
                ("output_samples", C.c_int), ("init_radius", C.c_double), ("seed", C.c_uint64), ("chain_id", C.c_int)]


def build(force=False):
    so = os.path.join(_HERE, "libgp_oracle.so")
    srcs = [os.path.join(_HERE, f) for f in ("gp_oracle.c", "gp_oracle_batch.c", "gp_oracle_advi.c", "gp_oracle.h")]
    if force or not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return so


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(build())
        dp, ip = C.POINTER(C.c_double), C.POINTER(C.c_int)
        L.orc_dim.argtypes = [C.POINTER(OrcProblem)]
        L.orc_num_outputs.argtypes = [C.POINTER(OrcProblem)]
        L.orc_logp_grad.argtypes = [C.POINTER(OrcProblem), dp, C.c_int, dp, dp]
        L.orc_write_array.argtypes = [C.POINTER(OrcProblem), dp, dp, C.POINTER(C.c_uint64)]
        L.orc_nuts_defaults.argtypes = [C.POINTER(OrcNutsArgs)]
        L.orc_nuts_chain.argtypes = [C.POINTER(OrcProblem), C.POINTER(OrcNutsArgs), dp, dp, dp, dp, dp]
        L.orc_nuts_chain.restype = C.c_long
        L.orc_logp_grad_batch.argtypes = [C.POINTER(OrcProblem), C.c_int, dp, C.c_int, dp, dp, ip, C.c_int]
        L.orc_nuts_screen.argtypes = [C.POINTER(OrcProblem), C.c_int, C.c_int, C.POINTER(OrcNutsArgs), C.c_int,
                                      dp, dp, C.c_int]
        L.orc_nuts_screen.restype = C.c_long
        L.orc_num_threads.restype = C.c_int
        L.orc_advi_defaults.argtypes = [C.POINTER(OrcAdviArgs)]
        L.orc_advi_screen.argtypes = [C.POINTER(OrcProblem), C.c_int, C.POINTER(OrcAdviArgs), C.c_int, C.c_int,
                                      dp, dp, dp, C.c_int]
        L.orc_advi_screen.restype = None
        _LIB = L
    return _LIB


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double)) if a is not None else None


class Problem:
    """Keeps the numpy buffers alive for the C struct."""

    def __init__(self, sd, model="gp_grid"):
        self.sd = sd
        self.y = np.ascontiguousarray(sd.y, dtype=np.float64)
        self.ii = np.ascontiguousarray(sd.ii_obs, dtype=np.int32)
        self.x1 = np.ascontiguousarray(sd.x1, dtype=np.float64)
        self.x2 = np.ascontiguousarray(sd.x2, dtype=np.float64)
        p = OrcProblem()
        p.model = 0 if model == "gp_grid" else 1
        p.n1, p.n2, p.N = sd.n1, sd.n2, self.y.shape[0]
        p.est_la, p.robust, p.pcprior, p.heteroscedastic = sd.est_la, sd.robust, sd.pcprior, sd.heteroscedastic
        p.kernel, p.nu_matern, p.est_alpha = sd.kernel, sd.nu_matern, sd.est_alpha
        if p.model == 1:
            p.est_alpha = 0
        p.lambda_, p.rho = sd.lambda_, sd.rho
        for k in range(4):
            p.pcprior_hypers[k] = float(sd.pcprior_hypers[k])
        p.y = _dp(self.y)
        p.ii_obs = self.ii.ctypes.data_as(C.POINTER(C.c_int))
        p.x1, p.x2 = _dp(self.x1), _dp(self.x2)
        self.c = p
        self.dim = lib().orc_dim(C.byref(p))
        self.n_out = lib().orc_num_outputs(C.byref(p))

    def logp_grad(self, u, jacobian=True, want_grad=True):
        u = np.ascontiguousarray(u, dtype=np.float64)
        assert u.shape == (self.dim,)
        lp = C.c_double()
        g = np.empty(self.dim) if want_grad else None
        st = lib().orc_logp_grad(C.byref(self.c), _dp(u), int(jacobian), C.byref(lp), _dp(g))
        return st, lp.value, g

    def logp_grad_batch(self, U, jacobian=True, nthreads=0):
        U = np.ascontiguousarray(U, dtype=np.float64)
        B = U.shape[0]
        lp = np.empty(B)
        g = np.empty((B, self.dim))
        st = np.empty(B, dtype=np.int32)
        lib().orc_logp_grad_batch(C.byref(self.c), B, _dp(U), int(jacobian), _dp(lp), _dp(g),
                                  st.ctypes.data_as(C.POINTER(C.c_int)), nthreads)
        return st, lp, g

    def write_array(self, u, rng_state=1):
        u = np.ascontiguousarray(u, dtype=np.float64)
        out = np.empty(self.n_out)
        r = C.c_uint64(rng_state)
        st = lib().orc_write_array(C.byref(self.c), _dp(u), _dp(out), C.byref(r))
        return st, out

    def nuts_chain(self, num_warmup=1000, num_samples=1000, seed=1, chain_id=1, adapt_delta=0.9,
                   max_treedepth=10, want_outputs=True):
        a = OrcNutsArgs()
        lib().orc_nuts_defaults(C.byref(a))
        a.num_warmup, a.num_samples, a.seed, a.chain_id = num_warmup, num_samples, seed, chain_id
        a.adapt_delta, a.max_treedepth = adapt_delta, max_treedepth
        draws = np.empty((num_samples, self.dim))
        outs = np.empty((num_samples, self.n_out)) if want_outputs else None
        sp = np.empty((num_samples, 7))
        eps = C.c_double()
        minv = np.empty(self.dim)
        ng = lib().orc_nuts_chain(C.byref(self.c), C.byref(a), _dp(draws), _dp(outs), _dp(sp), C.byref(eps), _dp(minv))
        return dict(n_grad=ng, draws=draws, outputs=outs, sampler_params=sp, stepsize=eps.value, inv_metric=minv)


def nuts_screen(problems, n_chains=4, num_warmup=1000, num_samples=1000, seed=1, adapt_delta=0.9,
                want_outputs=True, nthreads=0, metric=0, stepsize_jitter=0.0):
    """Batch of problems x chains over host threads.  Returns (n_grad_total, outputs, sampler_params)."""
    a = OrcNutsArgs()
    lib().orc_nuts_defaults(C.byref(a))
    a.num_warmup, a.num_samples, a.seed, a.adapt_delta = num_warmup, num_samples, seed, adapt_delta
    a.metric, a.stepsize_jitter = metric, stepsize_jitter
    arr = (OrcProblem * len(problems))(*[p.c for p in problems])
    stride = max(p.n_out for p in problems)
    outs = np.full((len(problems), n_chains, num_samples, stride), np.nan) if want_outputs else None
    sp = np.empty((len(problems), n_chains, num_samples, 7))
    ng = lib().orc_nuts_screen(arr, len(problems), n_chains, C.byref(a), stride, _dp(outs), _dp(sp), nthreads)
    return ng, outs, sp


ADVI_INFO = ("status", "eta", "iterations", "elbo", "converged", "n_grad", "n_lp", "_")


def advi_screen(problems, seed=1, want_outputs=True, nthreads=0, **kw):
    """Full-rank ADVI (rstan::vb defaults unless overridden in kw) for a batch of problems over host threads.
    Returns (mean_u[n, maxD], outputs[n, output_samples, max_out] or None, info[n, 8])."""
    a = OrcAdviArgs()
    lib().orc_advi_defaults(C.byref(a))
    a.seed = seed
    for k, v in kw.items():
        setattr(a, k, v)
    arr = (OrcProblem * len(problems))(*[p.c for p in problems])
    ld_u = max(p.dim for p in problems)
    stride = max(p.n_out for p in problems)
    mean = np.full((len(problems), ld_u), np.nan)
    outs = np.full((len(problems), a.output_samples, stride), np.nan) if want_outputs else None
    info = np.zeros((len(problems), 8))
    lib().orc_advi_screen(arr, len(problems), C.byref(a), ld_u, stride, _dp(mean), _dp(outs), _dp(info), nthreads)
    return mean, outs, info
