"""ctypes binding of the C ABI in include/bsyn.h (libbsyn.so, built in-tree by __graft_entry__.build()).

The product path has no CPU fallback: if the CUDA library is missing or no device is present, every
compute call raises.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("BSYN_LIB", os.path.join(_HERE, "lib", "libbsyn.so"))   # BSYN_LIB: A/B builds

BSYN_NQ = 12
BSYN_NSP = 7
Q_NAMES = ["ec50_1", "ec50_2", "dss_1", "dss_2", "s", "rVUS_f", "rVUS_p0", "VUS_Delta", "VUS_syn", "VUS_ant",
           "lp__", "_reserved"]
SP_NAMES = ["accept_stat__", "stepsize__", "treedepth__", "n_leapfrog__", "divergent__", "energy__", "lp__"]

EXPORTS = [
    "bsyn_batch_create", "bsyn_batch_destroy", "bsyn_last_error", "bsyn_version", "bsyn_num_params",
    "bsyn_num_outputs", "bsyn_num_param_blocks", "bsyn_param_block", "bsyn_max_params", "bsyn_max_outputs", "bsyn_num_problems", "bsyn_logp_grad",
    "bsyn_logp_grad_device", "bsyn_write_array", "bsyn_unconstrain", "bsyn_sampler_defaults", "bsyn_sample",
    "bsyn_sample_device", "bsyn_ess", "bsyn_bulk_ess", "bsyn_sample_to_ess", "bsyn_ess_target_defaults", "bsyn_launch_count", "bsyn_advi_defaults", "bsyn_advi", "bsyn_prepare", "bsyn_output_name",
]
VI_NAMES = ["status", "eta", "iterations", "elbo", "converged", "n_grad", "n_lp", "_reserved"]


class BsynProblem(C.Structure):
    _fields_ = [("n1", C.c_int32), ("n2", C.c_int32), ("nmissing", C.c_int32), ("nrep", C.c_int32),
                ("N", C.c_int32), ("y", C.POINTER(C.c_double)), ("ii_obs", C.POINTER(C.c_int32)),
                ("x1", C.POINTER(C.c_double)), ("x2", C.POINTER(C.c_double))]


class BsynOptions(C.Structure):
    _fields_ = [("model", C.c_int32), ("est_la", C.c_int32), ("robust", C.c_int32), ("pcprior", C.c_int32),
                ("heteroscedastic", C.c_int32), ("kernel", C.c_int32), ("nu_matern", C.c_int32),
                ("est_alpha", C.c_int32), ("pcprior_hypers", C.c_double * 4), ("rho", C.c_double),
                ("lambda_", C.c_double)]


class BsynSamplerArgs(C.Structure):
    _fields_ = [("chains", C.c_int32), ("iter", C.c_int32), ("warmup", C.c_int32), ("thin", C.c_int32),
                ("seed", C.c_uint64), ("init_r", C.c_double), ("adapt_delta", C.c_double),
                ("stepsize", C.c_double), ("stepsize_jitter", C.c_double), ("max_treedepth", C.c_int32),
                ("metric", C.c_int32), ("adapt_gamma", C.c_double), ("adapt_kappa", C.c_double),
                ("adapt_t0", C.c_double), ("adapt_init_buffer", C.c_int32), ("adapt_term_buffer", C.c_int32),
                ("adapt_window", C.c_int32), ("save_warmup", C.c_int32)]


class BsynSinks(C.Structure):
    _fields_ = [("draws", C.POINTER(C.c_double)), ("ld_draws", C.c_int64),
                ("udraws", C.POINTER(C.c_double)), ("ld_udraws", C.c_int64),
                ("qdraws", C.POINTER(C.c_double)), ("sampler_params", C.POINTER(C.c_double)),
                ("summary", C.POINTER(C.c_double)), ("cpo_mean", C.POINTER(C.c_double)), ("ld_cpo", C.c_int64),
                ("chain_stats", C.POINTER(C.c_double)), ("inv_metric", C.POINTER(C.c_double)),
                ("init_u", C.POINTER(C.c_double)), ("f_mean", C.POINTER(C.c_double)), ("ld_f", C.c_int64),
                ("surfaces", C.POINTER(C.c_double)), ("lpml", C.POINTER(C.c_double)),
                ("bridge_logml", C.POINTER(C.c_double)), ("bridge_niter", C.POINTER(C.c_int32)),
                ("elapsed_ms", C.POINTER(C.c_float))]


class BsynEssTarget(C.Structure):
    _fields_ = [("target_ess", C.c_double), ("q1", C.c_int32), ("q2", C.c_int32), ("max_rounds", C.c_int32),
                ("retry_adapt_delta", C.c_double), ("retry_warmup", C.c_int32), ("retry_rounds", C.c_int32)]


class BsynEssRun(C.Structure):
    _fields_ = [("total_grad_evals", C.c_int64), ("elapsed_ms", C.c_float), ("phase1_ms", C.c_float),
                ("rounds", C.c_int32), ("n_retried", C.c_int32)]


class BsynAdviArgs(C.Structure):
    _fields_ = [("iter", C.c_int32), ("grad_samples", C.c_int32), ("elbo_samples", C.c_int32),
                ("eval_elbo", C.c_int32), ("eta", C.c_double), ("adapt_engaged", C.c_int32),
                ("adapt_iter", C.c_int32), ("tol_rel_obj", C.c_double), ("output_samples", C.c_int32),
                ("init_r", C.c_double), ("seed", C.c_uint64)]


class BsynAdviSinks(C.Structure):
    _fields_ = [("mean", C.POINTER(C.c_double)), ("ld_mean", C.c_int64), ("info", C.POINTER(C.c_double)),
                ("draws", C.POINTER(C.c_double)), ("ld_draws", C.c_int64),
                ("udraws", C.POINTER(C.c_double)), ("ld_udraws", C.c_int64),
                ("qdraws", C.POINTER(C.c_double)), ("summary", C.POINTER(C.c_double)),
                ("cpo_mean", C.POINTER(C.c_double)), ("ld_cpo", C.c_int64)]


class BsynError(RuntimeError):
    pass


_LIB = None


def load():
    """Load libbsyn.so.  Raises BsynError (no fallback) when the library has not been built."""
    global _LIB
    if _LIB is not None:
        return _LIB
    if not os.path.exists(LIB_PATH):
        raise BsynError(f"{LIB_PATH} not found: run `python -c 'import __graft_entry__ as g; g.build()'` "
                        "(there is no CPU fallback for the hot path)")
    L = C.CDLL(LIB_PATH)
    dp, ip = C.POINTER(C.c_double), C.POINTER(C.c_int32)
    vp = C.c_void_p
    L.bsyn_batch_create.argtypes = [C.POINTER(BsynProblem), C.c_int32, C.POINTER(BsynOptions), C.c_int32, C.POINTER(vp)]
    L.bsyn_batch_destroy.argtypes = [vp]
    L.bsyn_batch_destroy.restype = None
    L.bsyn_last_error.restype = C.c_char_p
    for f in ("bsyn_num_params", "bsyn_num_outputs"):
        getattr(L, f).argtypes = [vp, C.c_int32]
    for f in ("bsyn_max_params", "bsyn_max_outputs", "bsyn_num_problems"):
        getattr(L, f).argtypes = [vp]
    L.bsyn_logp_grad.argtypes = [vp, C.c_int32, ip, dp, C.c_int32, C.c_int32, dp, dp, ip]
    L.bsyn_logp_grad_debug.argtypes = [vp, C.c_int32, ip, dp, C.c_int32, C.c_int32, dp, dp, ip, dp, C.c_int32]
    L.bsyn_debug_layout.argtypes = [vp, ip, ip, ip]
    L.bsyn_logp_grad_device.argtypes = [vp, C.c_int32, vp, vp, C.c_int32, C.c_int32, vp, vp, vp, vp]
    L.bsyn_write_array.argtypes = [vp, C.c_int32, ip, dp, C.c_int32, dp, C.c_int32, C.c_uint64, ip]
    L.bsyn_unconstrain.argtypes = [vp, C.c_int32, dp, dp]
    L.bsyn_sampler_defaults.argtypes = [C.POINTER(BsynSamplerArgs)]
    L.bsyn_sampler_defaults.restype = None
    L.bsyn_sample.argtypes = [vp, C.POINTER(BsynSamplerArgs), C.POINTER(BsynSinks), C.POINTER(C.c_int64)]
    L.bsyn_sample_device.argtypes = [vp, C.POINTER(BsynSamplerArgs), C.POINTER(C.c_int64), C.POINTER(C.c_float),
                                     C.POINTER(vp), C.POINTER(vp), C.POINTER(vp)]
    L.bsyn_ess.argtypes = [vp, C.c_int32, dp, dp]
    L.bsyn_bulk_ess.argtypes = [vp, C.c_int32, dp, dp, dp]
    L.bsyn_ess_target_defaults.argtypes = [C.POINTER(BsynEssTarget)]
    L.bsyn_ess_target_defaults.restype = None
    L.bsyn_sample_to_ess.argtypes = [vp, C.POINTER(BsynSamplerArgs), C.POINTER(BsynEssTarget), C.POINTER(BsynSinks), dp, ip,
                                     C.POINTER(BsynEssRun)]
    i32p = C.POINTER(C.c_int32)
    L.bsyn_prepare.argtypes = [dp, dp, C.c_int32, C.c_int32, dp, i32p, dp, i32p, dp, ip, i32p, i32p, i32p]
    L.bsyn_output_name.argtypes = [vp, C.c_int32, C.c_int32, C.c_char_p, C.c_int32]
    L.bsyn_num_param_blocks.argtypes = [vp]
    L.bsyn_param_block.argtypes = [vp, C.c_int32, C.c_int32, C.c_char_p, C.c_int32, i32p, i32p, i32p]
    L.bsyn_advi_defaults.argtypes = [C.POINTER(BsynAdviArgs)]
    L.bsyn_advi_defaults.restype = None
    L.bsyn_advi.argtypes = [vp, C.POINTER(BsynAdviArgs), C.POINTER(BsynAdviSinks), C.POINTER(C.c_int64),
                            C.POINTER(C.c_float)]
    L.bsyn_launch_count.restype = C.c_int64
    _LIB = L
    return L


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double)) if a is not None else None


def _ip(a):
    return a.ctypes.data_as(C.POINTER(C.c_int32)) if a is not None else None


def _check(rc, what):
    if rc != 0:
        raise BsynError(f"{what} failed (code {rc}): {load().bsyn_last_error().decode()}")


def options_from(sd, model="gp_grid") -> BsynOptions:
    o = BsynOptions()
    o.model = 0 if model == "gp_grid" else 1
    o.est_la, o.robust, o.pcprior, o.heteroscedastic = sd.est_la, sd.robust, sd.pcprior, sd.heteroscedastic
    o.kernel, o.nu_matern, o.est_alpha = sd.kernel, sd.nu_matern, sd.est_alpha
    for k in range(4):
        o.pcprior_hypers[k] = float(sd.pcprior_hypers[k])
    o.rho, o.lambda_ = sd.rho, sd.lambda_
    return o


def prepare_native(y, x):
    """bsyn_prepare: (x1, x2, y_obs, ii_obs, nrep, nmissing) from y [n] or [n, nrep] (NaN = missing) and x [n, 2]."""
    y = np.asarray(y, dtype=np.float64)
    x = np.ascontiguousarray(x, dtype=np.float64)
    if y.ndim == 1:
        y = y[:, None]
    n, nrep = y.shape
    yc = np.ascontiguousarray(y.T)                   # column-major: replicate r of row i at [r * n + i]
    x1, x2 = np.empty(n), np.empty(n)
    yo, io = np.empty(n * nrep), np.empty(n * nrep, dtype=np.int32)
    n1, n2, N, nr, nm = (C.c_int32() for _ in range(5))
    rc = load().bsyn_prepare(_dp(yc), _dp(x), n, nrep, _dp(x1), C.byref(n1), _dp(x2), C.byref(n2), _dp(yo),
                             io.ctypes.data_as(C.POINTER(C.c_int32)), C.byref(N), C.byref(nr), C.byref(nm))
    if rc == 5:
        raise ValueError("Missing monotherapy data! Make sure 'x' contains zero concentrations")
    if rc != 0:
        raise ValueError("bsyn_prepare: invalid concentrations")
    return x1[:n1.value].copy(), x2[:n2.value].copy(), yo[:N.value].copy(), io[:N.value].copy(), nr.value, nm.value


class Batch:
    """A set of experiments packed on one GPU (handle of bsyn_batch_create)."""

    def __init__(self, stan_data_list, model="gp_grid", device=0):
        L = load()
        self._keep = []
        n = len(stan_data_list)
        arr = (BsynProblem * n)()
        for k, sd in enumerate(stan_data_list):
            y = np.ascontiguousarray(sd.y, dtype=np.float64)
            ii = np.ascontiguousarray(sd.ii_obs, dtype=np.int32)
            x1 = np.ascontiguousarray(sd.x1, dtype=np.float64)
            x2 = np.ascontiguousarray(sd.x2, dtype=np.float64)
            self._keep += [y, ii, x1, x2]
            p = arr[k]
            p.n1, p.n2, p.nmissing, p.nrep, p.N = sd.n1, sd.n2, sd.nmissing, sd.nrep, y.shape[0]
            p.y, p.ii_obs, p.x1, p.x2 = _dp(y), _ip(ii), _dp(x1), _dp(x2)
        sd0 = stan_data_list[0]
        for sd in stan_data_list[1:]:
            same = all(getattr(sd, f) == getattr(sd0, f) for f in
                       ("est_la", "robust", "pcprior", "heteroscedastic", "kernel", "nu_matern", "est_alpha", "rho",
                        "lambda_")) and np.array_equal(sd.pcprior_hypers, sd0.pcprior_hypers)
            if not same:
                raise ValueError("all experiments of a batch must share the model options")
        opts = options_from(sd0, model)
        h = C.c_void_p()
        _check(L.bsyn_batch_create(arr, n, C.byref(opts), device, C.byref(h)), "bsyn_batch_create")
        self.h = h
        self.model = model
        self.n = n
        self.device = device
        self.max_params = L.bsyn_max_params(h)
        self.max_outputs = L.bsyn_max_outputs(h)
        self.dims = [L.bsyn_num_params(h, k) for k in range(n)] if n <= 4096 else None
        self.n_obs = [int(np.asarray(sd.y).shape[0]) for sd in stan_data_list]
        self.n_cells = [(sd.n1 + 1) * (sd.n2 + 1) for sd in stan_data_list]

    def close(self):
        if getattr(self, "h", None):
            load().bsyn_batch_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def num_params(self, k=0):
        return load().bsyn_num_params(self.h, k)

    def num_outputs(self, k=0):
        return load().bsyn_num_outputs(self.h, k)

    # -- log_prob / grad_log_prob -----------------------------------------------------------------
    def output_names(self, k=0):
        """Flat names of one write_array row of experiment k, from the library (bsyn_output_name)."""
        buf = C.create_string_buffer(48)
        out = []
        for i in range(self.num_outputs(k)):
            _check(load().bsyn_output_name(self.h, k, i, buf, 48), "bsyn_output_name")
            out.append(buf.value.decode())
        return out

    def logp_grad(self, u, prob_idx=None, jacobian=True, want_grad=True, debug=False):
        """u: [B, >=D].  Returns (lp[B], grad[B, ld] or None, status[B][, dbg])."""
        u = np.atleast_2d(np.asarray(u, dtype=np.float64))
        B = u.shape[0]
        ld = max(u.shape[1], self.max_params)
        ub = np.zeros((B, ld))
        ub[:, :u.shape[1]] = u
        idx = np.zeros(B, dtype=np.int32) if prob_idx is None else np.ascontiguousarray(prob_idx, dtype=np.int32)
        lp = np.empty(B)
        g = np.zeros((B, ld)) if want_grad else None
        st = np.empty(B, dtype=np.int32)
        L = load()
        if debug:
            nm, nn, gm = C.c_int32(), C.c_int32(), C.c_int32()
            L.bsyn_debug_layout(self.h, C.byref(nm), C.byref(nn), C.byref(gm))
            stride = 6 * nn.value + 4 * gm.value
            dbg = np.zeros((B, stride))
            _check(L.bsyn_logp_grad_debug(self.h, B, _ip(idx), _dp(ub), ld, int(jacobian), _dp(lp), _dp(g), _ip(st),
                                          _dp(dbg), stride), "bsyn_logp_grad_debug")
            return lp, g, st, dict(raw=dbg, NM=nm.value, NN=nn.value, GM=gm.value)
        _check(L.bsyn_logp_grad(self.h, B, _ip(idx), _dp(ub), ld, int(jacobian), _dp(lp), _dp(g), _ip(st)),
               "bsyn_logp_grad")
        return lp, g, st

    def write_array(self, u, prob_idx=None, seed=1):
        u = np.atleast_2d(np.asarray(u, dtype=np.float64))
        B = u.shape[0]
        ld = max(u.shape[1], self.max_params)
        ub = np.zeros((B, ld))
        ub[:, :u.shape[1]] = u
        idx = np.zeros(B, dtype=np.int32) if prob_idx is None else np.ascontiguousarray(prob_idx, dtype=np.int32)
        out = np.zeros((B, self.max_outputs))
        st = np.empty(B, dtype=np.int32)
        _check(load().bsyn_write_array(self.h, B, _ip(idx), _dp(ub), ld, _dp(out), self.max_outputs, seed, _ip(st)),
               "bsyn_write_array")
        return out, st

    def unconstrain(self, theta, k=0):
        theta = np.ascontiguousarray(theta, dtype=np.float64)
        u = np.empty(self.num_params(k))
        _check(load().bsyn_unconstrain(self.h, k, _dp(theta), _dp(u)), "bsyn_unconstrain")
        return u

    # -- sampling ---------------------------------------------------------------------------------
    @staticmethod
    def sampler_args(**kw) -> BsynSamplerArgs:
        a = BsynSamplerArgs()
        load().bsyn_sampler_defaults(C.byref(a))
        for k, v in kw.items():
            if not hasattr(a, k):
                raise TypeError(f"unknown sampler argument {k!r}")
            setattr(a, k, v)
        return a

    def param_blocks(self, k=0):
        """[(name, dims, flat_offset)]: what param_names() / param_dims() of the rstan module report."""
        L = load()
        out, buf = [], C.create_string_buffer(48)
        nd, off, dims = C.c_int32(), C.c_int32(), (C.c_int32 * 2)()
        for i in range(L.bsyn_num_param_blocks(self.h)):
            _check(L.bsyn_param_block(self.h, k, i, buf, 48, C.byref(nd), dims, C.byref(off)), "bsyn_param_block")
            out.append((buf.value.decode(), [dims[j] for j in range(nd.value)], off.value))
        return out

    def sample(self, want_draws=False, want_udraws=False, want_qdraws=True, want_sampler_params=True,
               want_cpo=False, want_inv_metric=False, want_inits=False, want_f_mean=False, want_surfaces=False, want_bridge=False,
               **kw):
        """Run adaptive NUTS for every experiment x chain in one launch.  Returns a dict of numpy arrays."""
        a = self.sampler_args(**kw)
        ns = (a.iter - a.warmup + a.thin - 1) // a.thin + ((a.warmup + a.thin - 1) // a.thin if a.save_warmup else 0)
        n, ch = self.n, a.chains
        s = BsynSinks()
        res = {}
        if want_draws:
            res["draws"] = np.zeros((n, ch, ns, self.max_outputs))
            s.draws, s.ld_draws = _dp(res["draws"]), self.max_outputs
        if want_udraws:
            res["udraws"] = np.zeros((n, ch, ns, self.max_params))
            s.udraws, s.ld_udraws = _dp(res["udraws"]), self.max_params
        if want_qdraws:
            res["qdraws"] = np.zeros((n, ch, ns, BSYN_NQ))
            s.qdraws = _dp(res["qdraws"])
        if want_sampler_params:
            res["sampler_params"] = np.zeros((n, ch, ns, BSYN_NSP))
            s.sampler_params = _dp(res["sampler_params"])
        res["summary"] = np.zeros((n, BSYN_NQ, 2))
        s.summary = _dp(res["summary"])
        res["chain_stats"] = np.zeros((n, ch, 8))
        s.chain_stats = _dp(res["chain_stats"])
        if want_cpo:
            ldc = max(self.n_obs)
            res["cpo_mean"] = np.zeros((n, ldc))
            s.cpo_mean, s.ld_cpo = _dp(res["cpo_mean"]), ldc
        if want_inv_metric:
            res["inv_metric"] = np.zeros((n, ch, self.max_params))
            s.inv_metric = _dp(res["inv_metric"])
        if want_inits:
            res["init_u"] = np.zeros((n, ch, self.max_params))
            s.init_u = _dp(res["init_u"])
        if want_f_mean or want_surfaces:
            s.ld_f = max(self.n_cells)
        if want_f_mean:
            res["f_mean"] = np.zeros((n, s.ld_f))
            s.f_mean = _dp(res["f_mean"])
        if want_surfaces:          # f (mean / median if robust), p0 (median), Delta (median) per grid cell, and LPML
            res["surfaces"] = np.zeros((n, 3, s.ld_f))
            res["lpml"] = np.zeros(n)
            s.surfaces, s.lpml = _dp(res["surfaces"]), _dp(res["lpml"])
        if want_bridge:            # bridge-sampling log marginal likelihood per experiment, on the device
            res["bridge_logml"] = np.zeros(n)
            res["bridge_niter"] = np.zeros(n, dtype=np.int32)
            s.bridge_logml, s.bridge_niter = _dp(res["bridge_logml"]), _ip(res["bridge_niter"])
        ems = C.c_float()
        s.elapsed_ms = C.pointer(ems)
        ng = C.c_int64()
        _check(load().bsyn_sample(self.h, C.byref(a), C.byref(s), C.byref(ng)), "bsyn_sample")
        res["n_grad"] = ng.value
        res["n_save"] = ns
        res["kernel_ms"] = ems.value
        return res

    def advi(self, want_draws=False, want_udraws=False, want_qdraws=True, want_cpo=False, **kw):
        """Full-rank ADVI (rstan::vb(algorithm="fullrank") defaults unless overridden in kw) for every experiment in
        one launch.  Returns a dict of numpy arrays; `info[:, 0]` != 0 where the reference would have thrown."""
        a = BsynAdviArgs()
        load().bsyn_advi_defaults(C.byref(a))
        for k, v in kw.items():
            if not hasattr(a, k):
                raise TypeError(f"unknown ADVI argument {k!r}")
            setattr(a, k, v)
        n, ns = self.n, a.output_samples
        s = BsynAdviSinks()
        res = {"mean": np.zeros((n, self.max_params)), "info": np.zeros((n, 8)), "summary": np.zeros((n, BSYN_NQ, 2))}
        s.mean, s.ld_mean, s.info, s.summary = _dp(res["mean"]), self.max_params, _dp(res["info"]), _dp(res["summary"])
        if want_draws:
            res["draws"] = np.zeros((n, 1, ns, self.max_outputs))
            s.draws, s.ld_draws = _dp(res["draws"]), self.max_outputs
        if want_udraws:
            res["udraws"] = np.zeros((n, 1, ns, self.max_params))
            s.udraws, s.ld_udraws = _dp(res["udraws"]), self.max_params
        if want_qdraws:
            res["qdraws"] = np.zeros((n, 1, ns, BSYN_NQ))
            s.qdraws = _dp(res["qdraws"])
        if want_cpo:
            ldc = max(self.n_obs)
            res["cpo_mean"] = np.zeros((n, ldc))
            s.cpo_mean, s.ld_cpo = _dp(res["cpo_mean"]), ldc
        ne, ms = C.c_int64(), C.c_float()
        _check(load().bsyn_advi(self.h, C.byref(a), C.byref(s), C.byref(ne), C.byref(ms)), "bsyn_advi")
        res["n_evals"], res["kernel_ms"], res["n_save"] = ne.value, ms.value, ns
        return res

    def sample_device(self, **kw):
        """Sampler with all results left on the device; returns (n_grad, kernel_ms)."""
        a = self.sampler_args(**kw)
        ng, ms = C.c_int64(), C.c_float()
        _check(load().bsyn_sample_device(self.h, C.byref(a), C.byref(ng), C.byref(ms), None, None, None),
               "bsyn_sample_device")
        return ng.value, ms.value

    def sample_to_ess(self, target=400.0, q=("VUS_syn", "VUS_ant"), block=150, max_rounds=6, retry_rounds=6,
                      retry_adapt_delta=0.99, retry_warmup=150, want_summary=False, want_qdraws=False, **kw):
        """Sample every experiment until min over `q` of its rank-normalised bulk ESS reaches `target`
        (bsyn_sample_to_ess): warm-up + `block` draws per chain, then extension launches of `block` draws over the
        experiments still below the target (phase 1, up to max_rounds launches); what is left re-adapts its step size to
        `retry_adapt_delta` and is extended up to `retry_rounds` more times (phase 2).  kw: sampler arguments."""
        kw = dict(kw)
        kw.setdefault("warmup", 1000)
        kw["iter"] = kw["warmup"] + block
        a = self.sampler_args(**kw)
        t = BsynEssTarget()
        load().bsyn_ess_target_defaults(C.byref(t))
        qi = [Q_NAMES.index(x) if isinstance(x, str) else int(x) for x in q]
        t.target_ess, t.q1, t.q2, t.max_rounds = float(target), qi[0], qi[1], max_rounds
        t.retry_adapt_delta, t.retry_warmup, t.retry_rounds = retry_adapt_delta, retry_warmup, retry_rounds
        n, ch, ld = self.n, a.chains, block * (max_rounds + retry_rounds)
        s = BsynSinks()
        res = {}
        if want_summary:
            res["summary"] = np.zeros((n, BSYN_NQ, 2))
            s.summary = _dp(res["summary"])
            res["chain_stats"] = np.zeros((n, ch, 8))
            s.chain_stats = _dp(res["chain_stats"])
        if want_qdraws:
            res["qdraws"] = np.zeros((n, ch, ld, BSYN_NQ))
            s.qdraws = _dp(res["qdraws"])
        ess = np.empty(n)
        nd = np.empty(n, dtype=np.int32)
        run = BsynEssRun()
        _check(load().bsyn_sample_to_ess(self.h, C.byref(a), C.byref(t), C.byref(s), _dp(ess), _ip(nd), C.byref(run)),
               "bsyn_sample_to_ess")
        res.update(ess=ess, draws_per_chain=nd, n_grad=run.total_grad_evals, kernel_ms=run.elapsed_ms,
                   phase1_ms=run.phase1_ms, rounds=run.rounds, n_retried=run.n_retried)
        return res

    def bulk_ess(self, q):
        """Rank-normalised bulk ESS, R-hat and tail ESS per experiment, on the device (bsyn_bulk_ess)."""
        qi = Q_NAMES.index(q) if isinstance(q, str) else int(q)
        ess, rhat, tail = np.empty(self.n), np.empty(self.n), np.empty(self.n)
        _check(load().bsyn_bulk_ess(self.h, qi, _dp(ess), _dp(rhat), _dp(tail)), "bsyn_bulk_ess")
        return ess, rhat, tail

    def ess(self, q):
        qi = Q_NAMES.index(q) if isinstance(q, str) else int(q)
        ess = np.empty(self.n)
        rhat = np.empty(self.n)
        _check(load().bsyn_ess(self.h, qi, _dp(ess), _dp(rhat)), "bsyn_ess")
        return ess, rhat
