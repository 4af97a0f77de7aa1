"""Host-side mirror of the reference's user API for the gp_grid / nointeraction path.

`bayesynergy()`  <-> R/bayesynergy.R:54-343      (one experiment; same arguments, same result fields)
`synergyscreen()` <-> R/synergyscreen.R:58-463   (a list of experiments; same screenSummary columns)

The reference is R; R is not in this image, so the host side above the C ABI is Python with the same
names, argument meaning and error behaviour (INTEGRATION.md shows the Rcpp binding a maintainer would add).
Everything inside the sampling loop runs on the GPU through bayesynergy_b200._lib (no CPU fallback).
"""
from __future__ import annotations

import warnings

import numpy as np

from . import _lib
from .dataprep import prepare, output_names, param_names
from .diagnostics import split_ess_rhat

SCREEN_COLUMNS = [
    "listID", "Experiment ID", "Drug A", "Drug B", "EC50 (Drug A)", "EC50 (Drug B)", "DSS (Drug A)", "DSS (Drug B)",
    "Synergy (mean)", "Synergy (sd)", "Antagonism (mean)", "Antagonism (sd)", "Efficacy (mean)", "Efficacy (sd)",
    "Non-interaction Efficacy (mean)", "Non-interaction Efficacy (sd)", "Interaction (mean)", "Interaction (sd)",
    "Synergy Score", "Antagonism Score", "s", "returnCode", "divergentTransitions",
]
Q = _lib.Q_NAMES


def _control_defaults(type, robust, control):
    """R/bayesynergy.R:106-116."""
    c = dict(adapt_delta=0.8)
    if type != 1:
        c = dict(adapt_delta=0.9)
    if robust:
        c["stepsize_jitter"] = 0.5
        c["metric"] = "dense_e"
    c.update(control or {})
    return c


def _sampler_kwargs(control, chains, iter, warmup, thin, seed):
    kw = dict(chains=chains, iter=iter, warmup=iter // 2 if warmup is None else warmup, thin=thin, seed=seed)
    m = dict(adapt_delta="adapt_delta", stepsize="stepsize", stepsize_jitter="stepsize_jitter",
             max_treedepth="max_treedepth", adapt_gamma="adapt_gamma", adapt_kappa="adapt_kappa",
             adapt_t0="adapt_t0", adapt_init_buffer="adapt_init_buffer", adapt_term_buffer="adapt_term_buffer",
             adapt_window="adapt_window")
    for k, v in control.items():
        if k == "metric":       # "diag_e" (default) or "dense_e" (the reference's choice when robust=TRUE)
            if v not in ("diag_e", "dense_e"):
                raise ValueError("control$metric must be 'diag_e' or 'dense_e'")
            kw["metric"] = 1 if v == "dense_e" else 0
            continue
        if k not in m:
            raise ValueError(f"unknown control argument {k!r}")
        kw[m[k]] = v
    return kw


_VB_ERRORS = {
    1: "Initialization failed.",
    2: "Cannot compute ELBO using the initial variational distribution. Your model may be either severely "
       "ill-conditioned or misspecified.",
    3: "All proposed step-sizes failed. Your model may be either severely ill-conditioned or misspecified.",
    4: "stan::variational::normal_fullrank::calc_grad: The number of dropped evaluations has reached its maximum "
       "amount (3). Your model may be either severely ill-conditioned or misspecified.",
    5: "stan::variational::advi::calc_ELBO: The number of dropped evaluations has reached its maximum amount. "
       "Your model may be either severely ill-conditioned or misspecified.",
}


def _diagnose(sp, chain_stats, qdraws, max_treedepth):
    """rstan-style warnings that make bayesynergy() set returnCode = 1 (R/bayesynergy.R:210-214)."""
    msgs = []
    ndiv = int(sp[..., 4].sum())
    if ndiv > 0:
        msgs.append(f"There were {ndiv} divergent transitions after warmup.")
    nmax = int((sp[..., 2] >= max_treedepth).sum())
    if nmax > 0:
        msgs.append(f"There were {nmax} transitions after warmup that exceeded the maximum treedepth.")
    if chain_stats[..., 6].any():
        msgs.append("Initialization failed for at least one chain.")
    for qn in ("VUS_syn", "VUS_ant", "s"):
        ess, rhat = split_ess_rhat(qdraws[:, :, Q.index(qn)])
        if rhat == rhat and rhat > 1.05:
            msgs.append(f"The largest R-hat is {rhat:.2f}, indicating chains have not mixed.")
            break
        if ess == ess and ess < 100 * qdraws.shape[0]:
            msgs.append("Bulk Effective Samples Size (ESS) is too low, indicating posterior means and medians may be unreliable.")
            break
    return msgs, ndiv


def _posterior_from_draws(sd, draws, lp, model="gp_grid"):
    """rstan::extract(fit) from the write_array rows draws[chains, n_save, n_out] (chains pooled, as extract does)."""
    names = output_names(sd, model)
    flat = draws[:, :, :len(names)].reshape(-1, len(names))
    n_save = flat.shape[0]
    col = {nm: i for i, nm in enumerate(names)}
    n1, n2, N = sd.n1, sd.n2, sd.N

    def block(first, count):
        return flat[:, col[first]:col[first] + count]

    def grid(first):
        return block(first, n1 * n2).reshape(n_save, n1, n2).transpose(0, 2, 1)

    posterior = {}
    for nm in param_names(sd, model):
        base = nm.split("[")[0]
        if base != "z":
            posterior.setdefault(base, flat[:, col[nm]])
    posterior["p0"] = grid("p0[1,1]")
    posterior["p01"] = block("p01[1]", n1)
    posterior["p02"] = block("p02[1]", n2)
    if model == "gp_grid":
        posterior["z"] = grid("z[1,1]")
        posterior["Delta"] = grid("Delta[1,1]")
        posterior["GP"] = grid("GP[1,1]")
    posterior["CPO"] = block("CPO[1]", N)
    for nm in ("ec50_1", "ec50_2", "dss_1", "dss_2", "rVUS_f", "rVUS_p0", "VUS_Delta", "VUS_syn", "VUS_ant"):
        if nm in col:
            posterior[nm] = flat[:, col[nm]]
    posterior["lp__"] = lp
    return posterior


def _surface_grids(sd, surfaces_k):
    """[3, ld_f] device surfaces -> f, p0, Delta as (n2+1) x (n1+1) matrices (cells are column-major on the device)."""
    ncf = (sd.n1 + 1) * (sd.n2 + 1)
    return tuple(surfaces_k[i, :ncf].reshape(sd.n1 + 1, sd.n2 + 1).T.copy() for i in range(3))


def _outlier_message(sd, f_cells, s_mean, lambda_):
    """The residual rule of R/bayesynergy.R:305-314 on the posterior surface of f (column-major cell vector)."""
    fo = f_cells[sd.ii_obs - 1]
    resid = sd.y - fo
    point_sd = np.sqrt(s_mean ** 2 * (fo + lambda_))
    if (np.abs(resid) > 3 * point_sd).any():
        return (f"Largest residuals from posterior median is {resid.max():.4g}, which is more than three times the "
                "observation noise. This could indicate the presence of an outlier.")
    return None


def _assemble_object(sd, meta, posterior, surfaces_k, lpml, sp, chain_stats, qdraws, options, max_treedepth):
    """The fields of the reference's `bayesynergy` object (R/bayesynergy.R:244-343) for one experiment.  The posterior
    surfaces (mean / median per dose-grid cell) and LPML come reduced from the device (bsyn_sample's `surfaces` / `lpml`
    sinks), not from the draws on the host."""
    f_mean, p0_mean, Delta_mean = _surface_grids(sd, surfaces_k)
    skip = {"z", "p0", "p01", "p02", "Delta", "CPO", "lp__"}
    posterior_mean = {k: (float(v.mean()) if v.ndim == 1 else v.mean(axis=0)) for k, v in posterior.items() if k not in skip}
    posterior_mean.update(f=f_mean, p0=p0_mean, Delta=Delta_mean)
    if options["method"] == "sampling":
        messages, ndiv = _diagnose(sp, chain_stats, qdraws, max_treedepth)
    else:
        messages, ndiv = [], float("nan")
    returnCode = 1 if messages else 0
    msg = _outlier_message(sd, f_mean.T.reshape(-1), posterior_mean["s"], options["lambda_"])
    if msg:
        warnings.warn(msg)
        returnCode = 1
        messages.append(msg)
    model = dict(type=options["type"], lower_asymptotes=options["lower_asymptotes"], method=options["method"],
                 bayes_factor=options["bayes_factor"], heteroscedastic=options["heteroscedastic"],
                 robust=options["robust"], pcprior=options["pcprior"], lambda_=options["lambda_"])
    if options["type"] == 3:
        model.update(nu=options["nu"], pcprior_hypers=tuple(options["pcprior_hypers"]))
    obj = dict(posterior=posterior, posterior_mean=posterior_mean,
               data=dict(y=meta["y"], x=meta["x"], drug_names=meta["drug_names"], experiment_ID=meta["experiment_ID"],
                         units=meta["units"], indices=sd.ii_obs),
               model=model, returnCode=returnCode, LPML=float(lpml), divergent=float("nan"))
    if returnCode:
        obj["messages"] = messages
        obj["divergent"] = ndiv
    return obj


def bayesynergy(y, x, type=3, drug_names=None, experiment_ID=None, units=None, lower_asymptotes=True,
                heteroscedastic=True, bayes_factor=False, robust=False, rho=0.90, pcprior=False,
                pcprior_hypers=(1, 0.1, 1, 0.2), lambda_=0.005, nu=1.5, method="sampling", control=None,
                chains=4, iter=2000, warmup=None, thin=1, seed=1, device=0, vb=None):
    """Fit one experiment (R/bayesynergy.R:54).  Returns a dict with the fields of the R object:
    posterior (= rstan::extract), posterior_mean, data, model, returnCode, LPML, divergent, messages.

    method="vb" runs full-rank ADVI (R/bayesynergy.R:225-231) with rstan::vb's defaults; `vb` is a dict of its
    arguments (iter, elbo_samples, eval_elbo, eta, adapt_engaged, adapt_iter, tol_rel_obj, output_samples).  A fit
    that rstan::vb would abort with an error raises RuntimeError here."""
    if method not in ("sampling", "vb"):
        raise ValueError("Method must be one of {'sampling','vb'}")
    if method == "vb" and bayes_factor:
        raise ValueError("bayes_factor needs method='sampling' (bridge sampling works on posterior draws)")
    sd = prepare(y, x, type=type, lower_asymptotes=lower_asymptotes, heteroscedastic=heteroscedastic, robust=robust,
                 rho=rho, pcprior=pcprior, pcprior_hypers=pcprior_hypers, lambda_=lambda_, nu=nu,
                 bayes_factor=bayes_factor, method=method, native=True)
    control = _control_defaults(type, robust, control)
    kw = _sampler_kwargs(control, chains, iter, warmup, thin, seed)
    meta = dict(y=np.asarray(y), x=np.asarray(x),
                drug_names=list(drug_names) if drug_names is not None else ["Drug1", "Drug2"],
                experiment_ID=experiment_ID if experiment_ID is not None else "Experiment",
                units=list(units) if units is not None else ["conc.", "conc."])
    options = dict(type=type, lower_asymptotes=lower_asymptotes, method=method, bayes_factor=bayes_factor,
                   heteroscedastic=heteroscedastic, robust=robust, pcprior=pcprior, lambda_=lambda_, nu=nu,
                   pcprior_hypers=pcprior_hypers)
    b = _lib.Batch([sd], model="gp_grid", device=device)
    try:
        if method == "vb":
            res = b.advi(want_draws=True, want_cpo=True, seed=seed, **(vb or {}))
            st = int(res["info"][0, 0])
            if st != 0:
                raise RuntimeError("variational inference failed: " + _VB_ERRORS.get(st, f"status {st}"))
        else:
            from .bridge import DEVICE_BRIDGE_MAX_D
            dev_bridge = bool(bayes_factor) and sd.dim_gp_grid <= DEVICE_BRIDGE_MAX_D
            res = b.sample(want_draws=True, want_udraws=bool(bayes_factor) and not dev_bridge, want_bridge=dev_bridge,
                           want_surfaces=True, want_inv_metric=True, want_inits=True, **kw)
            blocks = b.param_blocks(0)
    finally:
        b.close()
    if method == "sampling":
        lp = res["sampler_params"][0][:, :, 6].reshape(-1)
        surfaces_k, lpml = res["surfaces"][0], res["lpml"][0]
    else:
        lp = np.zeros(res["draws"][0].shape[0] * res["draws"][0].shape[1])      # rstan::vb writes lp__ = 0
        surfaces_k, lpml = _host_surfaces(sd, res["draws"][0], robust)
    posterior = _posterior_from_draws(sd, res["draws"][0], lp)
    obj = _assemble_object(sd, meta, posterior, surfaces_k, lpml,
                           res["sampler_params"][0] if method == "sampling" else None,
                           res["chain_stats"][0] if method == "sampling" else None,
                           res["qdraws"][0] if method == "sampling" else None, options, kw.get("max_treedepth", 10))
    if method == "sampling":
        from .stanfit import StanFit
        obj.update(sampler_params=res["sampler_params"][0], chain_stats=res["chain_stats"][0], n_grad=res["n_grad"])
        # object$stanfit (R/bayesynergy.R:317): extract / summary / get_sampler_params / get_divergent_iterations ...
        D = len(param_names(sd))
        obj["stanfit"] = StanFit(output_names(sd), blocks, res["draws"][0][:, :, :len(output_names(sd))],
                                 res["sampler_params"][0], res["chain_stats"][0], res["inv_metric"][0][:, :D],
                                 res["init_u"][0][:, :D], res["kernel_ms"], warmup=kw["warmup"], args=kw)
    else:
        obj.update(vb_info=dict(zip(_lib.VI_NAMES, res["info"][0])), vb_mean=res["mean"][0, :len(param_names(sd))],
                   n_grad=res["n_evals"])
    if bayes_factor:
        from .bridge import bayes_factor as _bf, bayes_factor_from_logml
        if dev_bridge:
            obj["bayesfactor"] = float(bayes_factor_from_logml([sd], res["bridge_logml"], kw, device=device)[0])
            obj["logml"] = float(res["bridge_logml"][0])
        else:
            obj["bayesfactor"] = _bf(sd, res["udraws"][0], kw, device=device)
    return obj


def _host_surfaces(sd, draws, robust):
    """Surfaces / LPML of the ADVI output draws (1000 draws of one experiment; the sampler path reduces them on the
    device): R/bayesynergy.R:244-283."""
    post = _posterior_from_draws(sd, draws, np.zeros(draws.shape[0] * draws.shape[1]))
    n_save, n1, n2 = post["p01"].shape[0], sd.n1, sd.n2
    f = np.empty((n_save, n2 + 1, n1 + 1))
    f[:, 0, 0] = 1.0
    f[:, 1:, 0] = post["p02"]
    f[:, 0, 1:] = post["p01"]
    f[:, 1:, 1:] = post["p0"] + post["Delta"]
    p0 = f.copy()
    p0[:, 1:, 1:] = post["p0"]
    Delta = np.zeros_like(f)
    Delta[:, 1:, 1:] = post["Delta"]
    surf = np.stack([(np.median(f, axis=0) if robust else f.mean(axis=0)).T.reshape(-1),
                     np.median(p0, axis=0).T.reshape(-1), np.median(Delta, axis=0).T.reshape(-1)])
    lpml = float(-np.sum(np.log(post["CPO"].sum(axis=0) / n_save)))
    return surf, lpml


# ------------------------------------------------------------------------------------------------------
def shard_slices(n_items, world_size):
    """Contiguous, balanced split of the combination list over ranks: rank r gets
    items [lo, hi).  All chains of a combination stay on one GPU."""
    base, rem = divmod(n_items, world_size)
    out, lo = [], 0
    for r in range(world_size):
        hi = lo + base + (1 if r < rem else 0)
        out.append((lo, hi))
        lo = hi
    return out


def _screen_rows(experiments, offset, res, return_codes):
    rows = []
    S = np.where(res["summary"] == 0.0, np.nan, res["summary"])     # failed fits have no draws: scores are NaN, not 0/0
    for k, e in enumerate(experiments):
        dn = e.get("drug_names", ["DrugA", "DrugB"])
        m = {q: S[k, Q.index(q), 0] for q in Q}
        sdv = {q: S[k, Q.index(q), 1] for q in Q}
        rows.append({
            "listID": offset + k + 1, "Experiment ID": e.get("experiment_ID", "Experiment"),
            "Drug A": dn[0], "Drug B": dn[1],
            "EC50 (Drug A)": m["ec50_1"], "EC50 (Drug B)": m["ec50_2"], "DSS (Drug A)": m["dss_1"],
            "DSS (Drug B)": m["dss_2"], "Synergy (mean)": m["VUS_syn"], "Synergy (sd)": sdv["VUS_syn"],
            "Antagonism (mean)": m["VUS_ant"], "Antagonism (sd)": sdv["VUS_ant"], "Efficacy (mean)": m["rVUS_f"],
            "Efficacy (sd)": sdv["rVUS_f"], "Non-interaction Efficacy (mean)": m["rVUS_p0"],
            "Non-interaction Efficacy (sd)": sdv["rVUS_p0"], "Interaction (mean)": m["VUS_Delta"],
            "Interaction (sd)": sdv["VUS_Delta"], "Synergy Score": m["VUS_syn"] / sdv["VUS_syn"],
            "Antagonism Score": m["VUS_ant"] / sdv["VUS_ant"], "s": m["s"],
            "returnCode": int(return_codes[k]), "divergentTransitions": float(res["chain_stats"][k, :, 1].sum()),
        })
    return rows


_PREPARE_KEYS = {"type", "lower_asymptotes", "heteroscedastic", "robust", "rho", "pcprior", "pcprior_hypers", "lambda_",
                 "nu", "bayes_factor"}


def _check_prepare_params(params):
    """Validate bayesynergy_params ONCE, before the per-experiment loop: an unknown key is the caller's error (a
    TypeError), not a failed experiment.  `lambda` is accepted as the reference spells it (R/bayesynergy.R:57)."""
    if "lambda" in params:
        params["lambda_"] = params.pop("lambda")
    for k in ("units", "drug_names", "experiment_ID"):    # presentation-only arguments of bayesynergy()
        params.pop(k, None)
    unknown = sorted(set(params) - _PREPARE_KEYS)
    if unknown:
        raise TypeError(f"unknown bayesynergy_params: {unknown}")


def _fit_batch(sds, kw, device, max_treedepth, want_samples=False, want_udraws=False, want_bridge=False):
    """One batched launch over a list of prepared experiments; returns (res, returnCode[n]).

    returnCode follows bayesynergy(): 1 where rstan would have warned (divergences, tree depth, R-hat, bulk / tail ESS)
    or where the residual rule fires (R/bayesynergy.R:305-314, on the device-accumulated posterior mean surface of f;
    the reference uses the median surface when robust = TRUE -- here only when the draws are kept anyway), 2 where the
    chains could not be initialised."""
    b = _lib.Batch(sds, model="gp_grid", device=device)
    try:
        res = b.sample(want_qdraws=want_samples, want_sampler_params=want_samples, want_f_mean=True,
                       want_draws=want_samples, want_surfaces=want_samples, want_udraws=want_udraws,
                       want_bridge=want_bridge, **kw)
        n_draws = res["n_save"] * res["chain_stats"].shape[1]
        if n_draws <= 8192:         # rank-normalised bulk ESS / R-hat and tail ESS, as rstan >= 2.21 warns on
            ess_syn, rhat_syn, tail_syn = b.bulk_ess("VUS_syn")
            ess_ant, rhat_ant, tail_ant = b.bulk_ess("VUS_ant")
            tail = np.fmin(tail_syn, tail_ant)
        else:
            ess_syn, rhat_syn = b.ess("VUS_syn")
            ess_ant, rhat_ant = b.ess("VUS_ant")
            tail = np.full(len(sds), np.inf)
    finally:
        b.close()
    cs = res["chain_stats"]
    n_chains = cs.shape[1]
    rc = np.zeros(len(sds), dtype=np.int64)
    warn = (cs[:, :, 1].sum(axis=1) > 0) | (cs[:, :, 2].sum(axis=1) > 0)
    rhat = np.fmax(rhat_syn, rhat_ant)
    ess = np.fmin(ess_syn, ess_ant)
    warn |= (rhat > 1.05) | (ess < 100 * n_chains) | (tail < 100 * n_chains)
    s_mean = res["summary"][:, Q.index("s"), 0]
    for k, sd in enumerate(sds):
        ncf = (sd.n1 + 1) * (sd.n2 + 1)
        f_cells = res["surfaces"][k, 0, :ncf] if want_samples else res["f_mean"][k, :ncf]
        if _outlier_message(sd, f_cells, s_mean[k], sd.lambda_):
            warn[k] = True
    rc[warn] = 1
    rc[cs[:, :, 6].sum(axis=1) > 0] = 2
    res["ess"] = ess
    res["rhat"] = rhat
    return res, rc


def _screen_sample_object(sd, e, res, i, params, kw):
    """The `fit` object synergyscreen(return_samples = TRUE) keeps per experiment (R/synergyscreen.R:264, 409-419):
    the same fields bayesynergy() returns, assembled from experiment i of a batched launch."""
    lp = res["sampler_params"][i][:, :, 6].reshape(-1)
    posterior = _posterior_from_draws(sd, res["draws"][i], lp)
    options = dict(type=params.get("type", 3), lower_asymptotes=params.get("lower_asymptotes", True), method="sampling",
                   bayes_factor=bool(params.get("bayes_factor", False)), heteroscedastic=params.get("heteroscedastic", True),
                   robust=params.get("robust", False), pcprior=params.get("pcprior", False),
                   lambda_=params.get("lambda_", 0.005), nu=params.get("nu", 1.5),
                   pcprior_hypers=params.get("pcprior_hypers", (1, 0.1, 1, 0.2)))
    meta = dict(y=np.asarray(e["y"]), x=np.asarray(e["x"]), drug_names=e.get("drug_names", ["DrugA", "DrugB"]),
                experiment_ID=e.get("experiment_ID", "Experiment"), units=e.get("units", ["conc.", "conc."]))
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")         # the screen reports through returnCode (suppressWarnings in the reference)
        obj = _assemble_object(sd, meta, posterior, res["surfaces"][i], res["lpml"][i], res["sampler_params"][i],
                               res["chain_stats"][i], res["qdraws"][i], options, kw.get("max_treedepth", 10))
    return obj


def _fit_batch_vb(sds, vbkw, seed, device):
    """One batched ADVI launch; returnCode 2 where rstan::vb would have stopped with an error."""
    b = _lib.Batch(sds, model="gp_grid", device=device)
    try:
        res = b.advi(want_qdraws=False, seed=seed, **vbkw)
    finally:
        b.close()
    rc = np.where(res["info"][:, 0] != 0, 2, 0).astype(np.int64)
    res["chain_stats"] = np.zeros((len(sds), 1, 8))
    return res, rc


def synergyscreen(experiments, return_samples=False, max_retries=3, bayesynergy_params=None, device=None,
                  rank=0, world_size=1):
    """Fit a list of experiments (dicts with y, x and optional drug_names / experiment_ID) in one batched
    device launch (R/synergyscreen.R:58-463).  Returns dict(screenSummary=[rows], failed=[experiments]).

    With world_size > 1 this rank fits its contiguous shard only (shard_slices); gather the rows with
    `gather_screen` (a final all_gather, the only collective of the path).
    method="vb" (bayesynergy_params) runs the reference's default, full-rank ADVI, with its retry rule: refit
    whatever failed or ended with s > 1, up to max_retries more times (R/synergyscreen.R:159-167); `vb` holds
    rstan::vb arguments.  Differences from the reference, on purpose: the default method is "sampling" (with the
    whole screen in one launch NUTS no longer needs the ADVI shortcut, R/synergyscreen.R:91-93) and
    save_raw / save_plots do not exist.
    """
    if isinstance(experiments, dict) and "failed" in experiments:     # rerun mode (R/synergyscreen.R:97-103)
        old = experiments.get("screenSummary", [])
        out = synergyscreen(experiments["failed"], return_samples, max_retries, bayesynergy_params, device, rank,
                            world_size)
        out["screenSummary"] = list(old) + out["screenSummary"]
        return out
    if device is None:      # one process per GPU: under torchrun every rank lands on its own device
        import os
        device = int(os.environ.get("LOCAL_RANK", "0")) if world_size > 1 else 0
    params = dict(bayesynergy_params or {})
    method = params.pop("method", "sampling")
    if method not in ("sampling", "vb"):
        raise ValueError("Method must be one of {'sampling','vb'}")
    vbkw = dict(params.pop("vb", None) or {})
    want_bf = bool(params.get("bayes_factor", False))
    if want_bf and method != "sampling":
        raise ValueError("Computing the Bayes factor requires method = 'sampling')")
    if return_samples and method != "sampling":
        raise NotImplementedError("return_samples=TRUE with method='vb': call bayesynergy(method='vb') per experiment")
    control = params.pop("control", None)
    chains, iter_, warmup = params.pop("chains", 4), params.pop("iter", 2000), params.pop("warmup", None)
    thin, seed = params.pop("thin", 1), params.pop("seed", 1)
    type_, robust = params.get("type", 3), params.get("robust", False)
    lo, hi = shard_slices(len(experiments), world_size)[rank]
    mine = experiments[lo:hi]
    sds, ok_idx, rows_failed, fail_reason = [], [], [], {}
    _lib.load()                                 # a missing / unbuilt libbsyn.so is not a per-experiment failure
    _check_prepare_params(params)
    for k, e in enumerate(mine):
        try:
            sds.append(prepare(e["y"], e["x"], native=True, **params))
            ok_idx.append(k)
        except (ValueError, NotImplementedError) as err:   # the mirrored stop() cases of bayesynergy(): tryCatch(...
            rows_failed.append(k)                           # error = returnCode 2), R/synergyscreen.R:143-146
            fail_reason[k] = f"{type(err).__name__}: {err}"

    rows = {}
    if sds and method == "vb":
        res, rc = _fit_batch_vb(sds, vbkw, seed, device)
        for r, k in zip(_screen_rows([mine[k] for k in ok_idx], 0, res, rc), ok_idx):
            r["listID"] = lo + k + 1
            rows[k] = r
        bad = lambda j: rows[ok_idx[j]]["returnCode"] == 2 or not (rows[ok_idx[j]]["s"] <= 1)   # noqa: E731
        retry = [j for j in range(len(sds)) if bad(j)]
        tries = 0
        while retry and tries <= max_retries:
            res2, rc2 = _fit_batch_vb([sds[j] for j in retry], vbkw, seed + 1000 * (tries + 1), device)
            new_rows = _screen_rows([mine[ok_idx[j]] for j in retry], 0, res2, rc2)
            for r, j in zip(new_rows, retry):
                r["listID"] = lo + ok_idx[j] + 1
                rows[ok_idx[j]] = r
            retry = [j for j in retry if bad(j)]
            tries += 1
    elif sds:
        # return_samples keeps every write_array draw on the host (as the reference keeps every stanfit): the launches are
        # then chunked so that a chunk's draws fit comfortably in host memory
        chunk = max(1, int(2e9 // (8 * chains * max(1, iter_ // 2) * 1024))) if return_samples else len(sds)
        samples, udraws_gp, kw_of, lml_gp = {}, {}, {}, {}
        from .bridge import DEVICE_BRIDGE_MAX_D
        dev_bridge = want_bf and max(sd.dim_gp_grid for sd in sds) <= DEVICE_BRIDGE_MAX_D

        def fit(js, kwj):
            """fit the experiments js (indices into sds) and store their rows / samples"""
            codes = {}
            for c0 in range(0, len(js), chunk):
                part = js[c0:c0 + chunk]
                res_, rc_ = _fit_batch([sds[j] for j in part], kwj, device, kwj.get("max_treedepth", 10),
                                       want_samples=return_samples, want_udraws=want_bf and not dev_bridge,
                                       want_bridge=dev_bridge)
                new_rows = _screen_rows([mine[ok_idx[j]] for j in part], 0, res_, rc_)
                for i, (r, j) in enumerate(zip(new_rows, part)):
                    r["listID"] = lo + ok_idx[j] + 1
                    rows[ok_idx[j]] = r
                    codes[j] = rc_[i]
                    kw_of[j] = kwj
                    if dev_bridge:
                        lml_gp[j] = res_["bridge_logml"][i]
                    elif want_bf:
                        udraws_gp[j] = res_["udraws"][i]
                    if return_samples:
                        samples[j] = _screen_sample_object(sds[j], mine[ok_idx[j]], res_, i, params, kwj)
            return codes

        ctl = _control_defaults(type_, robust, control)
        kw = _sampler_kwargs(ctl, chains, iter_, warmup, thin, seed)
        codes = fit(list(range(len(sds))), kw)
        # retry the experiments with warnings at adapt_delta = 0.99 (R/synergyscreen.R:149-171), as ONE smaller launch
        retry = [j for j in range(len(sds)) if codes[j] != 0]
        tries = 0
        while retry and tries <= max_retries:   # the reference's loop body runs max_retries + 1 times (quirk 9)
            ctl2 = dict(ctl, adapt_delta=0.99)
            kw2 = _sampler_kwargs(ctl2, chains, iter_, warmup, thin, seed + 1000 * (tries + 1))
            codes2 = fit(retry, kw2)
            retry = [j for j in retry if codes2[j] != 0]
            tries += 1
        if want_bf:
            # cfg4: "two models per combination" -- one nointeraction launch over every fitted experiment and one batched
            # log_prob launch per model for all bridge points; with <= 160 parameters the whole estimator runs on the device
            # and only the log marginal likelihoods come back (bsyn_sample's bridge_logml sink), else bridge.bayes_factor_many
            from .bridge import bayes_factor_many, bayes_factor_from_logml
            good = [j for j in range(len(sds)) if rows[ok_idx[j]]["returnCode"] != 2]
            for kwj in {id(kw_of[j]): kw_of[j] for j in good}.values():
                js = [j for j in good if kw_of[j] is kwj]
                if dev_bridge:
                    bfs = bayes_factor_from_logml([sds[j] for j in js], [lml_gp[j] for j in js], kwj, device=device)
                else:
                    bfs = bayes_factor_many([sds[j] for j in js], [udraws_gp[j] for j in js], kwj, device=device, seed=seed)
                for j, v in zip(js, bfs):
                    rows[ok_idx[j]]["Bayes Factor"] = float(v)
                    if return_samples and j in samples:
                        samples[j]["bayesfactor"] = float(v)
            for r in rows.values():
                r.setdefault("Bayes Factor", float("nan"))
    summary = [rows[k] for k in sorted(rows)]
    # s > 1 => returnCode 2; drop code-2 rows; collect failed (R/synergyscreen.R:424-451)
    for r in summary:
        if r["s"] > 1:
            r["returnCode"] = 2
    kept = [r for r in summary if r["returnCode"] != 2]
    failed_ids = sorted(set(rows_failed) | {r["listID"] - lo - 1 for r in summary if r["returnCode"] == 2})
    out = dict(screenSummary=kept)
    if return_samples and method == "sampling":
        out["screenSamples"] = [samples[j] for j in sorted(samples) if rows[ok_idx[j]]["returnCode"] != 2]
    if failed_ids:
        out["failed"] = [mine[k] for k in failed_ids]
        out["failed_reason"] = {lo + k + 1: fail_reason.get(k, "returnCode 2") for k in failed_ids}
    n_warn = sum(1 for r in kept if r["returnCode"] > 0)
    if kept and n_warn:
        warnings.warn(f"{100.0 * n_warn / len(kept):.2f}% of experiments returned a warning or error -- check these.")
    return out


def screen_to_ess(sds, target=400.0, batch=None, device=0, chains=4, warmup=1000, block=150, max_rounds=6, seed=1,
                  adapt_delta=0.9, retry="refit", retry_adapt_delta=0.99, retry_rounds=6, retry_block=300, retry_warmup=150,
                  want_summary=True):
    """north_star's target run: sample every prepared experiment until min(bulk-ESS(VUS_syn), bulk-ESS(VUS_ant)) >=
    `target` (rank-normalised, on the device; bsyn_sample_to_ess).

    Phase 1: warm-up + `block` draws per chain for everything, then extension launches over the experiments still below
    the target -- chains resume from their adapted state -- up to `max_rounds` launches.
    Phase 2, the reference's remedy for what is left (R/synergyscreen.R:149-171: adapt_delta = 0.99), `retry_rounds`
    launches at most, in one of two forms:
      retry="refit"  (what the reference does) the experiments still below the target are refitted FROM SCRATCH in one
                     small batch at `retry_adapt_delta` (new initial values, full warm-up, blocks of `retry_block`); the
                     fit with the larger ESS is kept.  Slower (the small batch is latency-bound), reaches more of them:
                     a fresh start also re-draws which mode each chain of a multimodal posterior falls into.
      retry="resume" inside the same bsyn_sample_to_ess call: their chains resume, re-adapt the STEP SIZE to
                     `retry_adapt_delta` for `retry_warmup` iterations (metric kept) and append blocks; earlier draws stay
                     pooled with the new ones.  Cheaper; chains stuck in different modes stay there (R-hat > 1.05).
      retry=None     no phase 2.
    `batch`: an existing _lib.Batch of `sds` (kept open); else one is created and closed here.
    Returns dict(ess, draws_per_chain, summary, chain_stats, n_grad, kernel_ms, phase1_ms, phase2_ms, rounds, n_retry,
    reached)."""
    if retry not in ("refit", "resume", None):
        raise ValueError('retry must be "refit", "resume" or None')
    own = batch is None
    b = _lib.Batch(sds, device=device) if own else batch
    try:
        r = b.sample_to_ess(target=target, block=block, max_rounds=max_rounds, chains=chains, warmup=warmup, seed=seed,
                            adapt_delta=adapt_delta, want_summary=want_summary,
                            retry_rounds=retry_rounds if retry == "resume" else 0, retry_adapt_delta=retry_adapt_delta,
                            retry_warmup=retry_warmup)
    finally:
        if own:
            b.close()
    ess, nd = r["ess"].copy(), r["draws_per_chain"].copy()
    out = dict(ess=ess, draws_per_chain=nd, n_grad=r["n_grad"], kernel_ms=r["kernel_ms"], rounds=r["rounds"],
               n_retry=int(r["n_retried"]), summary=r.get("summary"), chain_stats=r.get("chain_stats"),
               phase1_ms=r["phase1_ms"], phase2_ms=r["kernel_ms"] - r["phase1_ms"], retry=retry)
    left = np.where(~(ess >= target))[0]
    if retry == "refit" and len(left) and retry_rounds > 0:
        b2 = _lib.Batch([sds[k] for k in left], device=device)
        try:
            r2 = b2.sample_to_ess(target=target, block=retry_block, max_rounds=retry_rounds, retry_rounds=0, chains=chains,
                                  warmup=warmup, seed=seed + 7777, adapt_delta=retry_adapt_delta, want_summary=want_summary)
        finally:
            b2.close()
        better = np.nan_to_num(r2["ess"], nan=-1.0) > np.nan_to_num(ess[left], nan=-1.0)
        for i, k in enumerate(left):
            if better[i]:
                ess[k], nd[k] = r2["ess"][i], r2["draws_per_chain"][i]
                if want_summary:
                    out["summary"][k], out["chain_stats"][k] = r2["summary"][i], r2["chain_stats"][i]
        out["n_grad"] += r2["n_grad"]
        out["kernel_ms"] += r2["kernel_ms"]
        out["phase2_ms"] = r2["kernel_ms"]
        out["rounds"] += r2["rounds"]
        out["n_retry"] = int(len(left))
    out["reached"] = int((ess >= target).sum())
    return out


def gather_screen(local, world_size):
    """The path's only collective: gather every rank's screenSummary rows (torch.distributed, any backend)."""
    if world_size == 1:
        return local
    import torch.distributed as dist
    parts = [None] * world_size
    dist.all_gather_object(parts, local)
    out = dict(screenSummary=[r for p in parts for r in p["screenSummary"]])
    failed = [e for p in parts for e in p.get("failed", [])]
    if failed:
        out["failed"] = failed
    return out
