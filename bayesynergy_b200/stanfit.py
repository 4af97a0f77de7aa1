"""A `stanfit`-shaped view of one fitted experiment: what `bayesynergy()` keeps in `object$stanfit`
(R/bayesynergy.R:317) and what its callers read through rstan -- `extract`, `summary`, `get_sampler_params`,
`get_divergent_iterations`, `get_adaptation_info`, `get_elapsed_time`, `get_inits`.

It is assembled from the sinks of ONE `bsyn_sample` call (all chains of the experiment in one launch); the C++
counterpart that the Rcpp module wraps is `include/bsyn_stanfit.hpp`.  Flat names and their order are the library's
(`bsyn_output_name`, = src/stanExports_gp_grid.h:2256-2402), `lp__` last, as in rstan.
"""
from __future__ import annotations

import numpy as np

from .bridge import _ess_columns

SAMPLER_PARAMS = ["accept_stat__", "stepsize__", "treedepth__", "n_leapfrog__", "divergent__", "energy__"]


class StanFit:
    def __init__(self, flat_names, blocks, draws, sampler_params, chain_stats, inv_metric=None, inits=None,
                 kernel_ms=0.0, warmup=0, args=None, model_name="gp_grid"):
        """draws: [chains, n_save, n_out] write_array rows (post-warmup unless save_warmup); sampler_params:
        [chains, n_save, 7] (the 7th column is lp__); blocks: [(name, dims, flat_offset)] from Batch.param_blocks."""
        self.model_name = model_name
        self.flat_names = list(flat_names) + ["lp__"]
        self.blocks = list(blocks)
        self.chains, self.n_save = draws.shape[0], draws.shape[1]
        self._draws = np.concatenate([draws[:, :, :len(flat_names)], sampler_params[:, :, 6:7]], axis=2)
        self._sp = sampler_params[:, :, :6]
        self._cs = chain_stats
        self._inv_metric = inv_metric
        self._inits = inits
        self._kernel_ms = kernel_ms
        self.warmup = warmup
        self.args = dict(args or {})

    # -- rstan::extract ----------------------------------------------------------------------------------------------
    def extract(self, pars=None, permuted=True):
        """permuted=True: dict name -> array [chains * n_save, *dims] (chains pooled; rstan also permutes the draws, which
        no caller of bayesynergy depends on).  permuted=False: array [n_save, chains, n_flat] + the flat names."""
        if not permuted:
            return self._draws.transpose(1, 0, 2).copy(), list(self.flat_names)
        out = {}
        flat = self._draws.reshape(-1, self._draws.shape[2])
        for name, dims, off in self.blocks:
            if pars is not None and name not in pars:
                continue
            n = int(np.prod(dims)) if dims else 1
            if n == 0:
                out[name] = np.empty((flat.shape[0], 0))
                continue
            col = len(self.flat_names) - 1 if name == "lp__" else off
            v = flat[:, col:col + n]
            if len(dims) == 2:                      # column-major in the flat vector
                v = v.reshape(flat.shape[0], dims[1], dims[0]).transpose(0, 2, 1)
            elif not dims:
                v = v[:, 0]
            out[name] = v
        return out

    # -- rstan::summary ----------------------------------------------------------------------------------------------
    def summary(self, pars=None, probs=(0.025, 0.25, 0.5, 0.75, 0.975)):
        """dict(summary=[n_flat, 4 + len(probs) + 2], rownames, colnames): mean, se_mean, sd, quantiles, n_eff, Rhat per
        flat parameter -- n_eff / Rhat by the split-chain estimator rstan::summary uses (Geyer's initial sequences)."""
        idx = [i for i, nm in enumerate(self.flat_names)
               if pars is None or nm.split("[")[0] in pars]
        x = self._draws[:, :, idx]                                   # [chains, n, k]
        flat = x.reshape(-1, len(idx))
        mean, sd = flat.mean(axis=0), flat.std(axis=0, ddof=1)
        q = np.quantile(flat, probs, axis=0).T
        n = x.shape[1] // 2
        second = x.shape[1] - n
        with np.errstate(all="ignore"):
            ess = _ess_columns(x)
            sp = np.concatenate([np.stack([x[c, :n], x[c, second:second + n]]) for c in range(x.shape[0])])
            w = sp.var(axis=1, ddof=1).mean(axis=0)
            var_plus = w * (n - 1.0) / n + sp.mean(axis=1).var(axis=0, ddof=1)
            rhat = np.sqrt(var_plus / w)
            se = sd / np.sqrt(ess)
        table = np.column_stack([mean, se, sd, q, ess, rhat])
        cols = ["mean", "se_mean", "sd"] + [f"{100 * p:g}%" for p in probs] + ["n_eff", "Rhat"]
        return dict(summary=table, rownames=[self.flat_names[i] for i in idx], colnames=cols)

    # -- sampler diagnostics ---------------------------------------------------------------------------------------------
    def get_sampler_params(self, inc_warmup=False):
        """list (one per chain) of dict name -> array, as rstan::get_sampler_params(fit, inc_warmup = FALSE)."""
        return [{nm: self._sp[c, :, k].copy() for k, nm in enumerate(SAMPLER_PARAMS)} for c in range(self.chains)]

    def get_divergent_iterations(self):
        return (self._sp[:, :, 4] > 0).reshape(-1)

    def get_num_divergent(self):
        return int((self._sp[:, :, 4] > 0).sum())

    def get_max_treedepth_iterations(self, max_treedepth=10):
        return (self._sp[:, :, 2] >= max_treedepth).reshape(-1)

    def get_adaptation_info(self):
        """One string per chain in Stan's text format: step size and the (diagonal of the) inverse metric."""
        out = []
        for c in range(self.chains):
            s = f"# Adaptation terminated\n# Step size = {self._cs[c, 0]:g}\n# Diagonal elements of inverse mass matrix:\n# "
            if self._inv_metric is not None:
                s += ", ".join(f"{v:g}" for v in self._inv_metric[c])
            out.append(s + "\n")
        return out

    def get_elapsed_time(self):
        """[chains, 2] seconds (warmup, sample): the device time of the launch, shared by its chains, split by each
        chain's leapfrog counts in the two phases."""
        tot = 1e-3 * self._kernel_ms
        lw = self._cs[:, 3] - self._cs[:, 4]
        ls = self._cs[:, 4]
        with np.errstate(all="ignore"):
            fw = np.where(lw + ls > 0, lw / (lw + ls), 0.0)
        return np.column_stack([tot * fw, tot * (1.0 - fw)])

    def get_inits(self):
        """Initial values of every chain on the unconstrained scale ([chains, D]); constrain with Batch.write_array."""
        return self._inits

    def __repr__(self):
        return (f"<StanFit {self.model_name}: {self.chains} chains x {self.n_save} draws, {len(self.flat_names)} flat "
                f"parameters, {self.get_num_divergent()} divergent>")
