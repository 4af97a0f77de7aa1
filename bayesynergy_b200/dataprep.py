"""Host-side data preparation: (y, x) -> the `stan_data` problem descriptor.

Mirrors the reference's R code that runs before the sampler is entered
(R/bayesynergy.R:63-104 validation, :129-154 dose grid / ii_obs / nrep / nmissing,
:167-200 stan_data + kernel flags).  The descriptor produced here is exactly the
set of fields the generated model constructors read
(src/stanExports_gp_grid.h:510-636, src/stanExports_nointeraction.h:372-412).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import numpy as np

# kernel ids as in inst/stan/gp_grid.stan:112-141
KERNEL_RBF = 1
KERNEL_MATERN = 2
KERNEL_RQ = 3


@dataclass
class StanData:
    """One experiment's model inputs (field names = the Stan data block, gp_grid.stan:10-29)."""
    n1: int
    n2: int
    nmissing: int
    nrep: int
    y: np.ndarray            # float64 [N]
    ii_obs: np.ndarray       # int32 [N], 1-based into the column-major (n2+1) x (n1+1) grid f
    x1: np.ndarray           # float64 [n1] log10 doses of drug 1
    x2: np.ndarray           # float64 [n2]
    est_la: int = 1
    robust: int = 0
    pcprior: int = 0
    heteroscedastic: int = 1
    pcprior_hypers: np.ndarray = field(default_factory=lambda: np.array([1.0, 0.1, 1.0, 0.2]))
    rho: float = 0.9
    lambda_: float = 0.005
    kernel: int = KERNEL_MATERN
    nu_matern: int = 2
    est_alpha: int = 0

    @property
    def N(self) -> int:
        return int(self.y.shape[0])

    @property
    def dim_gp_grid(self) -> int:
        # src/stanExports_gp_grid.h:755-795
        return 13 + 2 * self.est_la + self.est_alpha + self.n1 * self.n2

    @property
    def dim_nointeraction(self) -> int:
        # src/stanExports_nointeraction.h (parameters block): 9 + 2*est_la
        return 9 + 2 * self.est_la

    def check(self):
        N_expected = (self.n1 + self.n2 + self.n1 * self.n2 + 1) * self.nrep - self.nmissing
        if N_expected != self.N:
            raise ValueError(
                f"size mismatch: (n1+n2+n1*n2+1)*nrep-nmissing = {N_expected} but len(y) = {self.N}")
        if self.ii_obs.shape[0] != self.N:
            raise ValueError("ii_obs and y must have the same length")
        ncell = (self.n1 + 1) * (self.n2 + 1)
        if self.N and (self.ii_obs.min() < 1 or self.ii_obs.max() > ncell):
            raise ValueError("ii_obs out of range")


def prepare(y, x, type=3, lower_asymptotes=True, heteroscedastic=True, robust=False, rho=0.90,
            pcprior=False, pcprior_hypers=(1, 0.1, 1, 0.2), lambda_=0.005, nu=1.5,
            bayes_factor=False, method="sampling", native=False) -> StanData:
    """R/bayesynergy.R:54-200 up to (not including) the sampler call.

    native=True does the grid / ii_obs / missing-value work (R/bayesynergy.R:129-154) in libbsyn's bsyn_prepare
    (what api.bayesynergy / api.synergyscreen use); native=False is the numpy statement of the same steps that the
    CPU-only tests use.

    `y`: [n] vector or [n, nrep] matrix (NaN = missing).  `x`: [n, 2] concentrations
    (zero = monotherapy).  Error messages follow the reference's `stop()` texts.
    """
    if method not in ("sampling", "vb"):
        raise ValueError("Method must be one of {'sampling','vb'}")
    if type not in (1, 2, 3, 4):
        raise ValueError("Argument 'type' must be one of {1,2,3,4}!")
    # NB: the reference wires this check to type == 2 (R/bayesynergy.R:70, a latent bug: the Matern
    # kernel is type 3).  We check it where it matters so an invalid nu is an error, not a Stan failure.
    if type == 3 and nu not in (0.5, 1.5, 2.5):
        raise ValueError("Argument 'nu' must be one of {1/2,3/2,5/2}!")
    if type == 1:
        raise NotImplementedError("type = 1 (splines model) is outside the gp_grid/nointeraction hot path")
    if type != 3 and pcprior:
        raise ValueError("PC prior only available for Matern covariance")
    if robust:
        lo = 2 * 0.5 * (1 + math.erf(1 / math.sqrt(2))) - 1
        if rho < lo or rho > 1:
            raise ValueError("rho must be between 0.6827 and 1")
    x = np.asarray(x, dtype=np.float64)
    y = np.asarray(y, dtype=np.float64)
    if x.ndim != 2 or x.shape[1] != 2:
        raise ValueError("Dimension mismatch! Argument 'x' should be a matrix with two columns "
                         "containing drug concentrations.")
    if y.ndim == 1:
        y = y[:, None]
    if y.shape[0] != x.shape[0]:
        raise ValueError("Dimension mismatch! Arguments 'y' and 'x' should be equally sized.")
    if x[:, 0].min() > 0 or x[:, 1].min() > 0:
        raise ValueError("Missing monotherapy data! Make sure 'x' contains zero concentrations")
    if bayes_factor and method != "sampling":
        raise ValueError("Computing the Bayes factor requires method = 'sampling')")

    if native:
        from . import _lib
        unq1, unq2, yv, ii_obs, nrep, nmissing = _lib.prepare_native(y, x)
        n1, n2 = len(unq1), len(unq2)
        return _finish(unq1, unq2, n1, n2, yv, ii_obs, nrep, nmissing, type, nu, lower_asymptotes, robust, pcprior,
                       heteroscedastic, pcprior_hypers, rho, lambda_)

    # R/bayesynergy.R:129-137
    with np.errstate(divide="ignore"):
        unq1 = np.log10(np.unique(x[:, 0]))[1:]
        unq2 = np.log10(np.unique(x[:, 1]))[1:]
    X1 = np.concatenate([[0.0], 10.0 ** unq1])
    X2 = np.concatenate([[0.0], 10.0 ** unq2])
    n1, n2 = len(unq1), len(unq2)
    # expand.grid(X1, X2) ordered by Var1 then Var2  ->  row a*(n2+1)+b  <->  (X1[a], X2[b])
    # The reference matches rows with match() on values rounded to 15 decimals after the
    # 10^log10(x) round trip (R/bayesynergy.R:137); unique doses are never merged, so matching each
    # row to its nearest grid level is the same map without depending on decimal rounding.
    a_idx = np.abs(x[:, 0][:, None] - X1[None, :]).argmin(axis=1)
    b_idx = np.abs(x[:, 1][:, None] - X2[None, :]).argmin(axis=1)
    if (np.abs(X1[a_idx] - x[:, 0]) > 1e-9 * np.maximum(1.0, np.abs(x[:, 0]))).any() or \
            (np.abs(X2[b_idx] - x[:, 1]) > 1e-9 * np.maximum(1.0, np.abs(x[:, 1]))).any():
        raise ValueError("dose level could not be matched to the grid")
    ii = a_idx.astype(np.int64) * (n2 + 1) + b_idx.astype(np.int64) + 1

    # R/bayesynergy.R:141-154
    nrep = y.shape[1]
    ngrid = n1 + n2 + n1 * n2 + 1
    if nrep > 1:
        ii_rep = np.tile(ii, nrep)
        yv = y.T.reshape(-1)                      # as.vector of a column-major matrix
        keep = ~np.isnan(yv)
        ii_obs = ii_rep[keep]
        nmissing = ngrid * nrep - ii_obs.shape[0]
        yv = yv[keep]
    else:
        counts = np.bincount(ii)
        nrep = int(counts.max())
        yv = y[:, 0]
        keep = ~np.isnan(yv)
        ii_obs = ii[keep]
        nmissing = ngrid * nrep - ii_obs.shape[0]
        yv = yv[keep]

    return _finish(unq1, unq2, n1, n2, yv, ii_obs, nrep, nmissing, type, nu, lower_asymptotes, robust, pcprior,
                   heteroscedastic, pcprior_hypers, rho, lambda_)


def _finish(unq1, unq2, n1, n2, yv, ii_obs, nrep, nmissing, type, nu, lower_asymptotes, robust, pcprior,
            heteroscedastic, pcprior_hypers, rho, lambda_) -> StanData:
    kernel, nu_matern, est_alpha = KERNEL_MATERN, 2, 0
    if type == 2:
        kernel, nu_matern, est_alpha = KERNEL_RBF, 0, 0
    elif type == 3:
        kernel = KERNEL_MATERN
        nu_matern = {0.5: 1, 1.5: 2, 2.5: 3}[nu]
    elif type == 4:
        kernel, nu_matern, est_alpha = KERNEL_RQ, 0, 1

    sd = StanData(n1=n1, n2=n2, nmissing=int(nmissing), nrep=int(nrep),
                  y=np.ascontiguousarray(yv, dtype=np.float64),
                  ii_obs=np.ascontiguousarray(ii_obs, dtype=np.int32),
                  x1=np.ascontiguousarray(unq1), x2=np.ascontiguousarray(unq2),
                  est_la=int(bool(lower_asymptotes)), robust=int(bool(robust)),
                  pcprior=int(bool(pcprior)), heteroscedastic=int(bool(heteroscedastic)),
                  pcprior_hypers=np.asarray(pcprior_hypers, dtype=np.float64),
                  rho=float(rho), lambda_=float(lambda_), kernel=kernel, nu_matern=nu_matern,
                  est_alpha=est_alpha)
    sd.check()
    return sd


# ---------------------------------------------------------------------------------------------
# Parameter / output naming (src/stanExports_gp_grid.h:1588-1704, :2256-2549; Appendix A of SURVEY)
# ---------------------------------------------------------------------------------------------

def param_names(sd: StanData, model="gp_grid"):
    """Flat names of the unconstrained vector, in the reference's order."""
    out = []
    if sd.est_la:
        out += ["la_1[1]", "la_2[1]"]
    out += ["log10_ec50_1", "log10_ec50_2", "theta_1", "theta_2", "slope_1", "slope_2"]
    if model == "gp_grid":
        out += ["b1", "b2", "ell", "sigma_f"]
        if sd.est_alpha:
            out += ["alpha[1]"]
        out += [f"z[{i+1},{j+1}]" for j in range(sd.n1) for i in range(sd.n2)]
    out += ["s", "s2_log10_ec50_1", "s2_log10_ec50_2"]
    return out


def output_names(sd: StanData, model="gp_grid"):
    """Flat names of one write_array row (params, transformed params, generated quantities)."""
    n1, n2, N = sd.n1, sd.n2, sd.N
    out = param_names(sd, model)
    out += [f"p0[{i+1},{j+1}]" for j in range(n1) for i in range(n2)]
    out += [f"p01[{j+1}]" for j in range(n1)]
    out += [f"p02[{i+1}]" for i in range(n2)]
    if model == "gp_grid":
        out += [f"Delta[{i+1},{j+1}]" for j in range(n1) for i in range(n2)]
        out += [f"GP[{i+1},{j+1}]" for j in range(n1) for i in range(n2)]
    out += ["ec50_1", "ec50_2"]
    out += [f"CPO[{n+1}]" for n in range(N)]
    out += ["dss_1", "dss_2", "rVUS_f", "rVUS_p0"]
    if model == "gp_grid":
        out += ["VUS_Delta", "VUS_syn", "VUS_ant"]
    return out


def synthetic_experiment(combo_id: int, n=10, nrep=3, seed0=20261019):
    """cfg5 generator: one synthetic 10x10x3 combination.

    doses per drug = 0 U 10^linspace(-3, 1, n); y is ((n+1)^2, nrep); generative model =
    gp_grid with Matern 3/2, ell=1, sigma_f=1, b1=b2=1 and a per-combo interaction strength.
    """
    rng = np.random.default_rng(seed0 + combo_id)
    la = rng.beta(1.0, 1.25, size=2) * 0.5
    slope = rng.gamma(2.0, 1.0, size=2) + 0.3
    m = rng.normal(-1.0, 1.0, size=2)
    strength = rng.choice([0.0, 0.0, 0.5, 1.5])
    z = rng.normal(size=(n, n)) * strength
    lx = np.linspace(-3.0, 1.0, n)
    d = np.abs(lx[:, None] - lx[None, :])
    K = (1 + math.sqrt(3) * d) * np.exp(-math.sqrt(3) * d) + 1e-10 * np.eye(n)
    L = np.linalg.cholesky(K)
    GP = L @ z @ L.T
    p01 = la[0] + (1 - la[0]) / (1 + 10 ** (slope[0] * (lx - m[0])))
    p02 = la[1] + (1 - la[1]) / (1 + 10 ** (slope[1] * (lx - m[1])))
    p0 = p02[:, None] * p01[None, :]
    lg = np.log(p0 / (1 - p0))
    Delta = -p0 / (1 + np.exp(GP + lg)) + (1 - p0) / (1 + np.exp(-GP - lg))
    f = np.empty((n + 1, n + 1))
    f[0, 0] = 1.0
    f[0, 1:] = p01
    f[1:, 0] = p02
    f[1:, 1:] = p0 + Delta
    doses = np.concatenate([[0.0], 10.0 ** lx])
    # rows ordered with drug-1 level a outer, drug-2 level b inner (row = a*(n+1)+b)
    a_idx, b_idx = np.meshgrid(np.arange(n + 1), np.arange(n + 1), indexing="ij")
    x = np.stack([doses[a_idx.reshape(-1)], doses[b_idx.reshape(-1)]], axis=1)
    fvec = f[b_idx.reshape(-1), a_idx.reshape(-1)]
    y = fvec[:, None] + rng.normal(size=(fvec.shape[0], nrep)) * (0.1 * np.sqrt(fvec + 0.005))[:, None]
    return y, x


def synthetic_oneil_experiment(combo_id: int, seed0=20261019 + 10_000_000, mono_reps=12):
    """cfg2/cfg3 generator: one synthetic combination on the exact dose layout of the shipped raw
    experiment `ONeil_A375$failed[[1]]`: 8 monotherapy doses per drug x `mono_reps` replicates on dose levels that
    do not align with the 4 x 4 combination block (x 4 replicates), no (0, 0) well  =>  11 distinct non-zero levels
    per axis (top dose shared), N = 256, 16 of 121 interior cells observed.  Returns (y[N], x[N, 2])."""
    rng = np.random.default_rng(seed0 + combo_id)
    m1 = np.array([0.0022, 0.0065, 0.02, 0.06, 0.19, 0.55, 1.7, 5.0])
    m2 = np.array([0.0045, 0.014, 0.04, 0.12, 0.37, 1.1, 3.4, 10.0])
    c1 = np.array([0.325, 0.8, 2.0, 5.0])
    c2 = np.array([0.223, 0.775, 2.75, 10.0])
    scale = 10.0 ** rng.uniform(-0.3, 0.3, size=2)          # per-combination dose scale
    m1, c1, m2, c2 = m1 * scale[0], c1 * scale[0], m2 * scale[1], c2 * scale[1]
    la = rng.beta(1.0, 1.25, size=2) * 0.5
    slope = rng.gamma(2.0, 1.0, size=2) + 0.3
    mid = np.array([np.log10(np.median(m1)), np.log10(np.median(m2))]) + rng.normal(0, 0.5, size=2)
    strength = rng.choice([0.0, 0.0, 0.5, 1.5])
    u1 = np.unique(np.concatenate([m1, c1]))
    u2 = np.unique(np.concatenate([m2, c2]))
    lx1, lx2 = np.log10(u1), np.log10(u2)

    def chol_k(lx):
        d = np.abs(lx[:, None] - lx[None, :])
        K = (1 + math.sqrt(3) * d) * np.exp(-math.sqrt(3) * d) + 1e-10 * np.eye(len(lx))
        return np.linalg.cholesky(K)

    z = rng.normal(size=(len(u2), len(u1))) * strength
    GP = chol_k(lx2) @ z @ chol_k(lx1).T
    p01 = la[0] + (1 - la[0]) / (1 + 10 ** (slope[0] * (lx1 - mid[0])))
    p02 = la[1] + (1 - la[1]) / (1 + 10 ** (slope[1] * (lx2 - mid[1])))
    p0 = p02[:, None] * p01[None, :]
    lg = np.log(p0 / (1 - p0))
    f = p0 + (-p0 / (1 + np.exp(GP + lg)) + (1 - p0) / (1 + np.exp(-GP - lg)))
    rows, vals = [], []
    for a, d in enumerate(u1):
        if d in m1:
            for _ in range(mono_reps):
                rows.append((d, 0.0)); vals.append(p01[a])
    for b, d in enumerate(u2):
        if d in m2:
            for _ in range(mono_reps):
                rows.append((0.0, d)); vals.append(p02[b])
    for d1 in c1:
        for d2 in c2:
            a, b = int(np.where(u1 == d1)[0][0]), int(np.where(u2 == d2)[0][0])
            for _ in range(4):
                rows.append((d1, d2)); vals.append(f[b, a])
    vals = np.array(vals)
    y = vals + rng.normal(size=vals.shape[0]) * 0.1 * np.sqrt(vals + 0.005)
    return y, np.array(rows)
