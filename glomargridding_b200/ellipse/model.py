"""`EllipseModel`: drop-in for `glomar_gridding.ellipse.EllipseModel` (reference `ellipse/model.py:156-579`) on a B200.

Same constructor, attributes (`fform`, `model_type`, `supercategory`, `supercategory_params`, `n_params`,
`default_guesses`, `default_bounds`) and methods (`negative_log_likelihood`, `negative_log_likelihood_function`, `fit`)
as the reference; the likelihood and the Nelder-Mead optimisation run on the device (`gmg_ellipse_nll_f64`,
`gmg_ellipse_fit_f64`, `csrc/ellipse_fit.cu` - a restatement of scipy's simplex method that follows the same path as
`scipy.optimize.minimize(method="Nelder-Mead", bounds=...)`), and there is no CPU fallback.

New: `fit_many` fits any number of independent problems (grid cells) in ONE launch - what
`EllipseBuilder.compute_params` uses instead of the reference's per-cell Python loop.

Not carried over: the bootstrap standard errors (`estimate_SE`, reference `model.py:500-579`, joblib workers) - only
`estimate_SE=None` is accepted - and optimisers other than Nelder-Mead.
"""

from __future__ import annotations

import ctypes
import math
from collections import OrderedDict
from collections.abc import Callable
from typing import Any

import numpy as np
import torch

from .. import _device as dev
from .._lib import check, lib

KM_PER_DEG = 60.0 * 1.852            # utils.deg_to_km: 60 nautical miles per degree of latitude


def deg_to_km(deg: float) -> float:
    """Latitude difference in degrees -> kilometres (reference `utils.py:498-521`)."""
    return KM_PER_DEG * deg


MODEL_TYPE_TO_SUPERCATEGORY = {
    "ps2006_kks2011_iso": "1_param_matern", "ps2006_kks2011_ani": "2_param_matern",
    "ps2006_kks2011_ani_r": "3_param_matern", "ps2006_kks2011_iso_pd": "1_param_matern_pd",
    "ps2006_kks2011_ani_pd": "2_param_matern_pd", "ps2006_kks2011_ani_r_pd": "3_param_matern_pd",
}
FFORM_TO_MODELTYPE = {
    "anisotropic_rotated": "ps2006_kks2011_ani_r", "anisotropic": "ps2006_kks2011_ani",
    "isotropic": "ps2006_kks2011_iso", "anisotropic_rotated_pd": "ps2006_kks2011_ani_r_pd",
    "anisotropic_pd": "ps2006_kks2011_ani_pd", "isotropic_pd": "ps2006_kks2011_iso_pd",
}


def _params(names, unit):
    out = OrderedDict((n, unit if n in ("Lx", "Ly", "R") else "radians") for n in names)
    out.update([("standard_deviation", "K"), ("qc_code", "1"), ("number_of_iterations", "1")])
    return out


SUPERCATEGORY_PARAMS = {
    "3_param_matern": _params(["Lx", "Ly", "theta"], "degrees"), "2_param_matern": _params(["Lx", "Ly"], "degrees"),
    "1_param_matern": _params(["R"], "degrees"), "3_param_matern_pd": _params(["Lx", "Ly", "theta"], "km"),
    "2_param_matern_pd": _params(["Lx", "Ly"], "km"), "1_param_matern_pd": _params(["R"], "km"),
}
_TWO_PI = 2.0 * math.pi
FFORM_PARAMETERS: dict[str, dict[str, Any]] = {
    "isotropic": {"n_params": 1, "default_guesses": [7.0], "default_bounds": [(0.5, 50.0)]},
    "isotropic_pd": {"n_params": 1, "default_guesses": [deg_to_km(7.0)],
                     "default_bounds": [(deg_to_km(0.5), deg_to_km(50.0))]},
    "anisotropic": {"n_params": 2, "default_guesses": [7.0, 7.0], "default_bounds": [(0.5, 50.0), (0.5, 30.0)]},
    "anisotropic_pd": {"n_params": 2, "default_guesses": [deg_to_km(7.0), deg_to_km(7.0)],
                       "default_bounds": [(deg_to_km(0.5), deg_to_km(50.0)), (deg_to_km(0.5), deg_to_km(30.0))]},
    "anisotropic_rotated": {"n_params": 3, "default_guesses": [7.0, 7.0, 0.0],
                            "default_bounds": [(0.5, 50.0), (0.5, 30.0), (-_TWO_PI, _TWO_PI)]},
    "anisotropic_rotated_pd": {"n_params": 3, "default_guesses": [deg_to_km(7.0), deg_to_km(7.0), 0.0],
                               "default_bounds": [(deg_to_km(0.5), deg_to_km(50.0)),
                                                  (deg_to_km(0.5), deg_to_km(30.0)), (-_TWO_PI, _TWO_PI)]},
}
_MESSAGES = {0: "Optimization terminated successfully.",
             1: "Maximum number of function evaluations has been exceeded.",
             2: "Maximum number of iterations has been exceeded."}


class FitResult(dict):
    """The fields of `scipy.optimize.OptimizeResult` the reference reads (`x`, `fun`, `nit`, `nfev`, `status`,
    `success`, `message`), with attribute access like the original."""

    __getattr__ = dict.__getitem__
    __setattr__ = dict.__setitem__


class EllipseModel:
    """Variogram / ellipse fitting on a standardised correlation matrix (reference `model.py:156-264`).

    Parameters: `anisotropic` (ellipse instead of circle), `rotated` (the ellipse may rotate; needs `anisotropic`),
    `physical_distance` (km instead of degrees), `v` (Matern shape, > 0), `unit_sigma` (fix the likelihood's sigma
    at 1 instead of fitting it)."""

    def __init__(self, anisotropic: bool, rotated: bool, physical_distance: bool, v: float,
                 unit_sigma: bool = False) -> None:
        if v <= 0:
            raise ValueError("'v' must be > 0")
        self.anisotropic = anisotropic
        self.rotated = rotated
        self.physical_distance = physical_distance
        self.v = v
        self.unit_sigma = unit_sigma
        self._get_model_names()
        self.supercategory_params = SUPERCATEGORY_PARAMS[self.supercategory]
        self.supercategory_n_params = len(self.supercategory_params)
        self._get_defaults()

    def _get_model_names(self) -> None:
        if self.rotated and not self.anisotropic:
            raise ValueError("Cannot have an isotropic rotated fform")
        parts = ["anisotropic"] if self.anisotropic else ["isotropic"]
        if self.rotated:
            parts.append("rotated")
        if self.physical_distance:
            parts.append("pd")
        self.fform = "_".join(parts)
        if self.fform not in FFORM_PARAMETERS:
            raise ValueError("Could not compute fform value from inputs")
        self.model_type = FFORM_TO_MODELTYPE[self.fform]
        self.supercategory = MODEL_TYPE_TO_SUPERCATEGORY[self.model_type]

    def _get_defaults(self) -> None:
        params = FFORM_PARAMETERS[self.fform]
        self.n_params: int = params["n_params"]
        self.default_guesses: list[float] = list(params["default_guesses"])
        self.default_bounds: list[tuple[float, float]] = list(params["default_bounds"])
        if self.anisotropic:
            self.cov_ij = lambda X, **p: cov_ij_anisotropic(self.v, 1, X[:, 0], X[:, 1], **p)
        else:
            self.cov_ij = lambda X, **p: cov_ij_isotropic(self.v, 1, X, **p)

    # -- device plumbing ----------------------------------------------------------------------------------------------
    def _csr(self, problems):
        """[(X, y), ...] -> device CSR arrays (offsets, x0, x1 | None, y, total)."""
        ns, x0s, x1s, ys = [], [], [], []
        for X, y in problems:
            X, y = np.asarray(X, dtype=np.float64), np.asarray(y, dtype=np.float64).reshape(-1)
            if self.anisotropic:
                if X.ndim != 2 or X.shape[1] != 2 or X.shape[0] != y.size:
                    raise ValueError(f"anisotropic models need X of shape (n, 2), got {X.shape} for {y.size} values")
                x0s.append(X[:, 0])
                x1s.append(X[:, 1])
            else:
                X = X.reshape(-1)
                if X.size != y.size:
                    raise ValueError(f"X ({X.size}) and y ({y.size}) differ in length")
                x0s.append(X)
            ns.append(y.size)
            ys.append(y)
        offs = np.concatenate([[0], np.cumsum(ns)]).astype(np.int64)
        cat = lambda parts: dev.to_device(np.concatenate(parts) if parts else np.zeros(0), torch.float64)  # noqa: E731
        return (dev.to_device(offs, torch.int64), cat(x0s), cat(x1s) if self.anisotropic else None, cat(ys),
                int(offs[-1]))

    def negative_log_likelihood(self, X: np.ndarray, y: np.ndarray, params: list[float],
                                arctanh_transform: bool = True) -> float:
        """Negative log-likelihood of the observed correlations `y` at displacements `X` given Matern `params`
        (reference `model.py:266-384`): Gaussian in the Fisher-transformed correlations, both sides clamped at
        0.999999.  Unlike the reference, `y` is not clamped in place (the clamp happens on the device copy)."""
        N = self.n_params + (0 if self.unit_sigma else 1)
        p = np.asarray(params, dtype=np.float64).reshape(-1)
        if p.size < N:
            raise ValueError(f"expected {N} parameters, got {p.size}")
        offs, x0, x1, yd, total = self._csr([(X, y)])
        scratch = torch.empty(max(2 * total, 2), dtype=torch.float64, device=x0.device)
        fun = torch.empty(1, dtype=torch.float64, device=x0.device)
        pd = dev.to_device(p[:N].copy(), torch.float64)
        check(lib().gmg_ellipse_nll_f64(offs.data_ptr(), 1, total, x0.data_ptr(), dev.ptr(x1), yd.data_ptr(),
                                        self.n_params, int(self.unit_sigma), int(arctanh_transform), float(self.v),
                                        pd.data_ptr(), scratch.data_ptr(), fun.data_ptr(), dev.stream_ptr()),
              "gmg_ellipse_nll_f64")
        return float(fun.item())

    def negative_log_likelihood_function(self, X: np.ndarray, y: np.ndarray) -> Callable[[list[float]], float]:
        """A function of the parameters alone (reference `model.py:386-400`)."""
        return lambda params: self.negative_log_likelihood(X, y, params)

    def _guesses_bounds(self, guesses, bounds):
        guesses = guesses or self.default_guesses
        bounds = bounds or self.default_bounds
        if (not self.unit_sigma) and len(guesses) != self.n_params + 1:      # model.py:468-470 (appends in place)
            guesses.append(0.1)
            bounds.append((0.0001, 0.5))
        N = self.n_params + (0 if self.unit_sigma else 1)
        if len(guesses) != N or len(bounds) != N:
            raise ValueError(f"expected {N} guesses and bounds, got {len(guesses)} and {len(bounds)}")
        return guesses, bounds

    def fit_many(self, problems, guesses: list[float] | None = None,
                 bounds: list[tuple[float, float]] | None = None, tol: float | None = None,
                 maxiter: int | None = None, maxfev: int | None = None):
        """Fit every (X, y) of `problems` in one launch.  Returns (list of `FitResult`, bounds)."""
        guesses, bounds = self._guesses_bounds(guesses, bounds)
        offs, x0, x1, yd, total = self._csr(problems)
        res = _run_fit(self, offs, x0, x1, yd, total, len(problems), guesses, bounds, tol, maxiter, maxfev)
        return res, bounds

    def fit(self, X: np.ndarray, y: np.ndarray, guesses: list[float] | None = None,
            bounds: list[tuple[float, float]] | None = None, opt_method: str = "Nelder-Mead",
            tol: float | None = None, estimate_SE: str | None = None, n_sim: int = 500, n_jobs: int = 4,
            backend: str = "loky", random_seed: int = 1234):
        """Nelder-Mead maximum-likelihood fit (reference `model.py:402-500`).  Returns (results, SE, bounds) with
        SE = None: the bootstrap standard errors of the reference are not available here."""
        if opt_method != "Nelder-Mead":
            raise NotImplementedError("only the reference's default optimiser, 'Nelder-Mead', runs on the device")
        if estimate_SE is not None:
            raise NotImplementedError("bootstrap standard errors (estimate_SE) are not implemented; pass estimate_SE=None")
        results, bounds = self.fit_many([(X, y)], guesses, bounds, tol)
        return results[0], None, bounds


def _run_fit(model: EllipseModel, offs, x0, x1, y, total: int, n_problems: int, guesses, bounds, tol, maxiter=None,
             maxfev=None) -> list[FitResult]:
    """`gmg_ellipse_fit_f64` on device CSR arrays -> one `FitResult` per problem."""
    N = model.n_params + (0 if model.unit_sigma else 1)
    g = (ctypes.c_double * N)(*[float(v) for v in guesses])
    lo = (ctypes.c_double * N)(*[float(b[0]) for b in bounds])
    hi = (ctypes.c_double * N)(*[float(b[1]) for b in bounds])
    if any(a > b for a, b in zip(lo, hi)):
        raise ValueError("Nelder Mead - one of the lower bounds is greater than an upper bound.")    # scipy's text
    device = x0.device
    xtol = 1e-4 if tol is None else float(tol)                    # minimize(tol=...) sets xatol = fatol = tol
    scratch = torch.empty(max(2 * total, 2), dtype=torch.float64, device=device)
    x_out = torch.empty((n_problems, N), dtype=torch.float64, device=device)
    fun = torch.empty(n_problems, dtype=torch.float64, device=device)
    nit, nfev, status = (torch.empty(n_problems, dtype=torch.int32, device=device) for _ in range(3))
    check(lib().gmg_ellipse_fit_f64(offs.data_ptr(), n_problems, total, x0.data_ptr(), dev.ptr(x1), y.data_ptr(),
                                    model.n_params, int(model.unit_sigma), 1, float(model.v), g, lo, hi, xtol, xtol,
                                    int(maxiter or 0), int(maxfev or 0), scratch.data_ptr(), x_out.data_ptr(),
                                    fun.data_ptr(), nit.data_ptr(), nfev.data_ptr(), status.data_ptr(),
                                    dev.stream_ptr()), "gmg_ellipse_fit_f64")
    xs, fs = x_out.cpu().numpy(), fun.cpu().numpy()
    its, evs, sts = nit.cpu().numpy(), nfev.cpu().numpy(), status.cpu().numpy()
    return [FitResult(x=xs[b].copy(), fun=float(fs[b]), nit=int(its[b]), nfev=int(evs[b]), status=int(sts[b]),
                      success=bool(sts[b] == 0), message=_MESSAGES[int(sts[b])]) for b in range(n_problems)]


# ---- module-level covariance functions (reference `model.py:582-686`) ---------------------------------------------------
def _matern(v: float, inner: np.ndarray) -> np.ndarray:
    """x^v K_v(x) / (Gamma(v) 2^(v-1)) on the device (`gmg_matern_f64`, Temme / CF2 Bessel-K)."""
    x = dev.to_device(np.ascontiguousarray(np.asarray(inner, dtype=np.float64).reshape(-1)), torch.float64)
    out = torch.empty_like(x)
    if x.numel():
        check(lib().gmg_matern_f64(None, float(v), 2, x.data_ptr(), x.numel(), out.data_ptr(), dev.stream_ptr()),
              "gmg_matern_f64")
    return out.cpu().numpy().reshape(np.shape(inner))


def cov_ij_anisotropic(v: float, stdev: float, delta_x, delta_y, Lx: float, Ly: float, stdev_j: float | None = None,
                       theta: float | None = None):
    """Covariance between two points under a locally constant ellipse (reference `model.py:582-644`)."""
    stdev_j = stdev_j or stdev
    dx, dy = np.asarray(delta_x, dtype=np.float64), np.asarray(delta_y, dtype=np.float64)
    lx2, ly2 = Lx ** 2.0, Ly ** 2.0
    if theta is None:
        s00, s01, s11 = lx2, 0.0, ly2
    else:
        c, s = math.cos(theta), math.sin(theta)
        s00, s01, s11 = c * c * lx2 + s * s * ly2, c * s * (lx2 - ly2), s * s * lx2 + c * c * ly2
    det = s00 * s11 - s01 * s01
    if det == 0:
        raise ValueError("Denominator is 0")
    tau = np.sqrt(dx * (dx * (s11 / det) + dy * (-s01 / det)) + dy * (dx * (-s01 / det) + dy * (s00 / det)))
    return (stdev * stdev_j) * _matern(v, 2.0 * tau * np.sqrt(v))


def cov_ij_isotropic(v: float, stdev: float, delta, R: float, stdev_j: float | None = None):
    """The circle: Lx = Ly = R (reference `model.py:647-686`)."""
    stdev_j = stdev_j or stdev
    return (stdev * stdev_j) * _matern(v, 2.0 * (np.abs(np.asarray(delta, dtype=np.float64)) / R) * np.sqrt(v))
