"""Oracle for the ellipse-parameter FIT (not yet on the device): the per-cell maximum-likelihood
estimate of (Lx, Ly, theta) that produces the inputs of `EllipseCovarianceBuilder`.

TEST INFRASTRUCTURE ONLY (see `oracle/__init__.py`).  Restates, from the formulas, the computational core of
`glomar_gridding/ellipse/model.py` in the reference:

    cov_ij_anisotropic / cov_ij_isotropic           model.py:582-686   (Matern correlation of a displacement, PS-2006
                                                                       with locally constant Sigma; distances.py:41-58,
                                                                       :328-353, :356-474 for the 2 x 2 algebra)
    EllipseModel.negative_log_likelihood            model.py:266-384   (Gaussian log-likelihood of the Fisher-transformed
                                                                       correlations, 0.999999 clamp)
    EllipseModel.fit (estimate_SE=None)             model.py:402-500   (scipy Nelder-Mead with bounds)
    default guesses / bounds per functional form    model.py:104-153

Parity status: PINNED - `tests/test_oracle_vs_golden.py::test_ellipse_fit_*` compares values, likelihoods and fitted
parameters with fixtures produced by the unmodified reference (`tests/golden/make_golden_fit.py`) and, where the
reference sources are present, with the live reference.  The bootstrap standard errors (`model.py:500-579`, joblib)
and the training-set selection of `ellipse/estimate.py` are NOT restated yet.
"""

from __future__ import annotations

import math
import warnings

import numpy as np
from scipy import stats
from scipy.optimize import minimize
from scipy.special import gamma, kv

KM_PER_DEG = 1.852 * 60.0          # utils.deg_to_km: nautical miles per degree of latitude (60) x km per nm (1.852)

# model.py:104-153: (number of parameters, default guesses, default bounds) per functional form
_DEG = {"isotropic": (1, [7.0], [(0.5, 50.0)]),
        "anisotropic": (2, [7.0, 7.0], [(0.5, 50.0), (0.5, 30.0)]),
        "anisotropic_rotated": (3, [7.0, 7.0, 0.0], [(0.5, 50.0), (0.5, 30.0), (-2.0 * np.pi, 2.0 * np.pi)])}


def form_parameters(fform: str):
    """`FFORM_PARAMETERS[fform]` (model.py:104-153); the `_pd` forms are the degree forms scaled to km."""
    base = fform[:-3] if fform.endswith("_pd") else fform
    n, guesses, bounds = _DEG[base]
    if not fform.endswith("_pd"):
        return n, list(guesses), [tuple(b) for b in bounds]
    k = KM_PER_DEG
    g = [x * k if i < 2 else x for i, x in enumerate(guesses)]
    b = [(lo * k, hi * k) if i < 2 else ((-2.0 * math.pi, 2.0 * math.pi)) for i, (lo, hi) in enumerate(bounds)]
    return n, g, b


def sigma_rot(Lx: float, Ly: float, theta: float | None) -> np.ndarray:
    """R(theta) diag(Lx^2, Ly^2) R(theta)^T (distances.py:328-353, rotation matrix :41-50)."""
    L = np.diag([Lx ** 2.0, Ly ** 2.0])
    if theta is None:
        return L
    c, s = math.cos(theta), math.sin(theta)
    R = np.array([[c, -s], [s, c]])
    return R @ L @ R.T


def mahalanobis(delta_x, delta_y, Lx: float, Ly: float, theta: float | None = None):
    """sqrt(d^T Sigma^-1 d) (distances.py:356-474, 2 x 2 inverse :53-58)."""
    sigma = sigma_rot(Lx, Ly, theta) if theta is not None else np.diag(np.array([Lx ** 2.0, Ly ** 2.0]))
    det = sigma[0, 0] * sigma[1, 1] - sigma[0, 1] * sigma[1, 0]
    if det == 0:
        raise ValueError("Denominator is 0")
    inv = np.array([[sigma[1, 1], -sigma[0, 1]], [-sigma[1, 0], sigma[0, 0]]]) / det
    return np.sqrt(delta_x * (delta_x * inv[0, 0] + delta_y * inv[0, 1])
                   + delta_y * (delta_x * inv[1, 0] + delta_y * inv[1, 1]))


def cov_ij_anisotropic(v, stdev, delta_x, delta_y, Lx, Ly, stdev_j=None, theta=None):
    """model.py:582-641: sigma_i sigma_j / (Gamma(v) 2^(v-1)) x^v K_v(x), x = 2 sqrt(v) tau."""
    stdev_j = stdev_j or stdev
    tau = mahalanobis(delta_x, delta_y, Lx, Ly, theta=theta)
    first = (stdev * stdev_j) / (gamma(v) * (2.0 ** (v - 1)))
    inner = 2.0 * tau * np.sqrt(v)
    return first * np.power(inner, v) * kv(v, inner)


def cov_ij_isotropic(v, stdev, delta, R, stdev_j=None):
    """model.py:650-686: the circle, tau = |delta| / R."""
    stdev_j = stdev_j or stdev
    inner = 2.0 * (np.abs(delta) / R) * np.sqrt(v)
    first = (stdev * stdev_j) / (gamma(v) * (2.0 ** (v - 1)))
    return first * (inner ** v) * kv(v, inner)


class EllipseModelOracle:
    """`EllipseModel(anisotropic, rotated, physical_distance, v, unit_sigma)` (model.py:156-264): the functional form,
    its defaults, the negative log-likelihood and the Nelder-Mead fit."""

    ARCTANH_THRESHOLD = 0.999999        # model.py:330

    def __init__(self, anisotropic: bool, rotated: bool, physical_distance: bool, v: float, unit_sigma: bool = False):
        if v <= 0:
            raise ValueError("'v' must be > 0")
        if rotated and not anisotropic:
            raise ValueError("Cannot have an isotropic rotated fform")
        self.v, self.unit_sigma = v, unit_sigma
        self.fform = ("anisotropic" if anisotropic else "isotropic") + ("_rotated" if rotated else "") + \
                     ("_pd" if physical_distance else "")
        self.n_params, self.default_guesses, self.default_bounds = form_parameters(self.fform)
        self.anisotropic = anisotropic

    def cov_ij(self, X, params):
        p = list(params)
        if self.n_params == 1:
            return cov_ij_isotropic(self.v, 1, X, p[0])
        theta = p[2] if self.n_params == 3 else None
        return cov_ij_anisotropic(self.v, 1, X[:, 0], X[:, 1], p[0], p[1], theta=theta)

    def negative_log_likelihood(self, X, y, params, arctanh_transform: bool = True) -> float:
        """model.py:266-384.  Like the reference, `y` is clamped IN PLACE when |y| reaches the threshold."""
        sigma = 1 if self.unit_sigma else params[self.n_params]
        y_LL = self.cov_ij(X, params)
        if not arctanh_transform:
            return -1.0 * np.sum(stats.norm.logpdf(y, loc=y_LL, scale=sigma))
        thr = self.ARCTANH_THRESHOLD
        max_abs_y, max_abs_yLL = np.max(np.abs(y)), np.max(np.abs(y_LL))
        if max_abs_y >= thr:
            warnings.warn(f"abs(y) >= {thr} detected; fudged to threshold; max(abs(y)) = {max_abs_y}", RuntimeWarning)
            big = np.abs(y) > thr
            y[big] = np.sign(y[big]) * thr
        if max_abs_yLL >= 1:
            warnings.warn(f"abs(y_LL) >= {thr} detected; fudged to threshold; max(abs(y_LL))={max_abs_yLL}",
                          RuntimeWarning)
            big = np.abs(y_LL) > thr
            y_LL[big] = np.sign(y_LL[big]) * thr
        return -1.0 * np.sum(stats.norm.logpdf(np.arctanh(y), loc=np.arctanh(y_LL), scale=sigma))

    def fit(self, X, y, guesses=None, bounds=None, opt_method: str = "Nelder-Mead", tol=None):
        """model.py:402-500 with `estimate_SE=None`: (OptimizeResult, None, bounds).  The sigma of the likelihood is
        appended to the parameters (guess 0.1, bounds (1e-4, 0.5)) unless `unit_sigma`."""
        guesses = list(guesses or self.default_guesses)
        bounds = list(bounds or self.default_bounds)
        if (not self.unit_sigma) and len(guesses) != self.n_params + 1:
            guesses.append(0.1)
            bounds.append((0.0001, 0.5))
        res = minimize(lambda p: self.negative_log_likelihood(X, y, p), guesses, bounds=bounds, method=opt_method,
                       tol=tol)
        return res, None, bounds
