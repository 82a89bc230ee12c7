"""Oracle helper: Matern variogram matrix used by the reference's GeoStats.jl known-answer test.

TEST INFRASTRUCTURE ONLY (see `oracle/__init__.py`).  Follows `MaternVariogram.fit`,
reference variogram.py:296-476 (`_left` :409-411, `_middle` :413-428, `_right` :430-447, fit
:449-476): gamma(h) = psill * (1 - left * middle(h/range) * right(h/range)) + nugget with the value at
h == 0 forced to 0 (K_nu(0) is infinite).
"""

from __future__ import annotations

import numpy as np
from scipy.special import gamma, kv


def matern_variogram(dist: np.ndarray, psill: float, nugget: float, range_: float, nu: float,
                     method: str = "sklearn") -> np.ndarray:
    h = dist / range_
    scale = {"sklearn": np.sqrt(2.0 * nu), "gstat": 1.0, "karspeck": 2.0 * np.sqrt(nu)}[method]
    left = 1.0 / (gamma(nu) * np.power(2.0, nu - 1.0))
    with np.errstate(invalid="ignore"):
        out = psill * (1 - left * np.power(scale * h, nu) * kv(nu, scale * h)) + nugget
    np.fill_diagonal(out, nugget)  # reference variogram.py:470-475 (diagonal set to the nugget)
    return out


def geostats_setup():
    """Inputs of reference test/test_kriging.py:67-192: 20x20 unit grid (lat, lon in 1..20, C order),
    euclidean distances, Matern(range=35/3, psill=4, nugget=0, nu=1.5) variogram matrix passed *as*
    the covariance, three observations."""
    lat, lon = np.meshgrid(np.arange(1.0, 21.0), np.arange(1.0, 21.0), indexing="ij")
    pts = np.column_stack([lat.ravel(), lon.ravel()])
    dist = np.sqrt(((pts[:, None, :] - pts[None, :, :]) ** 2).sum(-1))
    cov = matern_variogram(dist, psill=4.0, nugget=0.0, range_=35 / 3, nu=1.5)
    obs_latlon = [(5.0, 5.0), (15.0, 10.0), (10.0, 15.0)]
    idx = np.array([int((la - 1) * 20 + (lo - 1)) for la, lo in obs_latlon], dtype=np.int64)
    obs = np.array([1.0, 0.0, 1.0])
    return cov, idx, obs


def spherical_variogram(dist, psill, nugget, range_):
    """`SphericalVariogram.fit`, reference variogram.py:122-138."""
    out = 0.5 * psill * (3 * dist / range_ - np.power(dist / range_, 3)) + nugget
    out[dist >= range_] = nugget + psill
    return out


def gaussian_variogram(dist, psill, nugget, range_):
    """`GaussianVariogram.fit`, reference variogram.py:203-215."""
    return psill * (1.0 - np.exp(-(np.power(dist, 2.0) / np.power(range_, 2.0)))) + nugget


def exponential_variogram(dist, psill, nugget, range_):
    """`ExponentialVariogram.fit`, reference variogram.py:283-286."""
    return psill * (1.0 - np.exp(-(dist / range_))) + nugget
