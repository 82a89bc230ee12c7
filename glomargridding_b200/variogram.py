"""Variogram models on the GPU: drop-ins for `glomar_gridding.variogram`.

Same dataclasses, fields, `__post_init__` range logic and `fit(distance_matrix)` semantics as the reference
(`glomar_gridding/variogram.py:33-504`); the elementwise evaluation runs in `gmg_variogram_f64`.  `fit` accepts a
numpy array or a CUDA tensor (e.g. `distances.haversine_distance(..., return_device=True)`) and returns the same
kind.  Scalar parameters only (the reference also broadcasts array-valued parameters).  No CPU fallback.
"""

from __future__ import annotations

from abc import ABC, abstractmethod
from dataclasses import dataclass
from typing import Literal

import numpy as np
import torch

from . import _device as dev
from ._lib import check, lib

MaternModel = Literal["sklearn", "gstat", "karspeck"]
_MODEL = {"spherical": 0, "gaussian": 1, "exponential": 2, "matern": 3}
_METHOD = {"sklearn": 0, "gstat": 1, "karspeck": 2}
_RANGE_ERR = "range parameter must not be None, " + "it wasn't set by the __post_init__ method."


def _evaluate(model: str, distance_matrix, psill, nugget, range_, nu=0.5, method="sklearn", fill_diag=False,
              variance=None):
    for name, val in (("psill", psill), ("nugget", nugget), ("range", range_), ("nu", nu)):
        if np.ndim(val) != 0:
            raise NotImplementedError(f"array-valued '{name}' is not supported by the device variogram models")
    on_device = isinstance(distance_matrix, torch.Tensor)
    D = dev.to_device(distance_matrix, torch.float64)
    if D.dim() != 2:
        raise ValueError("distance_matrix must be 2-d")
    rows, cols = D.shape
    out = torch.empty_like(D)
    check(lib().gmg_variogram_f64(_MODEL[model], D.data_ptr(), rows, cols, D.stride(0), float(psill), float(nugget),
                                  float(range_), float(nu), _METHOD[method], int(fill_diag and rows == cols),
                                  int(variance is not None), 0.0 if variance is None else float(variance),
                                  out.data_ptr(), out.stride(0), dev.stream_ptr()), "gmg_variogram_f64")
    return out if on_device else dev.to_host(out)


@dataclass()
class Variogram(ABC):
    """Generic variogram model class (reference `variogram.py:33-55`)."""

    @abstractmethod
    def fit(self, distance_matrix):
        raise NotImplementedError("Not implemented for base Variogram class")


def _need_range(obj) -> None:
    if obj.range is None and obj.effective_range is None:
        raise ValueError("One of 'effective_range' and 'range' must be set")


@dataclass()
class SphericalVariogram(Variogram):
    """Spherical model (reference `variogram.py:58-138`)."""

    psill: float
    nugget: float
    effective_range: float | None = None
    range: float | None = None

    def __post_init__(self):
        _need_range(self)
        if self.range is None:
            self.range = self.effective_range
        if self.effective_range is None:
            self.effective_range = self.range

    def fit(self, distance_matrix):
        if self.range is None:
            raise ValueError(_RANGE_ERR)
        return _evaluate("spherical", distance_matrix, self.psill, self.nugget, self.range)


@dataclass()
class GaussianVariogram(Variogram):
    """Gaussian model (reference `variogram.py:141-218`); effective range = 2 * range."""

    psill: float
    nugget: float
    effective_range: float | None = None
    range: float | None = None

    def __post_init__(self):
        _need_range(self)
        if self.range is None:
            self.range = self.effective_range / 2
        if self.effective_range is None:
            self.effective_range = self.range * 2

    def fit(self, distance_matrix):
        if self.range is None:
            raise ValueError(_RANGE_ERR)
        return _evaluate("gaussian", distance_matrix, self.psill, self.nugget, self.range)


@dataclass()
class ExponentialVariogram(Variogram):
    """Exponential model (reference `variogram.py:221-289`); effective range = 3 * range."""

    psill: float
    nugget: float
    range: float | None = None
    effective_range: float | None = None

    def __post_init__(self):
        _need_range(self)
        if self.range is None:
            self.range = self.effective_range / 3
        if self.effective_range is None:
            self.effective_range = self.range * 3

    def fit(self, distance_matrix):
        if self.range is None:
            raise ValueError(_RANGE_ERR)
        return _evaluate("exponential", distance_matrix, self.psill, self.nugget, self.range)


@dataclass()
class MaternVariogram(Variogram):
    """Matern model in the "sklearn", "gstat" or "karspeck" scaling (reference `variogram.py:295-476`)."""

    psill: float
    nugget: float
    effective_range: float | None = None
    range: float | None = None
    nu: float = 0.5
    method: MaternModel = "sklearn"

    def __post_init__(self) -> None:
        _need_range(self)
        if self.range is None:
            self.range = self.effective_range / 2 if 0.5 <= self.nu <= 10 else self.effective_range / 3
        elif self.effective_range is None:
            self.effective_range = self.range * 2 if 0.5 <= self.nu <= 10 else self.range * 3

    def fit(self, distance_matrix):
        if self.range is None:
            raise ValueError(_RANGE_ERR)
        if self.method not in _METHOD:
            raise ValueError("Unexpected 'method' value")
        return _evaluate("matern", distance_matrix, self.psill, self.nugget, self.range, self.nu, self.method,
                         fill_diag=True)


def _model_of(variogram: Variogram) -> tuple[str, float, str, bool]:
    if isinstance(variogram, SphericalVariogram):
        return "spherical", 0.5, "sklearn", False
    if isinstance(variogram, GaussianVariogram):
        return "gaussian", 0.5, "sklearn", False
    if isinstance(variogram, ExponentialVariogram):
        return "exponential", 0.5, "sklearn", False
    if isinstance(variogram, MaternVariogram):
        if variogram.method not in _METHOD:
            raise ValueError("Unexpected 'method' value")
        return "matern", variogram.nu, variogram.method, True
    raise TypeError("variogram must be one of the Spherical/Gaussian/Exponential/Matern models")


def stationary_covariance(lats_deg, lons_deg, variogram: Variogram, variance=None, radius: float = 6371.0,
                          return_device: bool = False):
    """The reference workflow's stationary route in ONE pass (notebook cells 6-7):

        dist = haversine_distance_from_frame(frame)          # distances.py:191-218
        cov = variogram_to_covariance(variogram.fit(dist), variance)    # variogram.py:96-476, :479-504

    for cells at (`lats_deg`, `lons_deg`), decimal degrees, in the frame's row order.  The N x N result is written
    once (`gmg_stationary_cov_f64`: distance, model and `variance - variogram` fused, upper triangle evaluated and
    mirrored) and is bit-identical to the three separate calls of this package.  `variance=None` returns the
    variogram matrix itself."""
    model, nu, method, fill = _model_of(variogram)
    if variogram.range is None:
        raise ValueError(_RANGE_ERR)
    for name, val in (("psill", variogram.psill), ("nugget", variogram.nugget), ("range", variogram.range), ("nu", nu)):
        if np.ndim(val) != 0:
            raise NotImplementedError(f"array-valued '{name}' is not supported by the device variogram models")
    if variance is not None and np.ndim(variance) != 0:
        raise NotImplementedError("array-valued 'variance' is not supported by the fused stationary covariance")
    lat = dev.to_device(np.radians(np.asarray(lats_deg, dtype=np.float64)), torch.float64)
    lon = dev.to_device(np.radians(np.asarray(lons_deg, dtype=np.float64)), torch.float64)
    if lat.shape != lon.shape or lat.dim() != 1:
        raise ValueError("lats and lons must be 1-d arrays of equal length")
    N = lat.numel()
    out = torch.empty((N, N), dtype=torch.float64, device=lat.device)
    check(lib().gmg_stationary_cov_f64(_MODEL[model], lat.data_ptr(), lon.data_ptr(), N, float(radius),
                                       float(variogram.psill), float(variogram.nugget), float(variogram.range),
                                       float(nu), _METHOD[method], int(fill), int(variance is not None),
                                       0.0 if variance is None else float(variance), out.data_ptr(), out.stride(0),
                                       dev.stream_ptr()), "gmg_stationary_cov_f64")
    return out if return_device else dev.to_host(out)


def variogram_to_covariance(variogram, variance):
    """Covariance from a variogram matrix: variance - variogram (reference `variogram.py:479-504`)."""
    return variance - variogram
