"""`EllipseBuilder`: drop-in for `glomar_gridding.ellipse.EllipseBuilder` (reference `ellipse/estimate.py:41-760`).

Covariance / correlation of a (time, latitude, longitude) training cube over time, and - from the correlations of
every cell with its neighbourhood - the maximum-likelihood ellipse parameters of every unmasked cell.  The
reference loops over cells in Python and calls `scipy.optimize.minimize` per cell (58 s for its 5-degree notebook);
here `compute_params` selects the training pairs of all cells (`gmg_ellipse_train_f64`) and fits all cells
(`gmg_ellipse_fit_f64`) on the device, a few launches in total.  The sample covariance is one DMMA GEMM.

`coords`: the reference takes `xarray.Coordinates`; anything that maps "time" / "latitude" / "longitude" to an object
with a `.values` array (or to an array) works, in that key order.  `compute_params` returns an `xarray.Dataset` when
xarray is importable and a plain dict {parameter: 2-D array} otherwise.
"""

from __future__ import annotations

import ctypes
import math as maths
from typing import Any
from warnings import warn

import numpy as np
import torch

from .. import _dense as dn
from .. import _device as dev
from .._lib import check, lib
from ..constants import RADIUS_OF_EARTH_KM
from ..utils import cov_2_cor, mask_array, uncompress_masked
from .model import EllipseModel, _run_fit

_DISP = {None: -1, "Met_Office": 0, "Modified_Met_Office": 1}


def _values(c):
    return np.asarray(c.values if hasattr(c, "values") and not callable(c.values) else c)


class EllipseBuilder:
    """Spatial covariance / correlation of the training data and the ellipse fit of every cell
    (reference `estimate.py:41-117`)."""

    def __init__(self, data_array: np.ndarray, coords) -> None:
        self.data = mask_array(data_array)
        self.coords = coords
        self.xy_shape = self.data[0].shape
        if len(self.xy_shape) != 2:
            raise ValueError("Time slice maps should be 2D; check extra dims (ensemble?)")
        self.big_covar_size = np.prod(self.xy_shape)
        self._parse_coords()
        self._detect_mask()
        self.calc_cov()

    def _parse_coords(self) -> None:
        self.tcoord_pos = -1
        self.xycoords_pos = []
        self.xycoords_name = []
        for i, coord in enumerate(self.coords):
            if coord == "time":
                self.tcoord_pos = i
            if coord in ["latitude", "longitude"]:
                self.xycoords_pos.append(i)
                self.xycoords_name.append(coord)
        if self.tcoord_pos == -1:
            raise ValueError("Input cube needs a time dimension")
        if self.tcoord_pos != 0:
            raise ValueError("Input cube time dimension not at 0")
        if len(self.xycoords_pos) != 2:
            raise ValueError("Input cube need two spatial dimensions ('latitude' and 'longitude')")
        self.xycoords_pos = tuple(self.xycoords_pos)
        lon, lat = _values(self.coords["longitude"]), _values(self.coords["latitude"])
        self.xx, self.yy = np.meshgrid(lon, lat)
        self.xi, self.yi = np.meshgrid(np.arange(len(lon)), np.arange(len(lat)))
        self.time_n = len(_values(self.coords["time"]))

    def _detect_mask(self) -> None:
        self.data_has_mask = np.ma.is_masked(self.data)
        if self.data_has_mask:
            # cells masked at ANY time are discarded (reference :126-132)
            self.mask = np.any(np.ma.getmaskarray(self.data), axis=0)
            self.mask_1D = self.mask.flatten()
            self._self_mask()
            self.small_covar_size = np.sum(np.logical_not(self.mask))
        else:
            self.mask = np.zeros_like(np.ma.getdata(self.data[0]), dtype=bool)
            self.mask_1D = self.mask.flatten()
            self.small_covar_size = self.big_covar_size
        self.x_masked = np.ma.masked_where(self.mask, self.xx)
        self.y_masked = np.ma.masked_where(self.mask, self.yy)
        self.xi_masked = np.ma.masked_where(self.mask, self.xi).compressed()
        self.yi_masked = np.ma.masked_where(self.mask, self.yi).compressed()
        self.xy_masked = np.column_stack([self.x_masked.compressed(), self.y_masked.compressed()])
        self.xy_full = np.column_stack([self.x_masked.flatten(), self.y_masked.flatten()])

    def _self_mask(self) -> None:
        self.data = np.ma.masked_where(np.broadcast_to(self.mask, self.data.shape), self.data)

    def calc_cov(self, rounding: int | None = None) -> None:
        """Covariance and correlation over time of the unmasked cells (reference `:150-180`): anomalies with respect
        to each cell's own time mean, `A^T A / (n_time - 1)` as one DMMA GEMM, correlation by `cov_2_cor` - both kept
        on the device (`cov_device`, `cor_device`) and mirrored to numpy (`cov`, `cor`)."""
        flat = np.ma.getdata(self.data).reshape((self.time_n, int(self.big_covar_size)))[:, ~self.mask_1D]
        flat = np.asarray(flat, dtype=np.float64)
        flat = flat - np.mean(flat, axis=0, keepdims=True)
        A = dn.padded(flat)                                             # (t, n) with an even leading dimension
        n = flat.shape[1]
        cov = dn.gemm(A[:, :n], A[:, :n], transa=True, alpha=1.0 / (self.time_n - 1))[:n, :n]
        self.cov_device = cov.contiguous()
        self.cov = dev.to_host(self.cov_device)
        if rounding is not None:
            self.cov = np.round(self.cov, rounding)
            self.cov_device = dev.to_device(self.cov, torch.float64)
        self.cor_device = cov_2_cor(self.cov_device)
        self.cor = dev.to_host(self.cor_device)
        if rounding is not None:
            self.cor = np.round(self.cor, rounding)
            self.cor_device = dev.to_device(self.cor, torch.float64)

    # -- training data --------------------------------------------------------------------------------------------
    def _train_device(self, cells: np.ndarray, anisotropic: bool, physical_distance: bool, delta_x_method,
                      physical_distance_selection: bool, min_distance: float, max_distance: float):
        """CSR training pairs of `cells` on the device: (offsets, x0, x1 | None, y, total, counts on the host)."""
        if physical_distance and delta_x_method is None:
            raise ValueError("Cannot have physical_distance with unset delta_x_method")
        if delta_x_method not in _DISP:
            raise ValueError(f"Unknown 'delta_x_method' value, got '{delta_x_method}'")
        lat = dev.to_device(np.ascontiguousarray(self.xy_masked[:, 1], dtype=np.float64), torch.float64)
        lon = dev.to_device(np.ascontiguousarray(self.xy_masked[:, 0], dtype=np.float64), torch.float64)
        cells_t = dev.to_device(np.asarray(cells, dtype=np.int64), torch.int64)
        cor = self.cor_device
        n_cells, N = int(cells_t.numel()), int(lat.numel())
        counts = torch.zeros(n_cells, dtype=torch.int64, device=lat.device)
        args = (lat.data_ptr(), lon.data_ptr(), N, cor.data_ptr(), cor.stride(0), cells_t.data_ptr(), n_cells,
                int(anisotropic), int(physical_distance), _DISP[delta_x_method], int(physical_distance_selection),
                float(min_distance), float(max_distance))
        check(lib().gmg_ellipse_train_f64(*args, 0, counts.data_ptr(), None, None, None, None, dev.stream_ptr()),
              "gmg_ellipse_train_f64")
        offs = torch.zeros(n_cells + 1, dtype=torch.int64, device=lat.device)
        offs[1:] = torch.cumsum(counts, 0)
        total = int(offs[-1].item())
        x0 = torch.empty(max(total, 1), dtype=torch.float64, device=lat.device)
        x1 = torch.empty(max(total, 1), dtype=torch.float64, device=lat.device) if anisotropic else None
        y = torch.empty(max(total, 1), dtype=torch.float64, device=lat.device)
        if total:
            check(lib().gmg_ellipse_train_f64(*args, 1, None, offs.data_ptr(), x0.data_ptr(), dev.ptr(x1),
                                              y.data_ptr(), dev.stream_ptr()), "gmg_ellipse_train_f64")
        return offs, x0, x1, y, total, counts.cpu().numpy()

    def _get_train_data(self, xy_point: int, min_distance: float, max_distance: float, anisotropic: bool,
                        delta_x_method, physical_distance: bool = True,
                        physical_distance_selection: bool = True) -> tuple[np.ndarray, np.ndarray]:
        """(X_train, y_train) of one cell (reference `:441-513`), selected and computed on the device."""
        offs, x0, x1, y, total, _ = self._train_device(np.array([xy_point]), anisotropic, physical_distance,
                                                       delta_x_method, physical_distance_selection, min_distance,
                                                       max_distance)
        x0h, yh = x0[:total].cpu().numpy(), y[:total].cpu().numpy()
        if anisotropic:
            return np.column_stack([x0h, x1[:total].cpu().numpy()]), yh
        return x0h, yh

    # -- fits -----------------------------------------------------------------------------------------------------
    def _finish(self, xy_point: int, matern_ellipse: EllipseModel, results, bounds) -> list:
        """ModelParams of one cell from its optimiser result (reference `:361-398`)."""
        model_params = results.x.tolist()
        self._check_params(matern_ellipse, model_params)
        stdev = None
        if not matern_ellipse.unit_sigma:
            stdev = model_params.pop()
        fit_success = _get_fit_score(model_params, bounds, results.nit) if results.success else 9
        std_dev = np.sqrt(self.cov[xy_point, xy_point])
        model_params.append(std_dev)
        model_params.append(fit_success)
        model_params.append(results.nit)
        return model_params, fit_success, std_dev, stdev

    def fit_ellipse_model(self, xy_point: int, matern_ellipse: EllipseModel, max_distance: float = 6000,
                          min_distance: float = 0.3, delta_x_method="Modified_Met_Office",
                          guesses: list[float] | None = None, bounds: list[tuple[float, float]] | None = None,
                          opt_method: str = "Nelder-Mead", tol: float = 0.001, estimate_SE: str | None = None,
                          n_jobs: int = 4, n_sim: int = 500,
                          physical_distance_selection: bool = True) -> dict[str, Any] | None:
        """Fit the ellipse of one cell (reference `:187-400`)."""
        R2 = uncompress_masked(self.cor[xy_point, :], self.mask_1D).reshape(self.xy_shape)
        X_train, y_train = self._get_train_data(
            xy_point=xy_point, min_distance=min_distance, max_distance=max_distance,
            anisotropic=matern_ellipse.anisotropic, delta_x_method=delta_x_method,
            physical_distance=matern_ellipse.physical_distance, physical_distance_selection=physical_distance_selection)
        if len(y_train) == 0:
            warn(f"No training data for idx {xy_point}")
            return None
        results, SE, bounds = matern_ellipse.fit(X_train, y_train, guesses=guesses, bounds=bounds,
                                                 opt_method=opt_method, tol=tol, estimate_SE=estimate_SE,
                                                 n_jobs=n_jobs, n_sim=n_sim)
        model_params, fit_success, std_dev, stdev = self._finish(xy_point, matern_ellipse, results, bounds)
        return {"Correlation": R2, "Results": results, "ModelParams": model_params, "Success": fit_success,
                "StandardDeviation": std_dev, "StandardError": SE, "RMSE": stdev}

    def _check_params(self, ellipse: EllipseModel, model_params: list[Any]) -> None:
        """Ensure Lx > Ly and theta in (-pi, pi] - in place, exactly as the reference does (`:402-422`; note that for
        an un-rotated anisotropic model index 2 is the likelihood's sigma, which the reference shifts as well)."""
        if ellipse.anisotropic and model_params[1] > model_params[0]:
            model_params[0], model_params[1] = model_params[1], model_params[0]
            model_params[2] += np.pi / 2
        if not ellipse.rotated:
            return None
        if model_params[2] > np.pi:
            model_params[2] -= np.pi
        if model_params[2] <= -np.pi:
            model_params[2] += np.pi
        return None

    def compute_params(self, default_value: Any, matern_ellipse: EllipseModel, max_distance: float = 6000,
                       min_distance: float = 0.3, delta_x_method="Modified_Met_Office",
                       guesses: list[float] | None = None, bounds: list[tuple[float, float]] | None = None,
                       opt_method: str = "Nelder-Mead", tol: float = 1e-4, estimate_SE: str | None = None,
                       n_jobs: int = 4, n_sim: int = 500, physical_distance_selection: bool = True):
        """Ellipse parameters of ALL unmasked cells (reference `:515-715`): one training-pair selection and one
        batched Nelder-Mead launch instead of a Python loop.  Returns {parameter name: (nlat, nlon) array} - the
        parameters of `matern_ellipse.supercategory_params`, `default_value` where there is no fit - as an
        `xarray.Dataset` if xarray is available."""
        if opt_method != "Nelder-Mead":
            raise NotImplementedError("only the reference's default optimiser, 'Nelder-Mead', runs on the device")
        if estimate_SE is not None:
            raise NotImplementedError("bootstrap standard errors (estimate_SE) are not implemented; pass estimate_SE=None")
        names = list(matern_ellipse.supercategory_params.keys())
        defaults = list(default_value) if isinstance(default_value, (list, tuple, np.ndarray)) else [default_value] * len(names)
        if len(defaults) != len(names):
            raise ValueError("Length of default_value must equal the number of parameters of the EllipseModel")
        params = {n: np.full(self.xy_shape, d, dtype=np.float64) for n, d in zip(names, defaults)}
        n_cells = int(self.small_covar_size)
        cells = np.arange(n_cells, dtype=np.int64)
        offs, x0, x1, y, total, counts = self._train_device(
            cells, matern_ellipse.anisotropic, matern_ellipse.physical_distance, delta_x_method,
            physical_distance_selection, min_distance, max_distance)
        g, b = matern_ellipse._guesses_bounds(guesses, bounds)
        results = _run_fit(matern_ellipse, offs, x0, x1, y, total, n_cells, g, b, tol)
        self.fit_results = results
        for mask_i, (grid_i, grid_j) in enumerate(zip(self.xi_masked, self.yi_masked)):
            if counts[mask_i] == 0:
                warn(f"No training data for idx {mask_i}")
                continue
            model_params, *_ = self._finish(mask_i, matern_ellipse, results[mask_i], b)
            for i, name in enumerate(names):
                params[name][grid_j, grid_i] = model_params[i]
        try:
            import xarray as xr

            lat, lon = _values(self.coords["latitude"]), _values(self.coords["longitude"])
            return xr.Dataset({k: (("latitude", "longitude"), v) for k, v in params.items()},
                              coords={"latitude": lat, "longitude": lon})
        except ImportError:
            return params

    def find_nearest_xy_index_in_cov_matrix(self, lonlat: list[float], use_full: bool = False) -> tuple[int, np.ndarray]:
        """Nearest row / column of the covariance to a (lon, lat) position (reference `:717-730`)."""
        lon, lat, *_ = lonlat
        a = self.xy_full if use_full else self.xy_masked
        idx = int(((a[:, 0] - lon) ** 2.0 + (a[:, 1] - lat) ** 2.0).argmin())
        return idx, a[idx, :]

    def _xy_2_xy_full_index(self, xy_point: int) -> int:
        """Index among ALL cells of the xy_point-th unmasked cell (reference `:732-745`)."""
        return int(np.argwhere(np.all((self.xy_full - self.xy_masked[xy_point, :]) == 0, axis=1))[0])

    def __str__(self):
        return str(self.__class__)


def _get_fit_score(model_params, bounds, niter) -> int:
    """QC code of a converged fit (reference `:751-781`): 0 ok, 1 / 2 one parameter at a lower / upper bound, 3 several."""
    fit_success = 0
    for model_param, bb in zip(model_params, bounds):
        if maths.isclose(model_param, bb[0], rel_tol=0.01):
            fit_success = 1 if fit_success == 0 else 3
        if maths.isclose(model_param, bb[1], rel_tol=0.01):
            fit_success = 2 if fit_success == 0 else 3
    return fit_success
