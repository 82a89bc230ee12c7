"""`EllipseCovarianceBuilder`: drop-in for `glomar_gridding.ellipse.EllipseCovarianceBuilder`.

Same constructor signature, attributes and methods as the reference class
(`glomar_gridding/ellipse/covariance.py:40-509`); the N x N build runs as one tiled FP64 CUDA kernel
(`gmg_ellipse_cov_f64`, `csrc/ellipse_cov.cu`) instead of the reference's pair-index / numpy
pipeline.  Host code here only does what the reference's constructor does on O(N) data (casting,
mask compression, `np.deg2rad`); all per-pair arithmetic is on the GPU and there is no CPU fallback.

Differences from the reference, by design:

* `covariance_method` / `batch_size` are validated exactly like the reference but do not change how
  the matrix is computed (one kernel; no 10 000-point "low_memory" switch, no warning).
* The result lives on the GPU (`cov_ns_device`, a CUDA `torch.Tensor`).  `cov_ns` is materialised as
  a numpy array (dtype `precision`) on first access, so chaining the builder into the kriging classes
  never round-trips the matrix through the host.
* `rows=(row0, nrows)` builds only a row block (how the build shards across GPUs).
* With `precision=np.float32` the inputs are rounded to float32 exactly where the reference rounds them
  (`:121-129`), the arithmetic is FP64 and the result is rounded to float32 on store; the reference's
  float32 path differs from its own float64 path by ~1e-6 relative, which is also how close the two
  implementations are in that mode.
* The reference's progress `print`s are not reproduced.
"""

from __future__ import annotations

from typing import Any

import numpy as np
import torch

from .. import _device as dev
from .._lib import check, lib
from ..constants import RADIUS_OF_EARTH_KM
from ..types import CovarianceMethod, DeltaXMethod
from ..utils import cov_2_cor, mask_array

import ctypes

_DISP = {"Met_Office": 0, "Modified_Met_Office": 1}
_MATERN = {"auto": 0, "table": 1, "direct": 2}


class EllipseCovarianceBuilder:
    """Compute covariance from ellipse parameters and positions on a B200.

    Parameters are those of the reference class (`ellipse/covariance.py:95-109`):
    `Lx, Ly, theta, stdev` (2-D arrays or MaskedArrays, theta in radians), `lats`, `lons` (degrees),
    `v` (Matern shape), `delta_x_method`, `max_dist` (km or None), `precision`, `covariance_method`,
    `batch_size`.

    Extensions (keyword-only): `rows=(row0, nrows)` builds only that row block of the compressed
    matrix; `matern_mode` in {"auto", "table", "direct"} selects how K_nu is evaluated; `peer_blocks`
    (`ellipse.p2p.PeerRowBlocks`) builds this rank's row block of a multi-GPU build in which every pair
    is evaluated once across the box and mirrored into the owning peer over NVLink.
    """

    def __init__(
        self,
        Lx: np.ndarray,
        Ly: np.ndarray,
        theta: np.ndarray,
        stdev: np.ndarray,
        lats: np.ndarray,
        lons: np.ndarray,
        v: float,
        delta_x_method: DeltaXMethod | None = "Modified_Met_Office",
        max_dist: float | None = None,
        precision=np.float32,
        covariance_method: CovarianceMethod = "array",
        batch_size: int | None = None,
        *,
        rows: tuple[int, int] | None = None,
        matern_mode: str = "auto",
        peer_blocks=None,
    ) -> None:
        if max_dist is not None and not isinstance(max_dist, (int, float)):
            raise ValueError("max_dist must be a number")  # reference :116-117
        if precision not in (np.float32, np.float64):
            raise ValueError("precision must be numpy.float32 or numpy.float64")
        if matern_mode not in _MATERN:
            raise ValueError(f"Unknown matern_mode: {matern_mode}")

        self.v = v
        self.precision = precision
        self.Lx = mask_array(np.ma.asanyarray(Lx).astype(precision))
        self.Ly = mask_array(np.ma.asanyarray(Ly).astype(precision))
        self.theta = mask_array(np.ma.asanyarray(theta).astype(precision))
        self.stdev = mask_array(np.ma.asanyarray(stdev).astype(precision))
        self.max_dist = max_dist
        self.delta_x_method: DeltaXMethod | None = delta_x_method
        self.lats = np.asarray(lats).astype(precision)
        self.lons = np.asarray(lons).astype(precision)
        self.covariance_method: CovarianceMethod = covariance_method
        self.batch_size = batch_size
        self.matern_mode = matern_mode

        self.xy_shape = self.Lx.shape
        self.n_elements = np.prod(self.xy_shape)

        self._get_mask()
        self._peer_blocks = peer_blocks
        if peer_blocks is not None:
            rows = (peer_blocks.row0, peer_blocks.nrows)
        self._rows = rows
        self._calculate_covariance()

    # -- host-side O(N) set-up, as reference `_get_mask` :149-180 --------------------------------
    def _get_mask(self) -> None:
        self.data_has_mask = np.ma.is_masked(self.Lx)
        if self.data_has_mask:
            self.data_mask = np.ma.getmaskarray(self.Lx)
            self.covar_size = np.sum(np.logical_not(self.data_mask))
        else:
            self.data_mask = np.zeros_like(np.ma.getdata(self.Lx), dtype=bool)
            self.covar_size = self.n_elements
        self.Lx_compressed = self.Lx.compressed()
        self.Ly_compressed = self.Ly.compressed()
        self.theta_compressed = self.theta.compressed()
        self.stdev_compressed = self.stdev.compressed()

        self.x_grid, self.y_grid = np.meshgrid(self.lons, self.lats)  # row-major (lat, lon), :166
        self.x_mask = np.ma.masked_where(self.data_mask, self.x_grid)
        self.y_mask = np.ma.masked_where(self.data_mask, self.y_grid)
        self.lat_grid_compressed = self.y_mask.compressed()
        self.lon_grid_compressed = self.x_mask.compressed()
        self.lat_grid_compressed_rad = np.deg2rad(self.lat_grid_compressed)  # :171-172
        self.lon_grid_compressed_rad = np.deg2rad(self.lon_grid_compressed)
        self.xy_compressed = np.column_stack([self.lon_grid_compressed, self.lat_grid_compressed])
        self.xy_full = np.column_stack([self.x_mask.flatten(), self.y_mask.flatten()])

    def _validate(self) -> None:
        # same errors, same order of checks as reference `_get_disp_fn` :182-203 and the dispatch :226-236
        if self.delta_x_method not in _DISP:
            raise ValueError(f"Unknown 'delta_x_method' value: {self.delta_x_method}")
        if self.covariance_method not in ("low_memory", "array", "batched"):
            raise ValueError(f"Unknown covariance_method: {self.covariance_method}")
        if self.covariance_method == "batched" and self.batch_size is None:
            raise ValueError("batch_size must be set if using 'batched' method")

    # -- device build --------------------------------------------------------------------------
    def _prepare_device(self) -> None:
        L = lib()
        N = int(len(self.Lx_compressed))
        self._N = N
        ws_bytes = ctypes.c_size_t()
        check(L.gmg_ellipse_ws_bytes(N, ctypes.byref(ws_bytes)), "gmg_ellipse_ws_bytes")
        self._ws = torch.empty(ws_bytes.value, dtype=torch.uint8, device=dev.require_cuda())
        ins = [dev.to_device(a) for a in (self.Lx_compressed, self.Ly_compressed, self.theta_compressed,
                                          self.stdev_compressed, self.lat_grid_compressed_rad,
                                          self.lon_grid_compressed_rad)]
        check(L.gmg_ellipse_prepare_f64(*[t.data_ptr() for t in ins], N, float(self.v),
                                        _DISP[self.delta_x_method], self._ws.data_ptr(), ws_bytes.value,
                                        dev.stream_ptr()), "gmg_ellipse_prepare_f64")
        self._inputs_device = ins  # keep alive until the stream has consumed them

    def _calculate_covariance(self) -> None:
        self._validate()
        self._prepare_device()
        N = self._N
        row0, nrows = (0, N) if self._rows is None else (int(self._rows[0]), int(self._rows[1]))
        if row0 < 0 or nrows <= 0 or row0 + nrows > N:
            raise ValueError(f"rows={self._rows} outside the {N} unmasked grid points")
        self.row0, self.nrows = row0, nrows
        f32 = self.precision == np.float32
        pb = self._peer_blocks
        if pb is not None:
            # every pair once across the box: transposed tiles go straight into the owning peer's rows (NVLink)
            if pb.N != N or pb.local.dtype != (torch.float32 if f32 else torch.float64):
                raise ValueError("peer_blocks does not match the grid size / precision of this build")
            pb.barrier()                      # peers may still be reading the previous contents
            check(lib().gmg_ellipse_cov_p2p_f64(self._ws.data_ptr(), N, float(self.v),
                                                -1.0 if self.max_dist is None else float(self.max_dist),
                                                _MATERN[self.matern_mode], pb.world, pb.rank, pb.block_row0, pb.ptrs,
                                                N, int(f32), dev.stream_ptr()), "gmg_ellipse_cov_p2p_f64")
            pb.barrier()                      # this rank's block is complete once ALL ranks have finished
            self.cov_ns_device = pb.local
            self._cov_ns_host = None
            return
        out = torch.empty((nrows, N), dtype=torch.float32 if f32 else torch.float64, device=self._ws.device)
        check(lib().gmg_ellipse_cov_f64(self._ws.data_ptr(), N, row0, nrows, float(self.v),
                                        -1.0 if self.max_dist is None else float(self.max_dist),
                                        _MATERN[self.matern_mode], out.data_ptr(), N, int(f32),
                                        dev.stream_ptr()), "gmg_ellipse_cov_f64")
        self.cov_ns_device = out
        self._cov_ns_host = None

    # -- the reference's three execution strategies (`:250-403`) -------------------------------------
    # Each fills `cov_ns` with the off-diagonal covariances and a ZERO diagonal (the reference adds diag(stdev^2)
    # afterwards, `:246`); here all three are the same tile kernel - there is no pair-index array to batch and no
    # Python loop to fall back to - followed by a reset of the diagonal.
    def _off_diagonal_build(self) -> None:
        self._calculate_covariance()
        self.cov_ns_device.diagonal(offset=self.row0).zero_()      # (r, row0 + r) for a row block

    def calculate_covariance_array(self) -> None:
        """Reference `:250-279` (all pairs at once)."""
        self._off_diagonal_build()

    def calculate_covariance_loop(self) -> None:
        """Reference `:281-354` (scalar loop over pairs)."""
        self._off_diagonal_build()

    def calculate_covariance_batched(self) -> None:
        """Reference `:356-403` (`batch_size` pairs at a time)."""
        if self.batch_size is None:
            raise ValueError("batch_size must be set if using 'batched' method")   # :375-376
        self._off_diagonal_build()

    def _get_disp_fn(self) -> None:
        """Reference `:182-203`: `disp_fn` is the displacement helper the chosen strategy would call."""
        single = self.covariance_method == "low_memory"
        if self.delta_x_method == "Modified_Met_Office":
            self.disp_fn = _mod_mo_disp_single if single else _mod_mo_disp_multi
        elif self.delta_x_method == "Met_Office":
            self.disp_fn = _mo_disp_single if single else _mo_disp_multi
        else:
            raise ValueError(f"Unknown 'delta_x_method' value: {self.delta_x_method}")

    # -- results ---------------------------------------------------------------------------------
    @property
    def cov_ns(self) -> np.ndarray:
        """The covariance as a numpy array of dtype `precision` (copied from the GPU on first use)."""
        if self._cov_ns_host is None:
            self._cov_ns_host = dev.to_host(self.cov_ns_device)
        return self._cov_ns_host

    @cov_ns.setter
    def cov_ns(self, value) -> None:
        if isinstance(value, torch.Tensor):
            self.cov_ns_device, self._cov_ns_host = value, None
        else:
            self._cov_ns_host = np.asarray(value)
            self.cov_ns_device = None

    # reference computes these at :219-224 and keeps them as attributes; here they are evaluated on demand by the
    # same device helpers the module exports (`gmg_sigma_rot_f64`, `gmg_det22_f64`)
    @property
    def sigmas(self) -> np.ndarray:
        return _sigma_rot_func_multi(self.Lx_compressed, self.Ly_compressed, self.theta_compressed).astype(self.precision)

    @property
    def sqrt_dets(self) -> np.ndarray:
        return np.sqrt(_det_22_multi(self.sigmas))

    @property
    def gamma_v_term(self) -> float:
        import math

        return math.gamma(self.v) * (2 ** (self.v - 1))

    @property
    def sqrt_v_term(self) -> float:
        return float(np.sqrt(self.v) * 2)

    def c_ij_anisotropic_array(self, i_s, j_s) -> np.ndarray:
        """Covariances between pairs of ellipses (reference `:405-468`), evaluated on the GPU."""
        i_t = dev.to_device(np.asarray(i_s), torch.int64)
        j_t = dev.to_device(np.asarray(j_s), torch.int64)
        out = torch.empty(i_t.numel(), dtype=torch.float64, device=i_t.device)
        check(lib().gmg_ellipse_pairs_f64(self._ws.data_ptr(), self._N, i_t.data_ptr(), j_t.data_ptr(),
                                          i_t.numel(), float(self.v), _MATERN[self.matern_mode], out.data_ptr(), dev.stream_ptr()),
              "gmg_ellipse_pairs_f64")
        return out.cpu().numpy().astype(self.precision)

    def calculate_cor(self) -> None:
        """Correlation matrix from the covariance matrix (reference `:470-473`), on the device; `cor_ns` is copied to
        the host on first use, `cor_ns_device` is the CUDA tensor."""
        if self.cov_ns_device is None:
            self.cov_ns_device = dev.to_device(self._cov_ns_host,
                                               torch.float32 if self.precision == np.float32 else torch.float64)
        if self.cov_ns_device.shape[0] != self.cov_ns_device.shape[1]:
            raise ValueError("calculate_cor needs the full matrix (this builder holds a row block)")
        self.cor_ns_device = cov_2_cor(self.cov_ns_device)
        self._cor_ns_host = None

    @property
    def cor_ns(self) -> np.ndarray:
        if getattr(self, "cor_ns_device", None) is None:
            raise AttributeError("cor_ns")                       # reference: the attribute exists after calculate_cor()
        if self._cor_ns_host is None:
            self._cor_ns_host = dev.to_host(self.cor_ns_device)
        return self._cor_ns_host

    def uncompress_cov(self, diag_fill_value: Any = np.nan, fill_value: Any = np.nan) -> None:
        """Scatter the compressed matrix into full-grid size (reference `:475-509`) on the device; overwrites
        `cov_ns` / `cov_ns_device`."""
        if self.cov_ns_device is None:
            self.cov_ns_device = dev.to_device(self._cov_ns_host,
                                               torch.float32 if self.precision == np.float32 else torch.float64)
        cov = self.cov_ns_device
        flat = np.asarray(self.data_mask).flatten()
        if cov.shape[0] != cov.shape[1] or not np.sum(~flat) == cov.shape[0]:
            raise ValueError("Data mask and coordinates cannot be aligned")
        M = flat.size
        inv = np.full(M, -1, dtype=np.int32)
        inv[~flat] = np.arange(cov.shape[0], dtype=np.int32)
        inv_t = dev.to_device(inv, torch.int32)
        out = torch.empty((M, M), dtype=cov.dtype, device=cov.device)
        check(lib().gmg_uncompress(cov.data_ptr(), cov.stride(0), inv_t.data_ptr(), M,
                                   int(cov.dtype == torch.float32), float(fill_value), float(diag_fill_value),
                                   out.data_ptr(), M, dev.stream_ptr()), "gmg_uncompress")
        torch.cuda.current_stream().synchronize()                # inv_t dies with this frame
        self.cov_ns_device, self._cov_ns_host = out, None


# -- module-level geometry helpers (reference `:512-654`), evaluated on the GPU ------------------------------------
def _dtype_of(*arrays):
    """float32 only if every array input is float32 (numpy's promotion for the reference expressions)."""
    kinds = [np.asarray(a).dtype for a in arrays]
    return np.float32 if all(k == np.float32 for k in kinds) else np.float64


def _dev_vec(a) -> torch.Tensor:
    return dev.to_device(np.atleast_1d(np.asarray(a, dtype=np.float64)).reshape(-1), torch.float64)


def _sigma_rot_func_multi(Lx: np.ndarray, Ly: np.ndarray, theta: np.ndarray) -> np.ndarray:
    """Flattened Sigma matrices (n x 4) for Lx, Ly, theta triplets; reference `:512-533` (`gmg_sigma_rot_f64`)."""
    lx, ly, th = _dev_vec(Lx), _dev_vec(Ly), _dev_vec(theta)
    if not lx.numel() == ly.numel() == th.numel():
        raise ValueError("Lx, Ly and theta must have the same length")
    out = torch.empty((lx.numel(), 4), dtype=torch.float64, device=lx.device)
    check(lib().gmg_sigma_rot_f64(lx.data_ptr(), ly.data_ptr(), th.data_ptr(), lx.numel(), out.data_ptr(),
                                  dev.stream_ptr()), "gmg_sigma_rot_f64")
    return out.cpu().numpy().astype(_dtype_of(Lx, Ly, theta), copy=False)


def _det_22_multi(mats: np.ndarray) -> np.ndarray:
    """Determinants of an array of flattened 2 x 2 matrices; reference `:543-547` (`gmg_det22_f64`)."""
    m = np.asarray(mats)
    if m.ndim != 2 or m.shape[1] != 4:
        raise ValueError("mats must have shape (n, 4)")
    t = dev.to_device(m, torch.float64)
    out = torch.empty(t.shape[0], dtype=torch.float64, device=t.device)
    check(lib().gmg_det22_f64(t.data_ptr(), t.shape[0], out.data_ptr(), dev.stream_ptr()), "gmg_det22_f64")
    return out.cpu().numpy().astype(_dtype_of(m), copy=False)


def _det_22_single(mat: np.ndarray):
    """Determinant of one flattened 2 x 2 matrix; reference `:536-540`."""
    return _det_22_multi(np.asarray(mat).reshape(1, 4))[0]


def _pairs(lat0, lon0, lat1, lon1):
    ts = [_dev_vec(a) for a in (lat0, lon0, lat1, lon1)]
    if len({t.numel() for t in ts}) != 1:
        raise ValueError("lat0, lon0, lat1, lon1 must have the same length")
    return ts


def _haversine_multi(lat0: np.ndarray, lon0: np.ndarray, lat1: np.ndarray, lon1: np.ndarray) -> np.ndarray:
    """Haversine distance (km) for a list of pairs of points in radians; reference `:572-588`
    (`gmg_pair_haversine_f64`)."""
    ts = _pairs(lat0, lon0, lat1, lon1)
    out = torch.empty(ts[0].numel(), dtype=torch.float64, device=ts[0].device)
    check(lib().gmg_pair_haversine_f64(*[t.data_ptr() for t in ts], out.numel(), RADIUS_OF_EARTH_KM, out.data_ptr(),
                                       dev.stream_ptr()), "gmg_pair_haversine_f64")
    return out.cpu().numpy().astype(_dtype_of(lat0, lon0, lat1, lon1), copy=False)


def _haversine_single(lat0: float, lon0: float, lat1: float, lon1: float) -> float:
    """Haversine distance between a pair of points; reference `:550-569` (exactly 0 below 1e-6 rad)."""
    if abs(lon0 - lon1) < 1e-6 and abs(lat0 - lat1) < 1e-6:
        return 0
    return float(_haversine_multi([lat0], [lon0], [lat1], [lon1])[0])


def _disp(lat0, lon0, lat1, lon1, modified: int):
    ts = _pairs(lat0, lon0, lat1, lon1)
    out = torch.empty((2, ts[0].numel()), dtype=torch.float64, device=ts[0].device)
    check(lib().gmg_pair_displacement_f64(*[t.data_ptr() for t in ts], out.shape[1], modified, RADIUS_OF_EARTH_KM,
                                          out[0].data_ptr(), out[1].data_ptr(), dev.stream_ptr()),
          "gmg_pair_displacement_f64")
    o = out.cpu().numpy().astype(_dtype_of(lat0, lon0, lat1, lon1), copy=False)
    return o[0], o[1]


def _mod_mo_disp_multi(lat0, lon0, lat1, lon1) -> tuple[np.ndarray, np.ndarray]:
    """Modified Met Office displacements (dy, dx) in km between pairs of points; reference `:624-639`."""
    return _disp(lat0, lon0, lat1, lon1, 1)


def _mo_disp_multi(lat0, lon0, lat1, lon1) -> tuple[np.ndarray, np.ndarray]:
    """Met Office displacements (dy, dx) in km between pairs of points; reference `:642-654`."""
    return _disp(lat0, lon0, lat1, lon1, 0)


def _mod_mo_disp_single(lat0: float, lon0: float, lat1: float, lon1: float) -> tuple[float, float]:
    """Reference `:591-606`."""
    dy, dx = _disp([lat0], [lon0], [lat1], [lon1], 1)
    return float(dy[0]), float(dx[0])


def _mo_disp_single(lat0: float, lon0: float, lat1: float, lon1: float) -> tuple[float, float]:
    """Reference `:609-621`."""
    dy, dx = _disp([lat0], [lon0], [lat1], [lon1], 0)
    return float(dy[0]), float(dx[0])
