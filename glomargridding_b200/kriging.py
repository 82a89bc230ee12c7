"""`SimpleKriging` / `OrdinaryKriging`: drop-ins for `glomar_gridding.kriging` on a B200.

Same classes, constructor signatures, methods, attributes, warnings and errors as the reference
(`glomar_gridding/kriging.py:41-829`, deprecated functional API `:935-1118`).  The arithmetic runs on the
GPU through `libglomar_b200.so` and there is no CPU fallback:

  gather S = C[idx, idx] + E, C_cross = C[idx, :]            (reference :337-342, :571-583)
  S = L L^T      blocked Cholesky, trailing update on FP64 DMMA tiles      (reference: LU inside np.linalg.solve)
  T = L^-T L^-1 [C_cross | 1]     blocked substitution, every step a DMMA GEMM
  field / variance / constraint mask from ONE pass over T and C_cross      (reference: M x M temporaries)

Differences from the reference, by design:

* The reference solves with LU (`np.linalg.solve`); this implementation uses Cholesky and, when
  `C[idx, idx] + E` turns out not to be positive definite (the default ellipse covariance is
  indefinite, and the reference's GeoStats.jl known-answer test feeds a variogram matrix), warns and falls
  back to a device LU with partial pivoting - same results as the reference, without the DMMA speed.
* Ordinary kriging is solved through the Schur complement of the Lagrange-extended system instead of an
  LU of the (n+1) x (n+1) bordered matrix; the results agree to ~1e-12 (max-norm relative).
* `kriging_weights` (M x n, or M x (n+1)) is materialised on first access; `solve`, `get_uncertainty` and
  `constraint_mask` never need it.  `covariance` may also be a CUDA `torch.Tensor` (e.g.
  `EllipseCovarianceBuilder.cov_ns_device`), which avoids a host round trip of the matrix.
* Results are computed in float64 and cast to the dtype the reference would return.
"""

from __future__ import annotations

from abc import ABC, abstractmethod
from typing import Literal
from warnings import warn

import numpy as np
import torch

from . import _dense as dn
from . import _device as dev
from ._lib import check, lib
from .utils import adjust_small_negative, get_spatial_mean, intersect_mtlb


def _as_device_cov(covariance) -> torch.Tensor:
    """float64 CUDA tensor of the covariance (no copy if it already is one)."""
    if isinstance(covariance, torch.Tensor):
        t = covariance.to(device=dev.require_cuda(), dtype=torch.float64)
        return t if t.stride(-1) == 1 else t.contiguous()
    return dev.to_device(np.asarray(covariance), torch.float64)


def _cov_dtype(covariance):
    if isinstance(covariance, torch.Tensor):
        return np.float32 if covariance.dtype == torch.float32 else np.float64
    dt = np.asarray(covariance).dtype
    return dt if np.issubdtype(dt, np.floating) else np.float64


class Kriging(ABC):
    """Base class (reference `kriging.py:41-279`).  Use `SimpleKriging` or `OrdinaryKriging`.

    Parameters
    ----------
    covariance : numpy.ndarray | torch.Tensor
        The spatial covariance matrix (M x M).
    idx : numpy.ndarray
        1-d flat ("C" order) grid indices of the observations, each value at most once.
    obs : numpy.ndarray
        The observation values (one per observed grid box).
    error_cov : numpy.ndarray | None
        Error covariance between observation grid points (n x n) or on the full grid (M x M).
    """

    def __init__(self, covariance, idx: np.ndarray, obs: np.ndarray, error_cov: np.ndarray | None = None) -> None:
        if not hasattr(self, "method"):
            raise NotImplementedError(
                "Do not use the generic class directly, " + "use SimpleKriging or OrdinaryKriging"
            )
        self.covariance = covariance
        self.idx = idx
        self.obs = obs
        self.error_cov = error_cov
        self._batch: dn.KrigeBatch | None = None     # device state of the last factorisation
        self._user_weights = None                     # weights supplied through set_kriging_weights
        self.subset_error_covariance()

    # -- reference `subset_error_covariance`, kriging.py:89-121 (host side, O(n^2) bookkeeping) ----------
    def subset_error_covariance(self) -> None:
        """Adjust the error covariance matrix to the grid size, check for nans or zeros on the diagonal."""
        if self.error_cov is not None:
            if self.error_cov.shape[0] != len(self.idx):
                self.error_cov = self.error_cov[self.idx[:, None], self.idx[None, :]]
            if (
                mismatch := np.logical_or(np.isnan(self.error_cov.diagonal()), self.error_cov.diagonal() == 0)
            ).any():
                idx_keep = np.where(np.logical_not(mismatch))[0]
                drop_idx = self.idx[mismatch].tolist()
                msg = (
                    "Have nans or zeros on the error covariance diagonal. "
                    + "At positions "
                    + " ,".join(map(str, drop_idx))
                    + ". Filtering input accordingly"
                )
                warn(msg)
                self.idx = self.idx[idx_keep]
                self.obs = self.obs[idx_keep]
                self.error_cov = self.error_cov[idx_keep[:, None], idx_keep[None, :]]
            # NaN -> 0 (reference :120).  A sum is NaN whenever any entry is (or +inf meets -inf), and costs a tenth of
            # the boolean-mask assignment, which is therefore only run when there is something to replace
            # The batched driver skips this O(n^2) host scan altogether: the gather kernel reads NaN entries of E as
            # 0 (`csrc/kriging.cu`), and nobody sees the per-problem `error_cov` attribute there.
            ec = self.error_cov
            if getattr(self, "_nan_to_zero_on_device", False) and isinstance(ec, np.ndarray) and ec.dtype.kind == "f":
                return None
            if not (isinstance(ec, np.ndarray) and ec.dtype.kind == "f") or np.isnan(ec.sum()):
                ec[np.isnan(ec)] = 0.0
        return None

    # -- weights attribute: computed lazily from the device factorisation --------------------------------
    def _has_weights(self) -> bool:
        """What `hasattr(self, "kriging_weights")` means in the reference - WITHOUT evaluating the property, whose
        getter runs the backward triangular solve and copies the M x n weights matrix to the host (1.7 of the 7.7 ms
        of a 5-degree class call, for methods that never look at the weights)."""
        return self._user_weights is not None or self._batch is not None

    @property
    def kriging_weights(self) -> np.ndarray:
        if self._user_weights is not None:
            return self._user_weights
        if self._batch is None:
            raise AttributeError("kriging_weights")       # keeps `hasattr(self, "kriging_weights")` semantics
        if getattr(self, "_weights_host", None) is None:
            W = self._batch.weights(0, ordinary=self.method == "ordinary")
            self._weights_host = dev.to_host(W).astype(_cov_dtype(self.covariance), copy=False)
        return self._weights_host

    @kriging_weights.setter
    def kriging_weights(self, value) -> None:
        self.set_kriging_weights(value)

    @kriging_weights.deleter
    def kriging_weights(self) -> None:
        self._user_weights = None
        self._batch = None
        self._weights_host = None

    def set_kriging_weights(self, kriging_weights: np.ndarray) -> None:
        """Set Kriging Weights (reference `:123-135`)."""
        self._user_weights = kriging_weights
        self._batch = None
        self._weights_host = None
        return None

    @abstractmethod
    def get_kriging_weights(self) -> None:
        raise NotImplementedError("`get_kriging_weights` not implemented for default class")

    @abstractmethod
    def kriging_weights_from_inverse(self, inv: np.ndarray) -> None:
        raise NotImplementedError("`kriging_weights_from_inverse` not implemented for default class")

    @abstractmethod
    def solve(self) -> np.ndarray:
        raise NotImplementedError("`solve` not implemented for default class")

    @abstractmethod
    def get_uncertainty(self) -> np.ndarray:
        raise NotImplementedError("`get_uncertainty` not implemented for default class")

    @abstractmethod
    def constraint_mask(self) -> np.ndarray:
        raise NotImplementedError("`constraint_mask` not implemented for default class")

    # -- shared device plumbing ---------------------------------------------------------------------------
    def _cov_device(self) -> torch.Tensor:
        # keyed on the identity of `self.covariance`: assigning a new matrix to the attribute drops the device copy
        if getattr(self, "_cov_dev", None) is None or getattr(self, "_cov_src", None) is not self.covariance:
            self._cov_dev = _as_device_cov(self.covariance)
            self._cov_src = self.covariance
        return self._cov_dev

    def _problem(self):
        E = None if self.error_cov is None else np.asarray(self.error_cov)     # cast to float64 while packing
        return (np.asarray(self.idx, dtype=np.int64), np.asarray(self.obs, dtype=np.float64), E)

    def _factorise(self, ordinary: bool) -> dn.KrigeBatch:
        """gather -> Cholesky -> solves -> apply, once; results cached until the weights are replaced."""
        batch = dn.KrigeBatch(self._cov_device(), [self._problem()], ordinary=ordinary).run()
        self._out_host = batch.out[0].cpu().numpy()
        self._obs_at_factor = np.array(self.obs, dtype=np.float64, copy=True)
        return batch

    def _field_from_batch(self) -> np.ndarray:
        """Kriged field of the kept factorisation for the CURRENT `self.obs`.

        The weights do not depend on the observations, and the reference evaluates `kriging_weights @ self.obs` when
        `solve` is called (`kriging.py:417`, `:677-679`), so `k.obs = other_member; k.solve()` (ensemble members on the
        same cells) must see the new values.  The row cached by `_factorise` belongs to the observations of that
        moment; for different ones the field is recomputed from the factorisation - one extra right-hand side,
        z = S^-1 y, and one GEMM row:  C_cross^T z  [- alpha 1^T z  for ordinary kriging]."""
        obs = np.asarray(self.obs, dtype=np.float64)
        if obs.shape == self._obs_at_factor.shape and np.array_equal(obs, self._obs_at_factor, equal_nan=True):
            return self._out_host[dn.OUT_FIELD]
        b = self._batch
        if obs.shape != (b.ns[0],):
            raise ValueError(f"obs of shape {obs.shape} does not match the {b.ns[0]} observations of the weights")
        cov = self._cov_device()
        M = cov.shape[0]
        y = dev.to_device(obs, torch.float64)
        ordinary = b.ordinary and self.method != "simple"
        if 0 in b.lu_fallback:
            o = dn.apply_reduce(cov, b.Cx(0), b.T(0), y, ordinary=ordinary)
            return o[dn.OUT_FIELD].cpu().numpy()
        z = dn.padded(b.solve_vectors(0, y[:, None]))              # (n, 2): S^-1 y in column 0
        field = dn.gemm(z[:, :1], b.Cx(0), transa=True)[0, :M]
        if ordinary:
            field = field - b.out[0, dn.OUT_ALPHA] * z[:, 0].sum()
        return field.cpu().numpy()

    def _apply_user_weights(self, W: np.ndarray, n: int):
        """(W[:, :n] @ obs, rowdot(W[:, :n], C_cross^T)) on the device for externally supplied weights
        (`set_kriging_weights`, `kriging_weights_from_inverse`): T = W[:, :n]^T runs through the same apply pass."""
        cov = self._cov_device()
        M = cov.shape[0]
        W = np.asarray(W)
        if W.shape[0] != M or W.shape[1] < n:
            raise ValueError(f"kriging weights of shape {W.shape} do not match covariance ({M}) and idx ({n})")
        Tt = dev.to_device(np.ascontiguousarray(W[:, :n].T), torch.float64)
        return self._apply_T(Tt)

    def _apply_T(self, Tt: torch.Tensor):
        """Column reductions of an (n x M) transposed-weights tensor against obs and C_cross."""
        cov = self._cov_device()
        M = cov.shape[0]
        idx_t = dev.to_device(np.asarray(self.idx), torch.int64)
        Cx = dn.padded(cov[idx_t, :M])                  # plumbing copy of the C_cross rows
        y = dev.to_device(np.asarray(self.obs, dtype=np.float64), torch.float64)
        o = dn.apply_reduce(cov, Cx, dn.padded(Tt[:, :M]), y, ordinary=False).cpu().numpy()
        return o[dn.OUT_FIELD], o[dn.OUT_ROWDOT], o[dn.OUT_MASK]

    def _diag_cov(self) -> np.ndarray:
        return torch.diagonal(self._cov_device()).cpu().numpy()

    def _out_dtype(self):
        return np.result_type(_cov_dtype(self.covariance), np.asarray(self.obs).dtype)


def _finish_uncertainty(u2: np.ndarray, dtype) -> np.ndarray:
    """`adjust_small_negative` -> sqrt -> NaN -> 0 (reference `kriging.py:436-438`, `:702-704`)."""
    u2 = u2.astype(dtype, copy=False)
    if (u2 < 0).any():                         # the helper only acts on (and warns about) negative values
        u2 = adjust_small_negative(u2)
    with np.errstate(invalid="ignore"):
        uncert = np.sqrt(u2)
    nan = np.isnan(uncert)
    if nan.any():
        uncert[nan] = 0.0
    return uncert


class SimpleKriging(Kriging):
    """Simple kriging with a known constant mean (reference `kriging.py:282-491`)."""

    method: str = "simple"

    def get_kriging_weights(self) -> None:
        """(C_obs + E)^-1 C_cross by Cholesky on the device; sets `kriging_weights` (reference `:318-345`)."""
        self._user_weights = None
        self._weights_host = None
        self._batch = self._factorise(ordinary=False)
        return None

    def kriging_weights_from_inverse(self, inv: np.ndarray) -> None:
        """Weights from a pre-computed (C_obs + E)^-1 (reference `:347-378`): one DMMA GEMM."""
        if len(self.idx) != inv.shape[0]:
            raise ValueError("inv must be square with side length == len(idx)")
        cov = self._cov_device()
        idx_t = dev.to_device(np.asarray(self.idx), torch.int64)
        M = cov.shape[0]
        Cx = torch.zeros((len(self.idx), M + (M & 1)), dtype=torch.float64, device=cov.device)
        Cx[:, :M] = cov[idx_t, :M]
        n = inv.shape[0]
        inv_t = torch.zeros((n, n + (n & 1)), dtype=torch.float64, device=cov.device)
        inv_t[:, :n] = dev.to_device(np.asarray(inv), torch.float64)
        T = dn.gemm(inv_t[:, :n], Cx)                  # (n, M'): (inv @ C_cross)
        W = T[:, :M].t().contiguous()
        self.set_kriging_weights(dev.to_host(W).astype(np.result_type(np.asarray(inv).dtype,
                                                                      _cov_dtype(self.covariance)), copy=False))

    def solve(self, mean: np.ndarray | float = 0.0) -> np.ndarray:
        """W y + mean (reference `:380-417`)."""
        if not self._has_weights():
            self.get_kriging_weights()
        if self._batch is not None:
            field = self._field_from_batch()
        else:
            field, _, _ = self._apply_user_weights(self._user_weights, len(self.idx))
        return field.astype(self._out_dtype(), copy=False) + mean

    def get_uncertainty(self) -> np.ndarray:
        """sqrt(diag(C - W C_cross)) with small negatives clamped (reference `:419-439`)."""
        if not self._has_weights():
            raise KeyError("Please compute Kriging Weights first")
        if self._batch is not None:
            u2 = self._out_host[dn.OUT_U2]
        else:
            _, rowdot, _ = self._apply_user_weights(self._user_weights, len(self.idx))
            u2 = self._diag_cov() - rowdot
        return _finish_uncertainty(u2, _cov_dtype(self.covariance))

    def constraint_mask(self) -> np.ndarray:
        """diag(C[:, idx] W^T) / diag(C) (reference `:441-491`, Morice et al. 2021 eq. A14)."""
        if not self._has_weights():
            raise KeyError("Please compute Kriging Weights first")
        if self._batch is not None:
            mask = self._out_host[dn.OUT_MASK]
        else:
            _, _, mask = self._apply_user_weights(self._user_weights, len(self.idx))
        return mask.astype(_cov_dtype(self.covariance), copy=False)


class OrdinaryKriging(Kriging):
    """Ordinary kriging with an unknown constant mean (reference `kriging.py:494-812`)."""

    method: str = "ordinary"

    def get_kriging_weights(self) -> None:
        """Lagrange-extended weights (M x (n+1), multiplier in the last column; reference `:540-588`) through
        the Schur complement: o = S^-1 1, alpha = (o^T C_cross - 1) / (1^T o), W = (S^-1 C_cross - o alpha^T)^T."""
        self._user_weights = None
        self._weights_host = None
        self._batch = self._factorise(ordinary=True)
        return None

    def kriging_weights_from_inverse(self, inv: np.ndarray) -> None:
        """Weights from a pre-computed inverse of the EXTENDED system (reference `:590-636`)."""
        if len(self.idx) != inv.shape[0] - 1:
            raise ValueError("inv must be square with side length == len(idx)")
        cov = self._cov_device()
        idx_t = dev.to_device(np.asarray(self.idx), torch.int64)
        M, n1 = cov.shape[0], inv.shape[0]
        Cx = torch.zeros((n1, M + (M & 1)), dtype=torch.float64, device=cov.device)
        Cx[: n1 - 1, :M] = cov[idx_t, :M]
        Cx[n1 - 1, :M] = 1.0                           # Lagrange row (:630-633)
        inv_t = torch.zeros((n1, n1 + (n1 & 1)), dtype=torch.float64, device=cov.device)
        inv_t[:, :n1] = dev.to_device(np.asarray(inv), torch.float64)
        T = dn.gemm(inv_t[:, :n1], Cx)
        W = T[:, :M].t().contiguous()
        self.set_kriging_weights(dev.to_host(W).astype(_cov_dtype(self.covariance), copy=False))

    def solve(self) -> np.ndarray:
        """W_ext [y; 0] (reference `:638-679`)."""
        if not self._has_weights():
            self.get_kriging_weights()
        if self._batch is not None:
            field = self._field_from_batch()
        else:
            field, _, _ = self._apply_user_weights(self._user_weights, len(self.idx))
        return field.astype(self._out_dtype(), copy=False)

    def get_uncertainty(self) -> np.ndarray:
        """sqrt(diag(C - W_ext C_cross_ext) - alpha) (reference `:681-706`; the multiplier enters twice, H4)."""
        if not self._has_weights():
            raise KeyError("Please compute Kriging Weights first")
        if self._batch is not None:
            u2 = self._out_host[dn.OUT_U2]
        else:
            n = len(self.idx)
            _, rowdot, _ = self._apply_user_weights(self._user_weights, n)
            alpha = np.asarray(self._user_weights)[:, n].astype(np.float64)
            u2 = self._diag_cov() - rowdot - alpha - alpha
        return _finish_uncertainty(u2, _cov_dtype(self.covariance))

    def constraint_mask(self, simple_kriging_weights: np.ndarray | None = None) -> np.ndarray:
        """Constraint mask from the SIMPLE kriging weights (reference `:708-776`); they are taken from the
        device factorisation when it exists, otherwise solved for."""
        if simple_kriging_weights is not None:
            _, _, mask = self._apply_user_weights(simple_kriging_weights, len(self.idx))
        elif self._batch is not None:
            mask = self._out_host[dn.OUT_MASK]
        else:
            batch = dn.KrigeBatch(self._cov_device(), [self._problem()], ordinary=False).run()
            mask = batch.out[0, dn.OUT_MASK].cpu().numpy()
        return mask.astype(_cov_dtype(self.covariance), copy=False)

    def extended_inverse(self, simple_inv: np.ndarray) -> np.ndarray:
        """Inverse of the Lagrange-extended matrix from (C_obs + E)^-1 (reference `:778-812`)."""
        return _extended_inverse(simple_inv).astype(_cov_dtype(self.covariance))


def _extended_inverse(simple_inv: np.ndarray) -> np.ndarray:
    """Bordered-matrix inverse (Guttman 1946; reference `kriging.py:815-829`) on the device."""
    simple_inv = np.asarray(simple_inv)
    if len(simple_inv.shape) != 2:
        raise ValueError("S must be a matrix")
    n = simple_inv.shape[0]
    inv_t = dev.to_device(simple_inv, torch.float64)
    e = torch.empty(n, dtype=torch.float64, device=inv_t.device)
    out = torch.empty((n + 1, n + 1), dtype=torch.float64, device=inv_t.device)
    check(lib().gmg_extended_inverse_f64(inv_t.data_ptr(), n, inv_t.stride(0), e.data_ptr(), out.data_ptr(),
                                         out.stride(0), dev.stream_ptr()), "gmg_extended_inverse_f64")
    return dev.to_host(out).astype(simple_inv.dtype if np.issubdtype(simple_inv.dtype, np.floating)
                                   else np.float64, copy=False)


# ---- observation preparation helpers (host bookkeeping; reference `kriging.py:832-932`) ---------------------
def prep_obs_for_kriging(
    unmask_idx: np.ndarray,
    unique_obs_idx: np.ndarray,
    weights: np.ndarray,
    obs: np.ndarray,
    remove_obs_mean: int = 0,
    obs_bias: np.ndarray | None = None,
    error_cov: np.ndarray | None = None,
) -> tuple[np.ndarray, np.ndarray]:
    """Indices of observed cells among the unmasked cells and the weighted gridded observations
    (reference `kriging.py:832-904`)."""
    obs_idx = get_unmasked_obs_indices(unmask_idx, unique_obs_idx)
    if obs_bias is not None:
        print("With bias")
        grid_obs = weights @ (obs - obs_bias)
    else:
        grid_obs = weights @ obs
    grid_obs = np.squeeze(grid_obs) if grid_obs.ndim > 1 else grid_obs
    match remove_obs_mean:
        case 0:
            grid_obs_av = None
        case 1:
            grid_obs_av = np.ma.average(grid_obs)
        case 2:
            grid_obs_av = np.ma.median(grid_obs)
        case 3:
            if error_cov is None:
                raise ValueError("'remove_obs_mean = 3 requires error covariance")
            grid_obs_av = get_spatial_mean(grid_obs, error_cov)
        case _:
            raise ValueError("Unknown 'remove_obs_mean' value")
    if grid_obs_av is not None:
        grid_obs = grid_obs - grid_obs_av
    return obs_idx, grid_obs


def get_unmasked_obs_indices(unmask_idx: np.ndarray, unique_obs_idx: np.ndarray) -> np.ndarray:
    """Position of each observed grid cell inside the vector of unmasked cells (reference `:907-932`)."""
    unmask_idx = np.squeeze(unmask_idx) if unmask_idx.ndim > 1 else unmask_idx
    _, obs_idx, _ = intersect_mtlb(unmask_idx, unique_obs_idx)
    return obs_idx.astype(int)


# ---- deprecated functional API (reference `kriging.py:935-1118`), kept because the reference's own tests use
# it as the cross-check of the classes; the arithmetic goes through the same device kernels ----------------------
def _functional(obs_obs_cov, obs_grid_cov, grid_obs, interp_cov, ordinary: bool) -> np.ndarray:
    """(5, M) result rows for explicit (C_obs + E, C_cross, y, C): Cholesky solve + one apply pass."""
    Cx = dn.padded(np.asarray(obs_grid_cov))
    n, M = np.asarray(obs_grid_cov).shape
    T = dn.padded(np.asarray(obs_grid_cov), cols=M + 1)
    if ordinary:
        T[:, M] = 1.0
    S = dn.padded(np.asarray(obs_obs_cov))
    dn.cholesky_solve_inplace(S, T, "obs_obs_cov")   # the padding column of simple kriging is zero: harmless
    y = dev.to_device(np.zeros(n) if grid_obs is None else np.asarray(grid_obs, dtype=np.float64), torch.float64)
    cov = interp_cov if isinstance(interp_cov, torch.Tensor) else dev.to_device(np.asarray(interp_cov), torch.float64)
    return dn.apply_reduce(cov, Cx, T, y, ordinary).cpu().numpy()


def kriging_simple(
    obs_obs_cov: np.ndarray,
    obs_grid_cov: np.ndarray,
    grid_obs: np.ndarray,
    interp_cov: np.ndarray,
    mean: float | np.ndarray = 0.0,
) -> tuple[np.ndarray, np.ndarray]:
    """Deprecated functional simple kriging (reference `:935-989`): (field, uncertainty)."""
    warn(
        "kriging_simple is deprecated and will be removed in version v1.0.0, "
        + "use SimpleKriging class instead",
        DeprecationWarning,
    )
    o = _functional(obs_obs_cov, obs_grid_cov, grid_obs, interp_cov, ordinary=False)
    return o[dn.OUT_FIELD] + mean, _finish_uncertainty(o[dn.OUT_U2], np.float64)


def kriging_ordinary(
    obs_obs_cov: np.ndarray,
    obs_grid_cov: np.ndarray,
    grid_obs: np.ndarray,
    interp_cov: np.ndarray,
) -> tuple[np.ndarray, np.ndarray]:
    """Deprecated functional ordinary kriging (reference `:992-1052`): (field, uncertainty)."""
    warn(
        "kriging_ordinary is deprecated and will be removed in version v1.0.0, "
        + "use OrdinaryKriging class instead",
        DeprecationWarning,
    )
    o = _functional(obs_obs_cov, obs_grid_cov, grid_obs, interp_cov, ordinary=True)
    return o[dn.OUT_FIELD], _finish_uncertainty(o[dn.OUT_U2], np.float64)


def constraint_mask(
    obs_obs_cov: np.ndarray,
    obs_grid_cov: np.ndarray,
    interp_cov: np.ndarray,
) -> np.ndarray:
    """Deprecated functional constraint mask (reference `:1055-1118`)."""
    return _functional(obs_obs_cov, obs_grid_cov, None, interp_cov, ordinary=False)[dn.OUT_MASK]


KrigingMethod = Literal["simple", "ordinary"]
