"""`StochasticKriging` and `scipy_mv_normal_draw`: drop-ins for `glomar_gridding.stochastic` on a B200.

Two-stage kriging of Morice et al. (2021) as in the reference (`glomar_gridding/stochastic.py:32-392`): an
ordinary-kriging field from the observations plus a simple-kriging perturbation computed from a simulated
state and simulated observations.  Both sets of weights come from ONE device Cholesky of `C[idx, idx] + E`
(the reference inverts that matrix explicitly, `:153-161`); the Gaussian draws are `X = Z L^T` with
`L = chol(cov)` on DMMA tiles and `Z` from a device Philox generator.

Draw parity with the reference is impossible by construction: `np.random.multivariate_normal`
uses numpy's legacy MT19937 stream and an SVD factor.  What is kept: shapes, dtypes, the deterministic
sub-results, reproducibility under `np.random.seed` (the Philox key is drawn from numpy's global state), and the
distribution N(loc, cov).  For an INDEFINITE covariance numpy's SVD factor draws from V |Lambda| V^T (with a
RuntimeWarning); so does this implementation, through |C| = C sign(C) on the device; a merely singular matrix gets a
small diagonal lift (warning) before its Cholesky factorisation.
"""

from __future__ import annotations

import logging
from warnings import warn

import numpy as np
import torch

from . import _dense as dn
from . import _device as dev
from ._lib import check, lib
from .kriging import Kriging, _cov_dtype, _extended_inverse, _finish_uncertainty


class StochasticKriging(Kriging):
    """Stochastic (perturbed) kriging; reference `stochastic.py:32-392`.

    `error_cov` is mandatory.  After `solve()` the attributes `gridded_field`, `simulated_obs`,
    `simulated_grid` and `epsilon` are set as in the reference.
    """

    method: str = "stochastic"

    def __init__(self, covariance, idx: np.ndarray, obs: np.ndarray, error_cov: np.ndarray) -> None:
        if error_cov is None:
            raise ValueError("Error Covariance must be provided for StochasticKriging")
        self._user_simple_weights = None
        super().__init__(covariance=covariance, idx=idx, obs=obs, error_cov=error_cov)

    # -- simple kriging weights of the second stage --------------------------------------------------------
    def _has_simple_weights(self) -> bool:
        """`hasattr(self, "simple_kriging_weights")` without evaluating the property (see `Kriging._has_weights`)."""
        return self._user_simple_weights is not None or (self._batch is not None
                                                          and not getattr(self, "_simple_deleted", False))

    @property
    def simple_kriging_weights(self) -> np.ndarray:
        if self._user_simple_weights is not None:
            return self._user_simple_weights
        if self._batch is None or getattr(self, "_simple_deleted", False):
            raise AttributeError("simple_kriging_weights")
        if getattr(self, "_simple_weights_host", None) is None:
            W = self._batch.weights(0, ordinary=False)
            self._simple_weights_host = dev.to_host(W).astype(_cov_dtype(self.covariance), copy=False)
        return self._simple_weights_host

    @simple_kriging_weights.setter
    def simple_kriging_weights(self, value) -> None:
        self.set_simple_kriging_weights(value)

    @simple_kriging_weights.deleter
    def simple_kriging_weights(self) -> None:
        # the reference's tests delete and re-set this attribute (test/test_kriging.py:365-367); the device
        # factorisation is kept for the ordinary weights, only the simple weights become "unset"
        self._user_simple_weights = None
        self._simple_deleted = True

    def set_simple_kriging_weights(self, simple_kriging_weights: np.ndarray) -> None:
        """Set the simple kriging weights of the second stage (reference `:107-122`)."""
        self._user_simple_weights = simple_kriging_weights
        self._simple_deleted = False

    @property
    def kriging_weights(self) -> np.ndarray:
        if self._user_weights is not None:
            return self._user_weights
        if self._batch is None:
            raise AttributeError("kriging_weights")
        if getattr(self, "_weights_host", None) is None:
            W = self._batch.weights(0, ordinary=True)
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

    def get_kriging_weights(self) -> None:
        """Ordinary AND simple kriging weights from one factorisation (reference `:124-164`)."""
        self._user_weights = None
        self._user_simple_weights = None
        self._weights_host = None
        self._simple_weights_host = None
        self._simple_deleted = False
        self._batch = self._factorise(ordinary=True)
        return None

    def kriging_weights_from_inverse(self, inv: np.ndarray) -> None:
        """Both sets of weights from a pre-computed (C_obs + E)^-1 (reference `:166-218`)."""
        if len(self.idx) != inv.shape[0]:
            raise ValueError("inv must be square with side length == len(self.idx)")
        cov = self._cov_device()
        idx_t = dev.to_device(np.asarray(self.idx), torch.int64)
        M, n = cov.shape[0], inv.shape[0]
        Cx = torch.zeros((n + 1, M + (M & 1)), dtype=torch.float64, device=cov.device)
        Cx[:n, :M] = cov[idx_t, :M]
        Cx[n, :M] = 1.0                                 # Lagrange row (:213-215)
        dt = _cov_dtype(self.covariance)
        W_sk = dn.gemm(dn.padded(np.asarray(inv))[:, :n], Cx[:n])[:, :M].t().contiguous()
        ext = dn.padded(_extended_inverse(np.asarray(inv)))
        W_ok = dn.gemm(ext[:, : n + 1], Cx)[:, :M].t().contiguous()
        self.set_simple_kriging_weights(dev.to_host(W_sk).astype(dt, copy=False))
        self.set_kriging_weights(dev.to_host(W_ok).astype(dt, copy=False))

    def get_uncertainty(self) -> np.ndarray:
        """As `OrdinaryKriging.get_uncertainty` (reference `:220-243`)."""
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

    def constraint_mask(self) -> np.ndarray:
        """Constraint mask from the simple kriging weights (reference `:245-297`)."""
        if not self._has_simple_weights():
            raise KeyError("Please set kriging weights")
        if self._user_simple_weights is not None:
            _, _, mask = self._apply_user_weights(self._user_simple_weights, len(self.idx))
        else:
            mask = self._out_host[dn.OUT_MASK]
        return mask.astype(_cov_dtype(self.covariance), copy=False)

    def solve(self, simulated_state: np.ndarray | None = None) -> np.ndarray:
        """Ordinary-kriging field + simple-kriging perturbation (reference `:299-392`)."""
        if not self._has_weights():
            self.get_kriging_weights()
        if self.error_cov is None:
            raise ValueError("Error Covariance must be set to draw simulated observations")
        dt = _cov_dtype(self.covariance)
        M = self._cov_device().shape[0]
        if simulated_state is None:
            simulated_state = scipy_mv_normal_draw(loc=np.zeros(M), cov=self._cov_device(), ndraws=1).astype(dt)
        simulated_state = np.asarray(simulated_state)
        noise = scipy_mv_normal_draw(loc=np.zeros(self.error_cov.shape[0]), cov=self.error_cov, ndraws=1).astype(dt)
        self.simulated_obs = simulated_state[self.idx] + noise
        # second stage: simple kriging of the simulated observations (:386-387)
        self.simulated_grid = self._simple_apply(self.simulated_obs).astype(dt, copy=False)
        self.epsilon = self.simulated_grid - simulated_state
        # first stage: ordinary kriging of the observations (:390-392)
        if self._batch is not None:
            field = self._field_from_batch()
        else:
            field, _, _ = self._apply_user_weights(self._user_weights, len(self.idx))
        self.gridded_field = field.astype(self._out_dtype(), copy=False)
        return self.gridded_field + self.epsilon

    def _simple_apply(self, values: np.ndarray) -> np.ndarray:
        """W_simple @ values on the device without forming W_simple: C_cross^T (S^-1 values) - one vector solve
        with the existing Cholesky factor and one GEMM row."""
        cov = self._cov_device()
        M = cov.shape[0]
        y = dev.to_device(np.asarray(values, dtype=np.float64), torch.float64)
        b = self._batch
        if self._user_simple_weights is None and b is not None and not getattr(self, "_simple_deleted", False):
            if 0 in b.lu_fallback:
                o = dn.apply_reduce(cov, b.Cx(0), b.T(0), y, ordinary=False)
                return o[dn.OUT_FIELD].cpu().numpy()
            z = dn.padded(b.solve_vectors(0, y[:, None]))            # (n, 2): S^-1 values in column 0
            return dn.gemm(z[:, :1], b.Cx(0), transa=True)[0, :M].cpu().numpy()
        W = np.asarray(self.simple_kriging_weights)
        Tt = dn.padded(np.ascontiguousarray(W.T))
        idx_t = dev.to_device(np.asarray(self.idx), torch.int64)
        o = dn.apply_reduce(cov, dn.padded(cov[idx_t, :M]), Tt, y, ordinary=False)
        return o[dn.OUT_FIELD].cpu().numpy()


def _chol_with_jitter(cov_t: torch.Tensor, n: int) -> torch.Tensor:
    """Lower Cholesky factor (upper triangle zeroed) of the covariance the draws come from.

    A positive definite matrix is factored as it is.  Otherwise the reference's `np.random.multivariate_normal`
    (`stochastic.py:460`) does not fail: it warns and draws from its SVD factor, i.e. from V |Lambda| V^T.  The same
    here: an indefinite matrix (n <= covariance_tools.SPECTRAL_MAX) is replaced by |C| = C sign(C) (device
    eigenvalues + sign-function iteration on DMMA GEMMs) with numpy's warning; what is then still singular gets a
    diagonal lift of 1e-10, 1e-8, 1e-6 x mean(diag) (warning), after which the reference's error is raised."""
    from . import covariance_tools as ct

    src = cov_t[:, :n]
    mean_diag = float(torch.diagonal(src).mean().item())
    last = None
    tried_abs = False
    for jitter in (0.0, 0.0, 1e-10, 1e-8, 1e-6):
        L = dn.padded(src)
        if jitter:
            L[:, :n].diagonal().add_(jitter * mean_diag)
        try:
            dn.factor_inplace(L, n, "cov")
        except np.linalg.LinAlgError as e:
            last = e
            if not tried_abs:
                tried_abs = True
                if 2 <= n <= ct.SPECTRAL_MAX:
                    lam = ct.sym_eigvals(src.contiguous())
                    if float(lam[0].item()) < -1e-12 * float(lam.abs().max().item()):
                        warn("covariance is not symmetric positive-semidefinite.", RuntimeWarning)   # numpy's text
                        src = ct.matrix_abs(src.contiguous(), lam)
            continue
        if jitter:
            warn(f"covariance is not positive definite; drew from cov + {jitter:g} * mean(diag) * I")
            logging.warning("scipy_mv_normal_draw: Cholesky needed a diagonal jitter of %g x mean(diag)", jitter)
        check(lib().gmg_tril_f64(L.data_ptr(), n, L.stride(0), dev.stream_ptr()), "gmg_tril_f64")
        return L
    raise ValueError("Negative eigenvalues are unexpectedly large.") from last


def scipy_mv_normal_draw(
    loc: np.ndarray,
    cov,
    ndraws: int = 1,
    eigen_rtol: float = 1e-6,
    eigen_fudge: float = 1e-8,
    *,
    seed: int | None = None,
) -> np.ndarray:
    """Random multivariate normal draw(s) N(loc, cov) (reference `stochastic.py:395-502`).

    Returns shape (len(loc),) if `ndraws == 1`, else (ndraws, len(loc)).  `cov` may be a numpy array or a
    CUDA tensor.  `seed` keys the device Philox stream; by default it is taken from numpy's global random state,
    so `np.random.seed(s)` makes the result reproducible like it does for the reference.
    `eigen_rtol` / `eigen_fudge` are accepted for signature compatibility (the fallback here is a diagonal
    jitter, see module docstring)."""
    shape = tuple(cov.shape)
    if len(shape) != 2:
        raise ValueError("cov should be 2D.")
    if shape[0] != shape[1]:
        raise ValueError("cov is not a square matrix")
    n = shape[0]
    if seed is None:
        seed = int(np.random.randint(0, 2**31 - 1)) | (int(np.random.randint(0, 2**31 - 1)) << 31)
    cov_t = cov.to(dev.require_cuda(), torch.float64) if isinstance(cov, torch.Tensor) else dev.to_device(
        np.asarray(cov), torch.float64)
    L = _chol_with_jitter(cov_t, n)
    nd = int(ndraws)
    Z = torch.zeros((nd, n + (n & 1)), dtype=torch.float64, device=L.device)
    zflat = torch.empty(nd * n, dtype=torch.float64, device=L.device)
    check(lib().gmg_normal_f64(zflat.data_ptr(), nd * n, seed, 0, dev.stream_ptr()), "gmg_normal_f64")
    Z[:, :n] = zflat.view(nd, n)
    X = dn.gemm(Z[:, :n], L[:, :n], transb=True)       # X = Z L^T  (ndraws x n)
    draw = X[:, :n].cpu().numpy() + np.asarray(loc, dtype=np.float64)
    return draw[0] if ndraws == 1 else draw
