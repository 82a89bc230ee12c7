"""`eigenvalue_clip` on the GPU: drop-in for `glomar_gridding.covariance_tools.eigenvalue_clip` /
`explained_variance_clip` - the step that sits between the covariance build and kriging
in the reference's workflow (notebook cell 20) and makes an indefinite ellipse covariance positive definite.

Reference (`covariance_tools.py:633-785`): full `numpy.linalg.eigh`, keep the k largest eigenvalues whose sum
first exceeds `target_variance_fraction * trace`, replace ALL others by their mean (so the trace is preserved),
rebuild `V diag(lambda') V^T`.  Since the replaced eigenvalues all become one constant c,

    out = c I + V_k (Lambda_k - c I) V_k^T ,      c = (trace - sum(Lambda_k)) / (n - k)

needs only the k kept eigenpairs - or, equivalently, the spectral PROJECTOR P onto their span:

    out = c I + P (C - c I) ,      P = (I + sign(C - tau I)) / 2 ,   lambda_{k+1} < tau < lambda_k .

Default route (n <= 8192): all eigenVALUES from a Householder tridiagonalisation + Sturm bisection on the device
(`gmg_sym_eigvals_f64`, csrc/spectrum.cu) give k, c and tau; sign(C - tau I) is a Newton-Schulz iteration
X <- (mu/2) X (3 I - mu^2 X^2) - two DMMA GEMMs per step - whose scaling mu is chosen each step from the KNOWN
spectrum (the scalar recurrence is run on the eigenvalues alongside), which cuts the ~25 unscaled steps to ~12.
Larger matrices: orthogonal (subspace) iteration with Rayleigh-Ritz for the kept eigenpairs - every step a DMMA GEMM
against C, Cholesky-QR (Gram matrix, `gmg_potrf`, forward substitution) keeps the block orthonormal, and only the
small m x m projected matrix is diagonalised (`gmg_jacobi_eigh_f64`).  No CPU fallback.

Differences from the reference: `method="Laloux_2000"` is limited to n <= 8192; the progress `print`s are dropped; the
final "is it positive definite" check (`:779-784`, a second full eigendecomposition) is replaced by the sign of c
and of the kept eigenvalues, which decide it.
"""

from __future__ import annotations

import ctypes
from typing import Literal
from warnings import warn

import numpy as np
import torch

from . import _dense as dn
from . import _device as dev
from ._lib import check, lib
from .utils import cor_2_cov, cov_2_cor


FULL_JACOBI_MAX = 4096      # up to this order C can be diagonalised directly (Jacobi); beyond it, subspace iteration
SPECTRAL_MAX = 8192         # up to this order: tridiagonal eigenvalues + sign-function projector (the default route)
MIN_REL_GAP = 1e-9          # below this relative gap at the cut the projector is ill-conditioned: use eigenvectors


def sym_eigvals(C: torch.Tensor) -> torch.Tensor:
    """All eigenvalues (ascending) of a symmetric device matrix: `gmg_sym_eigvals_f64` on a padded copy."""
    n = C.shape[0]
    A = dn.padded(C)                                          # the kernel destroys its input
    w = torch.empty(n, dtype=torch.float64, device=C.device)
    nbytes = ctypes.c_size_t()
    check(lib().gmg_sym_eigvals_ws_bytes(n, ctypes.byref(nbytes)), "gmg_sym_eigvals_ws_bytes")
    ws = torch.empty(nbytes.value, dtype=torch.uint8, device=C.device)
    check(lib().gmg_sym_eigvals_f64(A.data_ptr(), A.stride(0), n, w.data_ptr(), ws.data_ptr(), nbytes.value,
                                    dev.stream_ptr()), "gmg_sym_eigvals_f64")
    return w


def matrix_sign(C: torch.Tensor, lam: torch.Tensor, tau: float = 0.0, max_steps: int = 80, zero_tol: float = 1e-13):
    """sign(C - tau I) of a symmetric device matrix by the scaled Newton-Schulz iteration X <- (mu/2) X (3 I - mu^2 X^2)
    (two DMMA GEMMs per step).  `lam` = all eigenvalues of C (`sym_eigvals`): the scalar recurrence is run on them
    alongside the matrix one, which gives the optimal scaling mu = sqrt(3 / (a^2 + a b + b^2)) of every step from the
    current extreme moduli a, b, and the step count.  Eigenvalues within `zero_tol` (relative) of tau are treated as
    zeros of the sign function.  Returns (X as an (n, even ld) buffer, steps)."""
    n = C.shape[0]
    X = dn.padded(C)
    X[:, :n].diagonal().sub_(tau)
    x = lam - tau                                             # the spectrum of X, iterated alongside
    s = float(x.abs().max().item()) * (1.0 + 1e-12)
    if s == 0.0:
        return torch.zeros_like(X), 0
    X.div_(s)
    x = x / s
    x = x[x.abs() > zero_tol]
    steps = 0
    for steps in range(max_steps + 1):
        if x.numel() == 0:
            break
        ax = x.abs()
        a, b = float(ax.min().item()), float(ax.max().item())
        if 1.0 - a <= 4e-16 and b - 1.0 <= 4e-16:
            break
        if steps == max_steps:
            raise np.linalg.LinAlgError("sign-function iteration did not converge")
        mu = (3.0 / (a * a + a * b + b * b)) ** 0.5
        Y = torch.zeros_like(X)
        dn.gemm(X[:, :n], X[:, :n], out=Y[:, :n])             # X^2
        Xn = X.clone()
        dn.gemm(X[:, :n], Y[:, :n], out=Xn[:, :n], alpha=-0.5 * mu ** 3, beta=1.5 * mu)
        x = 0.5 * mu * x * (3.0 - (mu * x) ** 2)
        X = Xn
    return X, steps


def matrix_abs(C: torch.Tensor, lam: torch.Tensor | None = None) -> torch.Tensor:
    """|C| = V |Lambda| V^T = C sign(C) of a symmetric device matrix - the covariance `numpy.random.multivariate_normal`
    actually draws from when it is handed an indefinite matrix (its SVD factor uses the singular values)."""
    n = C.shape[0]
    lam = sym_eigvals(C) if lam is None else lam
    X, _ = matrix_sign(C, lam)
    Cp = dn.padded(C)
    out = torch.zeros_like(Cp)
    dn.gemm(Cp[:, :n], X[:, :n], out=out[:, :n])
    res = out[:, :n]
    return (0.5 * (res + res.t())).contiguous()


def _projector_clip(C: torch.Tensor, lam: torch.Tensor, tau: float, c: float, max_steps: int = 80):
    """c I + P (C - c I) with P the spectral projector of C for the eigenvalues above tau; `lam` = all eigenvalues.
    Returns (matrix (n, n) view of an even-ld buffer, Newton-Schulz steps taken)."""
    n = C.shape[0]
    X, steps = matrix_sign(C, lam, tau, max_steps, zero_tol=0.0)
    Cc = dn.padded(C)
    Cc[:, :n].diagonal().sub_(c)
    out = Cc.clone()
    dn.gemm(X[:, :n], Cc[:, :n], out=out[:, :n], alpha=0.5, beta=0.5)   # (C - cI)/2 + sign (C - cI)/2 = P (C - cI)
    res = out[:, :n]
    res.diagonal().add_(c)
    return res, steps


def _jacobi_eigh(T: torch.Tensor, max_sweeps: int = 30, tol: float = 1e-15):
    """Eigen-decomposition of a small symmetric device matrix: (eigenvalues descending, eigenvectors as columns)."""
    m0 = T.shape[0]
    m = m0 + (m0 & 1)
    A = torch.zeros((m, m), dtype=torch.float64, device=T.device)
    A[:m0, :m0] = T
    U = torch.empty((m, m), dtype=torch.float64, device=T.device)
    nbytes = ctypes.c_size_t()
    check(lib().gmg_jacobi_eigh_ws_bytes(m, ctypes.byref(nbytes)), "gmg_jacobi_eigh_ws_bytes")
    ws = torch.zeros(nbytes.value, dtype=torch.uint8, device=T.device)
    done = torch.zeros(1, dtype=torch.int32, device=T.device)
    check(lib().gmg_jacobi_eigh_f64(A.data_ptr(), m, U.data_ptr(), m, m, max_sweeps, tol, ws.data_ptr(), nbytes.value,
                                    done.data_ptr(), dev.stream_ptr()), "gmg_jacobi_eigh_f64")
    w = torch.diagonal(A)[:m0].clone()
    if m != m0:                                   # drop the padding eigenpair (exactly decoupled: zero row / column)
        keep = torch.arange(m, device=T.device) != int(torch.argmax(U[m0].abs()))
        w, U = torch.diagonal(A)[keep].clone(), U[:m0][:, keep]
    order = torch.argsort(w, descending=True)     # plumbing: sort of m numbers
    return w[order], U[:, order].contiguous()


def top_eigenpairs(C: torch.Tensor, m: int, iters: int = 60, tol: float = 1e-13, seed: int = 0):
    """The m algebraically largest eigenpairs of the symmetric device matrix C by orthogonal iteration.
    Returns (theta (m,) descending, Vt (m, n) with the eigenvectors as ROWS, residual estimate per pair)."""
    n = C.shape[0]
    m = min(m, n)
    ldn = n + (n & 1)
    Cp = C if C.stride(0) % 2 == 0 else dn.padded(C)
    g = torch.Generator(device=C.device)
    g.manual_seed(seed)
    Qt = torch.zeros((m, ldn), dtype=torch.float64, device=C.device)
    Qt[:, :n] = torch.randn((m, n), dtype=torch.float64, device=C.device, generator=g)

    def orthonormalise(Yt):
        # Cholesky-QR twice: G = Y Y^T = L L^T, Q = L^-1 Y (forward substitution on DMMA tiles)
        for _ in range(2):
            G = dn.padded(dn.gemm(Yt[:, :n], Yt[:, :n], transb=True))
            ws = dn.factor_inplace(G, m, "Gram matrix of the subspace iteration")
            d = np.zeros(1, dtype=dn.SOLVE_DESC)
            d["S"], d["lds"], d["Linv"], d["P"], d["ldp"], d["info"] = (
                G.data_ptr(), G.stride(0), ws["Linv"].data_ptr(), ws["P"].data_ptr(), ws["ldp"], ws["info"].data_ptr())
            d["R"], d["ldr"], d["n"], d["nrhs"] = Yt.data_ptr(), Yt.stride(0), m, n
            dt = dn._desc_tensor(d, C.device)
            check(lib().gmg_potrs_batched_f64(dt.data_ptr(), 1, m, n, 1, dev.stream_ptr()), "gmg_potrs_batched_f64")
            torch.cuda.current_stream().synchronize()
        return Yt

    Qt = orthonormalise(Qt)
    theta_prev = None
    for it in range(iters):
        Yt = torch.zeros_like(Qt)
        dn.gemm(Qt[:, :n], Cp[:, :n], out=Yt[:, :n])              # Y^T = Q^T C   (C symmetric)
        if it % 8 == 7 or it == iters - 1:
            T = dn.gemm(Yt[:, :n], Qt[:, :n], transb=True)        # Rayleigh-Ritz matrix Q^T C Q
            T = 0.5 * (T + T.t())
            theta, U = _jacobi_eigh(T)
            Vt = torch.zeros_like(Qt)
            dn.gemm(dn.padded(U), Qt[:, :n], transa=True, out=Vt[:, :n])   # Ritz vectors (rows) = U^T Q^T
            if theta_prev is not None:
                half = max(1, m // 2)
                change = float(((theta[:half] - theta_prev[:half]).abs() / theta[:half].abs().clamp_min(1e-300)).max())
                if change <= tol:
                    Qt = Vt
                    break
            theta_prev = theta
            Qt = Vt
            Yt = torch.zeros_like(Qt)
            dn.gemm(Qt[:, :n], Cp[:, :n], out=Yt[:, :n])
        Qt = orthonormalise(Yt)
    # final Rayleigh-Ritz on the converged basis + residuals ||C v - theta v||
    Yt = torch.zeros_like(Qt)
    dn.gemm(Qt[:, :n], Cp[:, :n], out=Yt[:, :n])
    T = dn.gemm(Yt[:, :n], Qt[:, :n], transb=True)
    theta, U = _jacobi_eigh(0.5 * (T + T.t()))
    Vt = torch.zeros_like(Qt)
    dn.gemm(dn.padded(U), Qt[:, :n], transa=True, out=Vt[:, :n])
    CV = torch.zeros_like(Qt)
    dn.gemm(Vt[:, :n], Cp[:, :n], out=CV[:, :n])
    resid = (CV[:, :n] - theta[:, None] * Vt[:, :n]).norm(dim=1)
    return theta, Vt, resid


def _clip_keep_top(C: torch.Tensor, lam: torch.Tensor, k: int, total: float):
    """`_eigenvalue_clip` (reference `covariance_tools.py:731-785`) from the spectrum alone: keep the k largest
    eigenvalues of C, replace the others by their mean.  `lam` = all eigenvalues ascending, `total` = their sum.
    Returns the clipped matrix, or None when the eigenvalues at the cut are too close for the projector route."""
    n = C.shape[0]
    theta = lam.flip(0)
    kept = float(theta[:k].sum().item())
    if total < kept:
        raise ValueError("Variance explained by retained eigenvalues exceeds total variance. Resulting matrix will "
                         "have negative eigenvalues. Try using a lower threshold.")
    if k <= 0 or k >= n:                                      # nothing is clipped (reference: `new_eigvals[:0] = ...`)
        res, pd = C.clone(), float(lam[0].item()) > 0
    else:
        scale = float(theta.abs().max().item())
        if float((theta[k - 1] - theta[k]).item()) <= MIN_REL_GAP * scale:
            return None
        c = (total - kept) / (n - k)
        res, _ = _projector_clip(C, lam, float((0.5 * (theta[k - 1] + theta[k])).item()), c)
        pd = c > 0 and float(theta[k - 1].item()) > 0
    if not pd:
        warn("Resulting matrix is not positive-definite, and may not be a valid covariance matrix.")
    return 0.5 * (res + res.t())


def laloux_clip(cov, num_grid_pts: int | None = None, num_time_pts: int = 40, return_device: bool = False):
    """Clip the eigenvalues of the CORRELATION matrix below the Marchenko-Pastur noise level (1 + sqrt(q))^2,
    q = num_grid_pts / num_time_pts (inverted if < 1), to their mean and scale back to a covariance
    (reference `covariance_tools.py:488-630`; `cov_2_cor` / `cor_2_cov` `utils.py:611-671`)."""
    on_device = isinstance(cov, torch.Tensor)
    C = cov.to(dev.require_cuda(), torch.float64) if on_device else dev.to_device(np.asarray(cov), torch.float64)
    n = C.shape[0]
    if n > SPECTRAL_MAX:
        raise NotImplementedError(f"method='Laloux_2000' is limited to matrices of order <= {SPECTRAL_MAX} on the device")
    num_grid_pts = num_grid_pts or n
    variances = torch.diagonal(C).clone()
    cor = cov_2_cor(C)                                        # gmg_cov2cor; raises like the reference on a bad diagonal
    q = num_grid_pts / num_time_pts
    if q < 1.0:
        q = 1.0 / q
    lam = sym_eigvals(cor)
    k = int((lam > (1.0 + np.sqrt(q)) ** 2.0).sum().item())
    res = _clip_keep_top(cor, lam, k, float(lam.sum().item()))
    if res is None:
        if n > FULL_JACOBI_MAX:
            raise np.linalg.LinAlgError("degenerate eigenvalues at the Laloux cut")
        theta, U = _jacobi_eigh(cor)
        c = float((theta[k:].sum() / (n - k)).item())
        new = theta.clone()
        new[k:] = c
        res = dn.gemm(dn.padded(U * new[None, :]), dn.padded(U), transb=True)[:, :n]
        res = 0.5 * (res + res.t())
    out = cor_2_cov(res.contiguous(), variances)              # gmg_cor2cov
    return out if (on_device or return_device) else dev.to_host(out)


def explained_variance_clip(cov, target_variance_fraction: float = 0.95, return_device: bool = False):
    """Clip all eigenvalues beyond `target_variance_fraction` of explained variance to their mean
    (reference `covariance_tools.py:633-728`, `_eigenvalue_clip` `:731-785`)."""
    if not 0.0 < target_variance_fraction <= 1.0:
        raise ValueError("'target_variance_fraction' must be (0, 1.0]")
    on_device = isinstance(cov, torch.Tensor)
    C = cov.to(dev.require_cuda(), torch.float64) if on_device else dev.to_device(np.asarray(cov), torch.float64)
    n = C.shape[0]
    total = float(torch.diagonal(C).sum().item())            # trace = sum of all eigenvalues
    target = target_variance_fraction * total
    if n <= SPECTRAL_MAX and n >= 2:
        lam = sym_eigvals(C)                                  # ascending
        csum = torch.cumsum(lam.flip(0), 0)
        over = torch.nonzero(csum > target)
        if not over.numel():
            raise ValueError("Target Explained Variance not exceeded")
        res = _clip_keep_top(C, lam, int(over[0].item()) + 1, total)
        if res is not None:
            return res.contiguous() if (on_device or return_device) else dev.to_host(res.contiguous())
        # (nearly) degenerate eigenvalues at the cut: fall through to the eigenvector routes below
    if n <= FULL_JACOBI_MAX:
        # small enough for a full Jacobi diagonalisation of C itself (one cooperative kernel): cheaper than growing
        # the subspace when many eigenvalues have to be kept (the synthetic 5-degree fields keep 780 of 2592)
        theta, U = _jacobi_eigh(C)
        Vt = dn.padded(U.t().contiguous())
        resid = torch.zeros_like(theta)
        csum = torch.cumsum(theta, 0)
        over = torch.nonzero(csum > target)
        if not over.numel():
            raise ValueError("Target Explained Variance not exceeded")
        k = int(over[0].item()) + 1
    else:
        m = min(n, 512)
        while True:
            theta, Vt, resid = top_eigenpairs(C, m)
            csum = torch.cumsum(theta, 0)
            # only the better-converged leading part of the block is trusted for the cut (the rest is oversampling)
            trusted = m if m == n else max(1, (2 * m) // 3)
            over = torch.nonzero(csum[:trusted] > target)
            if over.numel():
                k = int(over[0].item()) + 1
                break
            if m == n:
                raise ValueError("Target Explained Variance not exceeded")
            m = min(n, 2 * m)
    kept = float(csum[k - 1].item())
    if total < kept:
        raise ValueError("Variance explained by retained eigenvalues exceeds total variance. Resulting matrix will "
                         "have negative eigenvalues. Try using a lower threshold.")
    c = (total - kept) / (n - k) if n > k else 0.0
    scale = float(theta[0].abs().item())
    if float(resid[:k].max().item()) > 1e-9 * scale:
        warn(f"eigenvalue_clip: kept eigenpairs converged to {float(resid[:k].max().item()) / scale:.1e} (relative "
             "residual) only")
    ld = Vt.stride(0)
    W = torch.zeros((k, ld), dtype=torch.float64, device=C.device)
    W[:, :n] = Vt[:k, :n] * (theta[:k, None] - c)             # (Lambda_k - c) V_k^T
    out = torch.zeros((n, ld), dtype=torch.float64, device=C.device)
    dn.gemm(Vt[:k, :n], W[:, :n], transa=True, out=out[:, :n])   # V_k (Lambda_k - c) V_k^T
    res = out[:, :n]
    res.diagonal().add_(c)
    if not (c > 0 and float(theta[k - 1].item()) > 0):
        warn("Resulting matrix is not positive-definite, and may not be a valid covariance matrix.")
    res = 0.5 * (res + res.t())
    return res.contiguous() if (on_device or return_device) else dev.to_host(res.contiguous())


def eigenvalue_clip(cov, method: Literal["explained_variance", "Laloux_2000"] = "explained_variance", **kwargs):
    """Denoise / repair a covariance matrix by eigenvalue clipping (reference `covariance_tools.py:788-847`)."""
    match method:
        case "explained_variance":
            return explained_variance_clip(cov, **kwargs)
        case "Laloux_2000":
            return laloux_clip(cov, **kwargs)
        case _:
            raise ValueError(f"Unknown clipping method: {method}")


# -- scalar helpers of the clipping path (reference `covariance_tools.py:393-517`): O(n) logic on a spectrum, host side --
def csum_up_to_val(vals: np.ndarray, target: float, reverse: bool = True, niter: int = 0, csum: float = 0.0):
    """Cumulative sum and index at which `target` is surpassed (reference `:393-460`; its own checks:
    `test/test_covariance_tools.py:140-170`).  With `reverse` the vector is summed from its end and the returned index
    is negative, i.e. valid for the un-reversed vector.  Warns when `vals` is empty or the target is never exceeded."""
    vals = np.asarray(vals)
    if vals.size == 0:
        warn("`vals` is empty")
        return csum, niter
    if len(vals) != vals.size:
        raise ValueError("`vals` must be a vector")
    vals = vals[::-1] if reverse else vals
    i = 0
    for v in vals:                                     # accumulate(vals, initial=csum): the initial value is entry 0
        if csum > target:
            return csum, niter + (-i if reverse else i)
        csum = csum + v
        i += 1
    if csum > target:
        return csum, niter + (-i if reverse else i)
    warn("Out of `vals`, target not exceeded.")
    return csum, niter + (-i if reverse else i)


def clean_small(matrix: np.ndarray, atol: float = 1e-5) -> np.ndarray:
    """Set entries with abs(x) < atol to 0 (reference `:463-470`)."""
    cleaned = np.array(matrix, copy=True)
    cleaned[np.abs(cleaned) < atol] = 0.0
    return cleaned


def check_symmetric(a: np.ndarray, rtol: float = 1e-5, atol: float = 1e-8) -> bool:
    """`np.allclose(a, a.T)` (reference `:108-114`)."""
    a = np.asarray(a)
    return bool(np.allclose(a, a.T, rtol=rtol, atol=atol))


def _find_index_explained_variance(eigvals, target: float = 0.95) -> int:
    """(Negative) index of the eigenvalue - ascending spectrum - at which the cumulative sum from the top exceeds
    `target` of the total variance (reference `:473-485`)."""
    eigvals = np.asarray(eigvals)
    total_variance = np.sum(eigvals)
    target_explained_variance = target * total_variance
    csum, i2goal = csum_up_to_val(eigvals, target_explained_variance)
    if csum <= target_explained_variance:
        raise ValueError("Target Explained Variance not exceeded")
    return i2goal


def _find_index_aspect_ratio(eigvals: np.ndarray, num_grid_pts: int = 180 * 360, num_times: int = 41 * 6) -> int:
    """(Negative) number of eigenvalues above the Marchenko-Pastur edge (1 + sqrt(q))^2, q = grid points / times
    (reference `:488-516`, 7.2.2 in Bun et al.)."""
    q = num_grid_pts / num_times
    if q < 1.0:
        q = 1.0 / q
    threshold = (1.0 + np.sqrt(q)) ** 2.0
    return -int(np.sum(np.asarray(eigvals) > threshold))
