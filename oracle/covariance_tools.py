"""CPU restatement of the reference's eigenvalue clipping (TEST INFRASTRUCTURE ONLY - see oracle/__init__.py).

Follows `glomar_gridding/covariance_tools.py:633-785` (`explained_variance_clip` + `_eigenvalue_clip`): keep the
leading eigenvalues that explain `target_variance_fraction` of the trace, replace every other eigenvalue by the
mean of the clipped ones (which preserves the trace) and reassemble `V diag(w') V^T`.
Pinned against the unmodified reference through `tests/golden/notebook_2014_03.npz` (the kriged fields of the
clipped covariance) and, when `/root/reference` is present, against the live function.
"""

import numpy as np


def explained_variance_clip(cov: np.ndarray, target_variance_fraction: float = 0.95):
    """Returns (clipped covariance, number of eigenvalues kept, ascending eigenvalues of the input)."""
    w, v = np.linalg.eigh(cov)
    total = w.sum()
    csum = np.cumsum(w[::-1])
    k = int(np.argmax(csum > target_variance_fraction * total)) + 1
    new = w.copy()
    new[: len(w) - k] = (total - w[len(w) - k:].sum()) / (len(w) - k)
    return (v * new) @ v.T, k, w


def laloux_clip(cov: np.ndarray, num_grid_pts: int | None = None, num_time_pts: int = 40) -> np.ndarray:
    """`laloux_clip` (`covariance_tools.py:519-630`): clip the eigenvalues of the correlation matrix that do not
    exceed the Marchenko-Pastur edge (1 + sqrt(q))^2 (`_find_index_aspect_ratio`, `:488-516`) to their mean, then
    scale back with the original variances (`utils.py:611-671`)."""
    num_grid_pts = num_grid_pts or cov.shape[0]
    sd = np.sqrt(np.diag(cov))
    norm = np.outer(sd, sd)
    cor = cov / norm
    np.fill_diagonal(cor, 1.0)
    cor[cov == 0] = 0
    q = num_grid_pts / num_time_pts
    if q < 1.0:
        q = 1.0 / q
    w, v = np.linalg.eigh(cor)
    k = int(np.sum(w > (1.0 + np.sqrt(q)) ** 2.0))
    new = w.copy()
    if k > 0:                                  # k == 0: the reference's `new_eigvals[:0] = ...` changes nothing
        new[: len(w) - k] = (w.sum() - w[len(w) - k:].sum()) / (len(w) - k)
    out = ((v * new) @ v.T) * norm
    out[(v * new) @ v.T == 0] = 0
    return out
