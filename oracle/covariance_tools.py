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
