"""Host-side helpers of the hot path, mirroring the reference's `glomar_gridding/utils.py`.

Only the helpers the ellipse / kriging classes use are provided: they act on
O(N)-sized vectors or are pure bookkeeping; the N x N arithmetic lives in the CUDA library.
"""

from __future__ import annotations

import logging
from itertools import islice
from typing import Iterable
from warnings import warn

import numpy as np


def adjust_small_negative(mat: np.ndarray, atol: float = 1e-8) -> np.ndarray:
    """Set negative values within `atol` of zero to 0, warning as the reference does
    (`utils.py:132-175`: same warning texts, same `print` of the affected values)."""
    small_negative_check = np.logical_and(np.isclose(mat, 0, atol=atol), mat < 0.0)
    ret = mat.copy()
    if small_negative_check.any():
        warn("Small negative vals are detected. Setting to 0.")
        print(mat[small_negative_check])
        ret[small_negative_check] = 0.0
    if (ret < 0).any():
        warn("Negative values are detected")
    return ret.astype(mat.dtype)


def cor_2_cov(cor: np.ndarray, variances: np.ndarray, rounding: int | None = None) -> np.ndarray:
    """Covariance from correlation and variances (reference `utils.py:611-634`)."""
    stdevs = np.sqrt(variances)
    cov = cor * np.outer(stdevs, stdevs)
    cov[cor == 0] = 0
    if rounding is not None:
        cov = np.round(cov, rounding)
    return cov


def cov_2_cor(cov: np.ndarray, rounding: int | None = None) -> np.ndarray:
    """Correlation from covariance, unit diagonal enforced, exact zeros kept (reference `utils.py:637-671`)."""
    stdevs = np.sqrt(np.diag(cov))
    cor = cov / np.outer(stdevs, stdevs)
    if not np.all(np.diag(cor) == 1.0):
        bad_val = np.max(np.abs(np.diag(cor) - 1.0))
        if bad_val > 1e-6:
            raise ValueError(
                "Correlation Diagonal contains values not close to 1. " + f"With difference to 1: {bad_val}"
            )
        print("Numerical error correction applied to correlation matrix diagonal "
              + f"With difference to 1: {bad_val}")
        np.fill_diagonal(cor, 1.0)
    cor[cov == 0] = 0
    if rounding is not None:
        cor = np.round(cor, rounding)
    return cor


def mask_array(arr: np.ndarray) -> np.ma.MaskedArray:
    """Force an array to be a `numpy.ma.MaskedArray` (reference `utils.py:674-694`)."""
    if isinstance(arr, np.ma.MaskedArray):
        return arr
    if isinstance(arr, np.ndarray):
        logging.info("Ad hoc conversion to np.ma.MaskedArray")
        return np.ma.MaskedArray(arr)
    raise TypeError("Input is not a numpy array.")


def uncompress_masked(compressed_array: np.ndarray, mask: np.ndarray, fill_value=0.0, apply_mask: bool = False,
                      dtype=None):
    """Scatter a compressed 1-D array back to the shape of `mask` (reference `utils.py:550-608`)."""
    compressed_array = np.asarray(compressed_array)
    dtype = dtype or compressed_array.dtype
    not_mask = np.logical_not(mask)
    if np.sum(not_mask) != len(compressed_array):
        raise ValueError("Length of compressed_array does not align with mask")
    out = np.empty_like(mask, dtype=dtype)
    np.place(out, not_mask, compressed_array)
    if apply_mask:
        return np.ma.masked_where(mask, out)
    np.place(out, mask, fill_value)
    return out


def batched(iterable: Iterable, n: int, *, strict: bool = False):
    """`itertools.batched` for Python < 3.12 (reference `utils.py:697-712`)."""
    if n < 1:
        raise ValueError("'n' must be >= 1")
    iterator = iter(iterable)
    while batch := tuple(islice(iterator, n)):
        if strict and len(batch) != n:
            raise ValueError("batched(): incomplete batch")
        yield batch


def sizeof_fmt(num: float, suffix: str = "B") -> str:
    """Human-readable byte count."""
    for unit in ("", "Ki", "Mi", "Gi", "Ti", "Pi", "Ei", "Zi"):
        if abs(num) < 1024.0:
            return f"{num:3.1f}{unit}{suffix}"
        num /= 1024.0
    return f"{num:.1f}Yi{suffix}"


def intersect_mtlb(a: np.ndarray, b: np.ndarray):
    """Intersection of two arrays and the index of each common value in `a` and in `b`
    (reference `utils.py:246-279`)."""
    a1, ia = np.unique(a, return_index=True)
    b1, ib = np.unique(b, return_index=True)
    both = np.sort(np.concatenate((a1, b1)))
    common = both[:-1][both[1:] == both[:-1]]
    return common, ia[np.isin(a1, common)], ib[np.isin(b1, common)]


def get_spatial_mean(grid_obs: np.ndarray, covx: np.ndarray) -> float:
    """Best linear unbiased estimate of a constant mean, (1^T C^-1 1)^-1 (1^T C^-1 z)
    (reference `utils.py:766-793`); C^-1 [1, z] by the device Cholesky."""
    from . import _dense as dn

    n = len(grid_obs)
    R = dn.padded(np.column_stack([np.ones(n), np.asarray(grid_obs, dtype=np.float64)]))
    dn.cholesky_solve_inplace(dn.padded(np.asarray(covx)), R, "covx")
    x = R[:, :2].cpu().numpy()
    return float(x[:, 1].sum() / x[:, 0].sum())
