"""Host-side helpers of the hot path, mirroring the reference's `glomar_gridding/utils.py`.

Only the helpers the ellipse / kriging classes use are provided.  The N x N ones (`cov_2_cor`,
`cor_2_cov`) run on the GPU through the C ABI; the rest act on O(N)-sized vectors or are pure bookkeeping.
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


def _as_matrix(a):
    """numpy / torch matrix -> (CUDA tensor float32|float64 with unit column stride, was_tensor, numpy dtype)."""
    import torch

    from . import _device as dev

    if isinstance(a, torch.Tensor):
        dt = a.dtype if a.dtype in (torch.float32, torch.float64) else torch.float64
        t = a.to(dev.require_cuda(), dt)
        return (t if t.stride(1) == 1 else t.contiguous()), True, (np.float32 if dt == torch.float32 else np.float64)
    arr = np.asarray(a)
    npdt = np.float32 if arr.dtype == np.float32 else np.float64
    return dev.to_device(arr, torch.float32 if npdt == np.float32 else torch.float64), False, npdt


def cor_2_cov(cor, variances, rounding: int | None = None, return_device: bool = False):
    """Covariance from correlation and variances (reference `utils.py:611-634`): `cor * outer(sd, sd)` with the
    exact zeros of `cor` kept, evaluated on the GPU (`gmg_cor2cov`) in the matrix' own precision.
    numpy in -> numpy out; CUDA tensor in (or `return_device`) -> CUDA tensor out."""
    import torch

    from . import _device as dev
    from ._lib import check, lib

    C, was_tensor, npdt = _as_matrix(cor)
    n = C.shape[0]
    var = variances if isinstance(variances, torch.Tensor) else torch.from_numpy(np.array(variances, copy=True))
    sd = torch.sqrt(var.to(C.device, C.dtype)).contiguous()                  # O(n) plumbing
    out = torch.empty((n, n), dtype=C.dtype, device=C.device)
    check(lib().gmg_cor2cov(C.data_ptr(), C.stride(0), n, int(npdt == np.float32), sd.data_ptr(), out.data_ptr(), n,
                            dev.stream_ptr()), "gmg_cor2cov")
    if rounding is not None:
        out = torch.round(out, decimals=rounding)
    return out if (was_tensor or return_device) else dev.to_host(out)


def cov_2_cor(cov, rounding: int | None = None, return_device: bool = False):
    """Correlation from covariance, unit diagonal enforced, exact zeros kept (reference `utils.py:637-671`),
    evaluated on the GPU (`gmg_cov2cor`) in the matrix' own precision.
    numpy in -> numpy out; CUDA tensor in (or `return_device`) -> CUDA tensor out."""
    import torch

    from . import _device as dev
    from ._lib import check, lib

    C, was_tensor, npdt = _as_matrix(cov)
    n = C.shape[0]
    sd = torch.empty(n, dtype=C.dtype, device=C.device)
    stat = torch.zeros(2, dtype=torch.int64, device=C.device)
    cor = torch.empty((n, n), dtype=C.dtype, device=C.device)
    check(lib().gmg_cov2cor(C.data_ptr(), C.stride(0), n, int(npdt == np.float32), sd.data_ptr(), cor.data_ptr(), n,
                            stat.data_ptr(), dev.stream_ptr()), "gmg_cov2cor")
    n_bad, dev_bits = (int(x) for x in stat.cpu().numpy())
    if n_bad:
        bad_val = float(np.array([dev_bits], dtype=np.int64).view(np.float64)[0])
        if not bad_val <= 1e-6:
            raise ValueError(
                "Correlation Diagonal contains values not close to 1. " + f"With difference to 1: {bad_val}"
            )
        print("Numerical error correction applied to correlation matrix diagonal "
              + f"With difference to 1: {bad_val}")
        cor.fill_diagonal_(1.0)
    if rounding is not None:
        cor = torch.round(cor, decimals=rounding)
    return cor if (was_tensor or return_device) else dev.to_host(cor)


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
