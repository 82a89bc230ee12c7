"""Oracle: simple / ordinary kriging exactly as the reference formulates it (LU on the
Lagrange-extended system, non-textbook 2*alpha variance).

TEST INFRASTRUCTURE ONLY (see `oracle/__init__.py`).  Functional restatement of
`glomar_gridding/kriging.py`; each function cites the reference lines it follows.
"""

from __future__ import annotations

import warnings

import numpy as np


def subset_error_covariance(idx, obs, error_cov):
    """Subset E to the observed cells and drop obs whose E diagonal is NaN or 0.

    Follows `Kriging.subset_error_covariance`, kriging.py:89-121 (same warning text).
    Returns (idx, obs, error_cov); inputs are not mutated.
    """
    idx = np.asarray(idx)
    obs = np.asarray(obs)
    if error_cov is None:
        return idx, obs, None
    E = np.asarray(error_cov)
    if E.shape[0] != len(idx):
        E = E[idx[:, None], idx[None, :]]
    else:
        E = E.copy()
    d = E.diagonal()
    bad = np.logical_or(np.isnan(d), d == 0)
    if bad.any():
        keep = np.where(~bad)[0]
        msg = ("Have nans or zeros on the error covariance diagonal. At positions "
               + " ,".join(map(str, idx[bad].tolist())) + ". Filtering input accordingly")
        warnings.warn(msg)
        idx, obs, E = idx[keep], obs[keep], E[keep[:, None], keep[None, :]]
    E[np.isnan(E)] = 0.0
    return idx, obs, E


def adjust_small_negative(vec: np.ndarray, atol: float = 1e-8) -> np.ndarray:
    """(-atol, 0) -> 0; `utils.adjust_small_negative`, utils.py:132-175 (warnings omitted)."""
    out = vec.copy()
    out[np.logical_and(np.isclose(vec, 0, atol=atol), vec < 0.0)] = 0.0
    return out.astype(vec.dtype)


def _obs_system(cov, idx, error_cov):
    S = cov[idx[:, None], idx[None, :]]  # kriging.py:337, :571 (fancy-index copy)
    Cx = cov[idx, :]  # :338, :572
    if error_cov is not None:
        S = S + error_cov  # :342, :576
    return S, Cx


def _sqrt_clamped(u2):
    u = np.sqrt(adjust_small_negative(u2))  # :436-438, :702-704
    u[np.isnan(u)] = 0.0
    return u


def simple_kriging(cov, idx, obs, error_cov=None, mean=0.0, big_temporaries=True):
    """SimpleKriging: weights, field, uncertainty, constraint mask.

    Follows kriging.py:318-345 (weights by `np.linalg.solve`), :414-417 (solve), :419-439
    (uncertainty) and :441-491 (constraint mask).  With `big_temporaries` the M x M products the
    reference forms (:434, :487) are formed too (that is what the CPU baseline times); otherwise
    only their diagonals are evaluated (needed at 1 degree where M x M does not fit).
    """
    idx, obs, E = subset_error_covariance(idx, obs, error_cov)
    S, Cx = _obs_system(cov, idx, E)
    W = np.linalg.solve(S, Cx).T  # (M, n), :343
    z = W @ obs + mean
    if big_temporaries:
        u2 = np.diag(cov - W @ Cx)
        num = np.diag(cov[:, idx] @ W.T)
    else:
        u2 = np.diag(cov) - np.einsum("mk,km->m", W, Cx)
        num = np.einsum("mk,mk->m", cov[:, idx], W)
    return {"weights": W, "field": z, "uncertainty": _sqrt_clamped(u2),
            "constraint_mask": num / np.diag(cov), "idx": idx, "obs": obs, "error_cov": E}


def ordinary_kriging(cov, idx, obs, error_cov=None, big_temporaries=True):
    """OrdinaryKriging: extended weights (M, n+1), field, uncertainty, constraint mask.

    Follows kriging.py:540-588 (Lagrange-extended LU solve, weights cast to cov.dtype), :674-679
    (solve appends a 0 to obs), :681-706 (uncertainty: diag(C - W_ext @ Cx_ext) - alpha, i.e. the
    Lagrange column is subtracted twice - H4 in SURVEY.md) and :708-776 (constraint mask from a
    second, simple-kriging solve).
    """
    idx, obs, E = subset_error_covariance(idx, obs, error_cov)
    n, M = len(idx), cov.shape[0]
    S, Cx = _obs_system(cov, idx, E)
    one_n = np.ones((1, n), dtype=cov.dtype)
    S_ext = np.block([[S, one_n.T], [one_n, np.zeros((1, 1), dtype=cov.dtype)]])
    Cx_ext = np.concatenate((Cx, np.ones((1, M), dtype=cov.dtype)), axis=0)
    W = np.linalg.solve(S_ext, Cx_ext).T.astype(cov.dtype)  # (M, n+1)
    z = W @ np.append(obs, 0).astype(obs.dtype)
    alpha = W[:, -1]
    W_sk = np.linalg.solve(S, Cx).T
    if big_temporaries:
        u2 = np.diag(cov - W @ Cx_ext) - alpha
        num = np.diag(cov[:, idx] @ W_sk.T)
    else:
        u2 = np.diag(cov) - np.einsum("mk,km->m", W, Cx_ext) - alpha
        num = np.einsum("mk,mk->m", cov[:, idx], W_sk)
    return {"weights": W, "simple_weights": W_sk, "field": z, "uncertainty": _sqrt_clamped(u2),
            "constraint_mask": num / np.diag(cov), "idx": idx, "obs": obs, "error_cov": E}


def extended_inverse(simple_inv: np.ndarray) -> np.ndarray:
    """Inverse of [[S, 1], [1^T, 0]] from S^-1 by the bordered-matrix (Guttman) identity.

    Follows `_extended_inverse`, kriging.py:815-829.
    """
    if simple_inv.ndim != 2:
        raise ValueError("S must be a matrix")
    ones = np.ones((simple_inv.shape[0], 1), dtype=simple_inv.dtype)
    e = simple_inv @ ones
    finv = 1 / (0 - ones.T @ e)
    g = finv * e.T
    return np.block([[simple_inv + e @ g, -g.T], [-g, finv]])
