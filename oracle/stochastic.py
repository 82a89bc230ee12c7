"""Oracle: two-stage stochastic kriging and multivariate-normal draws.

TEST INFRASTRUCTURE ONLY (see `oracle/__init__.py`).  Restatement of
`glomar_gridding/stochastic.py`; each function cites the reference lines it follows.

Draw parity is UNPINNED by construction: the reference draws with
`np.random.multivariate_normal` (legacy global MT19937 RandomState + SVD factor,
stochastic.py:460) and no reference test checks drawn values (test/test_kriging.py:319-372 only
checks shapes and the deterministic ordinary-kriging part).  The oracle therefore exposes the
deterministic sub-results and takes the random state / simulated observations as *inputs*, so the
CUDA path can be compared on injected draws; distributional checks live in the tests.
"""

from __future__ import annotations

import numpy as np

from .kriging import _obs_system, _sqrt_clamped, extended_inverse, subset_error_covariance


def mv_normal_draw(loc, cov, ndraws=1):
    """`scipy_mv_normal_draw` main branch, stochastic.py:452-461: (M,) if ndraws == 1 else (ndraws, M)."""
    cov = np.asarray(cov)
    if cov.ndim != 2:
        raise ValueError("cov should be 2D.")
    if cov.shape[0] != cov.shape[1]:
        raise ValueError("cov is not a square matrix")
    draw = np.random.multivariate_normal(loc, cov, size=ndraws)
    return draw[0] if ndraws == 1 else draw


def stochastic_kriging(cov, idx, obs, error_cov, simulated_state=None, simulated_obs_noise=None,
                       big_temporaries=True):
    """StochasticKriging weights + solve with injectable draws.

    Follows stochastic.py:124-164 (explicit `np.linalg.inv` of C_obs + E), :166-218 (simple weights
    from the inverse, ordinary weights from the bordered inverse), :220-297 (uncertainty and
    constraint mask) and :299-392 (solve: simulated state, simulated obs = state[idx] + noise,
    epsilon = W_sk @ sim_obs - state, result = W_ok @ [obs; 0] + epsilon).
    """
    if error_cov is None:
        raise ValueError("Error Covariance must be provided for StochasticKriging")
    idx, obs, E = subset_error_covariance(idx, obs, error_cov)
    M = cov.shape[0]
    S, Cx = _obs_system(cov, idx, E)
    S_inv = np.linalg.inv(S).astype(cov.dtype)
    W_sk = (S_inv @ Cx).T
    Cx_ext = np.concatenate((Cx, np.ones((1, M), dtype=cov.dtype)), axis=0)
    W_ok = (extended_inverse(S_inv) @ Cx_ext).T
    alpha = W_ok[:, -1]
    if big_temporaries:
        u2 = np.diag(cov - W_ok @ Cx_ext) - alpha
        num = np.diag(cov[:, idx] @ W_sk.T)
    else:
        u2 = np.diag(cov) - np.einsum("mk,km->m", W_ok, Cx_ext) - alpha
        num = np.einsum("mk,mk->m", cov[:, idx], W_sk)
    if simulated_state is None:
        simulated_state = mv_normal_draw(np.zeros(M), cov, 1).astype(cov.dtype)
    if simulated_obs_noise is None:
        simulated_obs_noise = mv_normal_draw(np.zeros(E.shape[0]), E, 1).astype(cov.dtype)
    sim_obs = simulated_state[idx] + simulated_obs_noise
    sim_grid = W_sk @ sim_obs
    eps = sim_grid - simulated_state
    field = W_ok @ np.append(obs, 0).astype(obs.dtype)
    return {"weights": W_ok, "simple_weights": W_sk, "gridded_field": field, "epsilon": eps,
            "simulated_obs": sim_obs, "simulated_grid": sim_grid, "result": field + eps,
            "uncertainty": _sqrt_clamped(u2), "constraint_mask": num / np.diag(cov),
            "idx": idx, "obs": obs, "error_cov": E}
