"""Deterministic synthetic inputs for the benchmark and parity configurations.

Pure numpy input generation - no covariance or kriging arithmetic happens here.  Shared by
`bench.py`, the tests and `tests/golden/make_golden.py` so that every arm sees identical inputs.
"""

from __future__ import annotations

import numpy as np


def global_grid(res_deg: float) -> tuple[np.ndarray, np.ndarray]:
    """Cell-centre latitudes / longitudes of a regular global grid (5 deg -> 36 x 72, 1 deg -> 180 x 360)."""
    lats = np.arange(-90 + res_deg / 2, 90, res_deg, dtype=np.float64)
    lons = np.arange(-180 + res_deg / 2, 180, res_deg, dtype=np.float64)
    return lats, lons


def ellipse_fields(lats: np.ndarray, lons: np.ndarray, land_mask: bool = False):
    """Smooth, longitude-periodic ellipse parameter fields (Lx > Ly everywhere), each (nlat, nlon).

    Lx in 900..2700 km, Ly in 600..1200 km, theta in +-0.6 rad, stdev in 0.8..1.2.  With
    `land_mask` about 30 % of the cells are masked (returned as numpy MaskedArrays).
    """
    lon_g, lat_g = np.meshgrid(np.deg2rad(lons), np.deg2rad(lats))
    Lx = 1800.0 + 900.0 * np.cos(lat_g) * np.cos(2 * lon_g)
    Ly = 900.0 + 300.0 * np.sin(2 * lat_g) * np.sin(lon_g)
    theta = 0.6 * np.sin(lon_g) * np.cos(lat_g)
    stdev = 0.8 + 0.4 * np.cos(lat_g) ** 2
    if land_mask:
        land = (np.sin(3 * lon_g + 0.7) * np.cos(2 * lat_g - 0.3) + 0.35 * np.sin(lat_g)) > 0.42
        Lx, Ly, theta, stdev = (np.ma.masked_where(land, a) for a in (Lx, Ly, theta, stdev))
    return Lx, Ly, theta, stdev


def month_observations(month: int, n_cells: int, n_min: int = 600, n_max: int = 1600):
    """Observation cells and values for synthetic month `month` (config 2): sorted unique idx, N(0,1) obs."""
    rng = np.random.default_rng(month)
    n = int(rng.integers(n_min, n_max + 1))
    idx = np.sort(rng.choice(n_cells, size=n, replace=False)).astype(np.int64)
    obs = rng.standard_normal(n)
    return idx, obs


def fixed_observations(seed: int, n_cells: int, n_obs: int):
    """`n_obs` distinct sorted cells and N(0,1) values (config 4 uses seed 4, n_obs 20 000)."""
    rng = np.random.default_rng(seed)
    idx = np.sort(rng.choice(n_cells, size=n_obs, replace=False)).astype(np.int64)
    return idx, rng.standard_normal(n_obs)


def hadcrut_style_error_cov(idx: np.ndarray, nlon: int, seed: int = 7, n_groups: int = 12):
    """Obs-space (n x n) "HadCRUT-style" error covariance: sum_g s_g^2 1_g 1_g^T + diag(u^2).

    Groups are longitude bands ("platform groups"); s_g ~ U(0.05, 0.2), u ~ U(0.2, 0.4).  Mirrors
    the block + diagonal structure produced by reference `error_covariance.py:99-186`.
    """
    rng = np.random.default_rng(seed)
    s = rng.uniform(0.05, 0.2, size=n_groups)
    u = rng.uniform(0.2, 0.4, size=len(idx))
    g = ((idx % nlon) * n_groups) // nlon
    n = len(idx)
    s2 = s[g] ** 2
    E = np.empty((n, n), dtype=np.float64)
    for r0 in range(0, n, 2048):                  # row slabs: no n x n temporaries (3.2 GB each at n = 20 000)
        r1 = min(n, r0 + 2048)
        np.multiply(g[r0:r1, None] == g[None, :], s2[r0:r1, None], out=E[r0:r1])
    E[np.arange(n), np.arange(n)] += u ** 2
    return E


def full_grid_error_cov(idx: np.ndarray, n_cells: int, nlon: int, seed: int = 7):
    """Full-grid (M x M) variant: NaN off the observed cells' block and a zero diagonal at
    unobserved cells, exercising the subset + zero/NaN-diagonal filter (reference kriging.py:94-120)."""
    E = np.full((n_cells, n_cells), np.nan)
    E[np.arange(n_cells), np.arange(n_cells)] = 0.0
    E[np.ix_(idx, idx)] = hadcrut_style_error_cov(idx, nlon, seed)
    return E


def training_cube(res_deg: float = 10.0, nt: int = 48, seed: int = 11):
    """Synthetic training data for the ellipse fit (`EllipseBuilder`): `nt` time steps on a regular global grid drawn
    from a smooth non-stationary exponential covariance (length scale 1 500 km at the poles, 2 400 km at the equator),
    with the land mask of `ellipse_fields` (all times) and three cells masked at a single time each (the time-varying
    mask the reference discards, estimate.py:126-132).  Returns (numpy MaskedArray (nt, nlat, nlon), lats, lons)."""
    lats = np.arange(-90.0 + res_deg / 2, 90.0, res_deg)
    lons = np.arange(-180.0 + res_deg / 2, 180.0, res_deg)
    lon_g, lat_g = np.meshgrid(np.deg2rad(lons), np.deg2rad(lats))
    xyz = np.stack([np.cos(lat_g) * np.cos(lon_g), np.cos(lat_g) * np.sin(lon_g), np.sin(lat_g)], -1).reshape(-1, 3)
    chord = np.linalg.norm(xyz[:, None, :] - xyz[None, :, :], axis=-1) * 6371.0
    scale = (1500.0 + 900.0 * np.cos(lat_g).ravel())
    L = np.sqrt(scale[:, None] * scale[None, :])
    C = np.exp(-chord / L) + 1e-6 * np.eye(chord.shape[0])
    w, V = np.linalg.eigh(C)
    A = V * np.sqrt(np.clip(w, 1e-9, None))
    rng = np.random.default_rng(seed)
    data = (rng.standard_normal((nt, A.shape[0])) @ A.T).reshape(nt, lats.size, lons.size)
    land = (np.sin(3 * lon_g + 0.7) * np.cos(2 * lat_g - 0.3) + 0.35 * np.sin(lat_g)) > 0.42
    mask = np.broadcast_to(land, data.shape).copy()
    for t, i, j in ((5, 3, 4), (17, 9, 20), (30, 12, 1)):
        mask[t % nt, i % lats.size, j % lons.size] = True
    return np.ma.masked_array(data, mask), lats, lons
