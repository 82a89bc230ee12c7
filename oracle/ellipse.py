"""Oracle: Paciorek-Schervish non-stationary Matern covariance from per-cell ellipse parameters.

TEST INFRASTRUCTURE ONLY (see `oracle/__init__.py`).  numpy/scipy restatement of
`glomar_gridding/ellipse/covariance.py` in the reference; every function cites the lines it
follows.  The arithmetic order of the pair formula is kept identical to the reference so that the
restatement is bit-for-bit equal to it in float64 (checked in tests/test_oracle_vs_golden.py against fixtures made by the reference, and there against the live reference when /root/reference is present).

The structure is NOT the reference's: instead of materialising all `itertools.combinations`
index pairs (reference `:255`, 34 GB of indices at 1 degree) the matrix is evaluated by row strips
of the upper triangle, and arbitrary (i, j) samples or row blocks can be evaluated without
building the whole matrix - which is what the 1-degree parity tests and the CPU baseline need.
"""

from __future__ import annotations

import numpy as np
from scipy.special import gamma as _gamma
from scipy.special import kv as _kv

RADIUS_OF_EARTH_KM = 6371.0  # reference constants.py:18
TWO_PI = 2 * np.pi  # reference ellipse/covariance.py:37


def sigma_rot(Lx: np.ndarray, Ly: np.ndarray, theta: np.ndarray) -> np.ndarray:
    """Flattened Sigma_i = R(theta) diag(Lx^2, Ly^2) R(theta)^T, (N, 4).

    Follows `_sigma_rot_func_multi`, ellipse/covariance.py:512-533.
    """
    ct, st = np.cos(theta), np.sin(theta)
    c2, s2, cs = np.power(ct, 2), np.power(st, 2), ct * st
    Lx2, Ly2 = np.power(Lx, 2), np.power(Ly, 2)
    off = cs * (Lx2 - Ly2)
    return np.column_stack([c2 * Lx2 + s2 * Ly2, off, off, s2 * Lx2 + c2 * Ly2])


def det22(m: np.ndarray) -> np.ndarray:
    """Determinants of flattened 2x2 matrices; `_det_22_multi`, ellipse/covariance.py:543-547."""
    return m[..., 0] * m[..., 3] - m[..., 1] * m[..., 2]


def haversine_km(lat0, lon0, lat1, lon1):
    """asin-form haversine in km; `_haversine_multi`, ellipse/covariance.py:572-588."""
    dlon = lon0 - lon1
    dlat = lat0 - lat1
    a = np.sin(dlat / 2) ** 2 + np.cos(lat0) * np.cos(lat1) * np.sin(dlon / 2) ** 2
    return RADIUS_OF_EARTH_KM * (2 * np.arcsin(np.sqrt(a)))


def displacements_km(lat0, lon0, lat1, lon1, method: str):
    """(dy, dx) in km between point sets, radians in.

    Follows `_mo_disp_multi` / `_mod_mo_disp_multi`, ellipse/covariance.py:624-654: orientation
    (0) - (1), strict-inequality wrap of dx into [-pi, pi], and for "Modified_Met_Office" the
    mean-cos-latitude scaling applied AFTER the wrap.
    """
    dy = lat0 - lat1
    dx = lon0 - lon1
    dx = np.where(dx > np.pi, dx - TWO_PI, dx)
    dx = np.where(dx < -np.pi, dx + TWO_PI, dx)
    if method == "Modified_Met_Office":
        dx = dx * (0.5 * (np.cos(lat0) + np.cos(lat1)))
    elif method != "Met_Office":
        raise ValueError(f"Unknown 'delta_x_method' value: {method}")
    return RADIUS_OF_EARTH_KM * dy, RADIUS_OF_EARTH_KM * dx


class EllipseOracle:
    """Per-point state of the reference builder + pair / strip / full evaluators.

    Mirrors `EllipseCovarianceBuilder.__init__` / `_get_mask` (ellipse/covariance.py:95-180) and the
    pre-terms of `_calculate_covariance` (:218-224).  `precision` casts the inputs exactly where the
    reference does (:121-129, :219-221); float32 then runs float32 arithmetic end to end.
    """

    def __init__(self, Lx, Ly, theta, stdev, lats, lons, v,
                 delta_x_method="Modified_Met_Office", max_dist=None, precision=np.float64):
        if delta_x_method not in ("Modified_Met_Office", "Met_Office"):
            raise ValueError(f"Unknown 'delta_x_method' value: {delta_x_method}")
        self.v = v
        self.precision = precision
        self.method = delta_x_method
        self.max_dist = max_dist
        Lx = np.ma.MaskedArray(np.ma.asarray(Lx).astype(precision))
        mask = np.ma.getmaskarray(Lx)
        keep = ~mask.ravel()
        self.data_mask = mask
        comp = lambda a: np.ma.getdata(np.ma.asarray(a)).astype(precision).ravel()[keep]  # noqa: E731
        self.Lx, self.Ly, self.theta, self.stdev = (np.ma.getdata(Lx).ravel()[keep], comp(Ly),
                                                    comp(theta), comp(stdev))
        lats = np.asarray(lats).astype(precision)
        lons = np.asarray(lons).astype(precision)
        xg, yg = np.meshgrid(lons, lats)  # row-major (lat, lon) order, :166
        self.lat_deg = yg.ravel()[keep]
        self.lon_deg = xg.ravel()[keep]
        self.lat = np.deg2rad(self.lat_deg)  # :171-172
        self.lon = np.deg2rad(self.lon_deg)
        self.N = int(keep.sum())
        self.sigmas = sigma_rot(self.Lx, self.Ly, self.theta).astype(precision)  # :219-221
        self.sqrt_dets = np.sqrt(det22(self.sigmas))  # :222
        self.gamma_v_term = _gamma(v) * (2 ** (v - 1))  # :223
        self.sqrt_v_term = np.sqrt(v) * 2  # :224

    # -- the pair kernel -------------------------------------------------------------------
    def pairs(self, i_s, j_s) -> np.ndarray:
        """Covariance for index pairs; follows `c_ij_anisotropic_array`, :405-468, line by line."""
        i_s = np.asarray(i_s)
        j_s = np.asarray(j_s)
        dy, dx = displacements_km(self.lat[i_s], self.lon[i_s], self.lat[j_s], self.lon[j_s],
                                  self.method)
        c = (self.stdev[i_s] * self.stdev[j_s]) / self.gamma_v_term
        sb = 0.5 * (self.sigmas[i_s] + self.sigmas[j_s])
        sb_det = det22(sb)
        c = c * np.sqrt((self.sqrt_dets[i_s] * self.sqrt_dets[j_s]) / sb_det)
        inner = self.sqrt_v_term * np.sqrt(
            (dx * (dx * sb[..., 3] - dy * sb[..., 1]) + dy * (-dx * sb[..., 2] + dy * sb[..., 0]))
            / sb_det
        )
        c = c * np.power(inner, self.v)
        c = c * _kv(self.v, inner)
        c = c.astype(self.precision)
        if self.max_dist is not None:  # :256-266 (pairs beyond max_dist stay 0)
            d = haversine_km(self.lat[i_s], self.lon[i_s], self.lat[j_s], self.lon[j_s])
            c = np.where(d > self.max_dist, self.precision(0), c)
        return c

    def rows(self, r0: int, r1: int) -> np.ndarray:
        """Rows [r0, r1) of the full symmetric matrix, (r1-r0, N), diagonal = stdev^2 (:246)."""
        N = self.N
        out = np.empty((r1 - r0, N), dtype=self.precision)
        cols = np.arange(N)
        for r in range(r0, r1):
            lo = np.minimum(r, cols)
            hi = np.maximum(r, cols)
            sel = cols != r
            row = np.zeros(N, dtype=self.precision)
            row[sel] = self.pairs(lo[sel], hi[sel])  # always evaluated as (i<j), like :255
            row[r] = self.stdev[r] ** 2
            out[r - r0] = row
        return out

    def build(self, strip: int = 256) -> np.ndarray:
        """Full N x N matrix: upper-triangle strips, mirror (:277), then += diag(stdev^2) (:246)."""
        N = self.N
        C = np.zeros((N, N), dtype=self.precision)
        for r0 in range(0, N, strip):
            r1 = min(N, r0 + strip)
            ii, jj = np.meshgrid(np.arange(r0, r1), np.arange(r0, N), indexing="ij")
            up = jj > ii
            i_s, j_s = ii[up], jj[up]
            if i_s.size:
                C[i_s, j_s] = self.pairs(i_s, j_s)
        C = C + C.T
        C += np.diag(self.stdev ** 2).astype(self.precision)
        return C


def ellipse_covariance(Lx, Ly, theta, stdev, lats, lons, v, delta_x_method="Modified_Met_Office",
                       max_dist=None, precision=np.float64) -> np.ndarray:
    """`EllipseCovarianceBuilder(...).cov_ns` (ellipse/covariance.py:95-147)."""
    return EllipseOracle(Lx, Ly, theta, stdev, lats, lons, v, delta_x_method, max_dist,
                         precision).build()


def cov_2_cor(cov: np.ndarray) -> np.ndarray:
    """`utils.cov_2_cor`, utils.py:637-671 (unit diagonal forced, zeros kept)."""
    sd = np.sqrt(np.diag(cov))
    cor = cov / np.outer(sd, sd)
    if not np.all(np.diag(cor) == 1.0):
        bad = np.max(np.abs(np.diag(cor) - 1.0))
        if bad > 1e-6:
            raise ValueError("Correlation Diagonal contains values not close to 1. "
                             + f"With difference to 1: {bad}")
        np.fill_diagonal(cor, 1.0)
    cor[cov == 0] = 0
    return cor


def uncompress_cov(cov: np.ndarray, data_mask: np.ndarray, diag_fill_value=np.nan,
                   fill_value=np.nan) -> np.ndarray:
    """`EllipseCovarianceBuilder.uncompress_cov`, ellipse/covariance.py:475-509."""
    flat = np.asarray(data_mask).ravel()
    if int((~flat).sum()) != cov.shape[0]:
        raise ValueError("Data mask and coordinates cannot be aligned")
    M = flat.size
    out = np.full((M, M), fill_value, dtype=cov.dtype)
    out[np.arange(M), np.arange(M)] = diag_fill_value
    keep = np.flatnonzero(~flat)
    out[np.ix_(keep, keep)] = cov
    return out
