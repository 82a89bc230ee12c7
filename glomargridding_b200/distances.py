"""Distance helpers on the hot path (drop-ins for the array-producing part of `glomar_gridding.distances`).

Only what SURVEY.md §2 marks as on the path: `haversine_distance_from_frame` (`distances.py:191-218`, the
sklearn asin-form haversine matrix times a radius - evaluated by `gmg_haversine_f64`) and the scalar `radial_dist`
(`:97-136`).  The Transverse-Mercator / geopandas helpers are out of scope.
"""

from __future__ import annotations

import numpy as np
import torch

from . import _device as dev
from ._lib import check, lib


def haversine_distance(lats_deg, lons_deg, radius: float = 6371.0, return_device: bool = False):
    """Pairwise haversine distances (N x N) between points given in decimal degrees, on the GPU."""
    lat = dev.to_device(np.radians(np.asarray(lats_deg, dtype=np.float64)), torch.float64)
    lon = dev.to_device(np.radians(np.asarray(lons_deg, dtype=np.float64)), torch.float64)
    if lat.shape != lon.shape or lat.dim() != 1:
        raise ValueError("lats and lons must be 1-d arrays of equal length")
    N = lat.numel()
    D = torch.empty((N, N), dtype=torch.float64, device=lat.device)
    for r0 in range(0, N, 65535):
        nrows = min(65535, N - r0)
        check(lib().gmg_haversine_f64(lat.data_ptr(), lon.data_ptr(), N, r0, nrows, float(radius),
                                      D[r0:].data_ptr(), N, dev.stream_ptr()), "gmg_haversine_f64")
    return D if return_device else dev.to_host(D)


def haversine_distance_from_frame(df, radius: float = 6371) -> np.ndarray:
    """Great-circle distance matrix between the rows of a frame with exactly the columns "lat" and "lon"
    (decimal degrees); reference `distances.py:191-218`.  Any frame type with `.columns` and `df["lat"]`
    (polars, pandas) or a dict of arrays is accepted."""
    cols = list(df.keys()) if isinstance(df, dict) else list(df.columns)
    if cols != ["lat", "lon"]:
        raise ValueError("Input must only contain 'lat' and 'lon' columns")
    return haversine_distance(np.asarray(df["lat"]), np.asarray(df["lon"]), radius)


def radial_dist(lat1: float, lon1: float, lat2: float, lon2: float) -> float:
    """Great-circle distance (atan2 form) between two points in km, R = 6371.0 (reference `distances.py:97-136`)."""
    R = 6371.0
    lat1r, lat2r = np.radians(lat1), np.radians(lat2)
    dlon, dlat = np.radians(lon2 - lon1), lat2r - lat1r
    a = np.sin(dlat / 2) ** 2 + np.cos(lat1r) * np.cos(lat2r) * np.sin(dlon / 2) ** 2
    return R * (2 * np.arctan2(np.sqrt(a), np.sqrt(1 - a)))
