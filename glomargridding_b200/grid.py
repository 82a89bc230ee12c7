"""Grid helpers of the kriging path (drop-in for the index-producing part of `glomar_gridding.grid`).

`map_to_grid` (reference `grid.py:28-124`) assigns every observation to its nearest grid point and produces the flat,
row-major `grid_idx` that `SimpleKriging` / `OrdinaryKriging` / `StochasticKriging` take as `idx`; the nearest-point
search and the index arithmetic run on the device (`gmg_nearest_index_f64`).  `grid_from_resolution`
(`grid.py:127-195`) builds the coordinate vectors.

The reference works on `polars.DataFrame` / `xarray.DataArray`; neither is required here: `obs` may be a
`pandas.DataFrame`, a `polars.DataFrame` (when polars is installed) or a dict of equal-length arrays, `grid` anything
with `.coords[name].values` (an `xarray.DataArray`) or a dict {coordinate name: 1-d array} / `Grid`.  The result has the
type of `obs`, with the columns the reference adds (`grid_idx`, `grid_<coord>`), sorted by `grid_idx` like the reference.
"""

from __future__ import annotations

from collections.abc import Iterable

import numpy as np
import torch

from . import _device as dev
from ._lib import check, lib


class Grid(dict):
    """{coordinate name: 1-d coordinate vector}, in axis order; `shape` as `xarray.DataArray.shape`."""

    @property
    def shape(self) -> tuple[int, ...]:
        return tuple(len(v) for v in self.values())


def grid_from_resolution(resolution, bounds: list[tuple[float, float]], coord_names: list[str]) -> Grid:
    """Coordinate vectors `np.arange(lower, upper, resolution)` per axis (reference `grid.py:127-195`; the reference
    wraps them in an empty `xarray.DataArray`)."""
    if not isinstance(resolution, Iterable):
        resolution = [resolution for _ in range(len(bounds))]
    if len(resolution) != len(coord_names) or len(bounds) != len(coord_names):
        raise ValueError("Input lists must have the same length")
    return Grid((name, np.arange(lo, hi, res)) for name, (lo, hi), res in zip(coord_names, bounds, resolution))


def _grid_coords(grid, names) -> list[np.ndarray]:
    if hasattr(grid, "coords") and not isinstance(grid, dict):
        return [np.asarray(grid.coords[n].values, dtype=np.float64) for n in names]
    return [np.asarray(grid[n], dtype=np.float64) for n in names]


def nearest_flat_index(grid_axes: list[np.ndarray], positions: list[np.ndarray]):
    """(flat row-major index, [per-axis index]) of the grid point nearest to every position - on the device."""
    n = len(np.asarray(positions[0]))
    shape = [len(a) for a in grid_axes]
    device = dev.require_cuda()
    flat = torch.zeros(n, dtype=torch.int64, device=device)
    per_axis = []
    for d, (axis, pos) in enumerate(zip(grid_axes, positions)):
        stride = int(np.prod(shape[d + 1:], dtype=np.int64))
        a, p = dev.to_device(axis, torch.float64), dev.to_device(np.asarray(pos, dtype=np.float64), torch.float64)
        one = torch.zeros(n, dtype=torch.int64, device=device)
        check(lib().gmg_nearest_index_f64(a.data_ptr(), a.numel(), p.data_ptr(), n, one.data_ptr(), 1, 0,
                                          dev.stream_ptr()), "gmg_nearest_index_f64")
        check(lib().gmg_nearest_index_f64(a.data_ptr(), a.numel(), p.data_ptr(), n, flat.data_ptr(), stride,
                                          int(d > 0), dev.stream_ptr()), "gmg_nearest_index_f64")
        per_axis.append(one.cpu().numpy())
    return flat.cpu().numpy(), per_axis


def map_to_grid(obs, grid, obs_coords: list[str] = ["lat", "lon"], grid_coords: list[str] = ["latitude", "longitude"],
                sort: bool = True, bounds: list[tuple[float, float]] | None = None, add_grid_pts: bool = True,
                grid_prefix: str = "grid_"):
    """Align observations to a grid (reference `grid.py:28-124`): every observation gets the flat row-major index of
    its nearest grid point (`grid_idx`) and, with `add_grid_pts`, that point's coordinates (`grid_<coord>`); rows are
    sorted by `grid_idx`.  `bounds` first restricts the grid (inclusive coordinate slices, like `xarray.sel`) and the
    observations (lower bound inclusive, upper exclusive, like the reference's `filter_bounds`)."""
    kind = "dict"
    if hasattr(obs, "to_pandas") and not hasattr(obs, "iloc"):         # polars
        kind, frame = "polars", obs.to_pandas()
    elif hasattr(obs, "iloc"):
        kind, frame = "pandas", obs.copy()
    else:
        frame = {k: np.asarray(v) for k, v in obs.items()}
    axes = _grid_coords(grid, grid_coords)
    cols = {k: np.asarray(frame[k]) for k in (frame.columns if kind != "dict" else frame.keys())}
    if bounds is not None:
        if len(bounds) != len(obs_coords):
            raise ValueError("Length of 'bounds' must equal length of 'bound_cols'")
        axes = [a[(a >= lo) & (a <= hi)] for a, (lo, hi) in zip(axes, bounds)]
        keep = np.ones(len(cols[obs_coords[0]]), dtype=bool)
        for c, (lo, hi) in zip(obs_coords, bounds):
            keep &= (cols[c] >= lo) & (cols[c] < hi)
        cols = {k: v[keep] for k, v in cols.items()}
    flat, per_axis = nearest_flat_index(axes, [cols[c] for c in obs_coords])
    cols[grid_prefix + "idx"] = flat
    if add_grid_pts:
        for axis, ai, c in zip(axes, per_axis, obs_coords):
            cols[grid_prefix + c] = axis[ai]
    if sort:
        order = np.argsort(cols["grid_idx"], kind="stable")            # the reference sorts by the literal "grid_idx"
        cols = {k: v[order] for k, v in cols.items()}
    if kind == "dict":
        return cols
    import pandas as pd

    out = pd.DataFrame(cols)
    if kind == "polars":
        import polars as pl

        return pl.from_pandas(out)
    return out
