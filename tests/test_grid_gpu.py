"""`grid.map_to_grid` (reference grid.py:28-124): the observation -> grid index production on the device against the
reference's own expressions (`utils.find_nearest`: `int(np.argmin(np.abs(array - value)))` per coordinate, then
`np.ravel_multi_index(..., order="C")`) - ties (a value exactly between two grid points takes the first), values
outside the grid, bounds, sorting - and the chain map_to_grid -> OrdinaryKriging."""

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def G():
    import torch

    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import glomargridding_b200.grid as g

    return g


def _reference_index(axes, positions):
    idx = [[int(np.argmin(np.abs(a - v))) for v in p] for a, p in zip(axes, positions)]     # utils.find_nearest
    return np.ravel_multi_index(idx, [len(a) for a in axes], order="C"), idx


def test_map_to_grid_vs_reference_expressions(G):
    rng = np.random.default_rng(5)
    grid = G.grid_from_resolution(5, [(-87.5, 90), (-177.5, 180)], ["lat", "lon"])
    assert grid.shape == (36, 72)
    n = 5000
    lat = rng.uniform(-95, 95, n)                              # some outside the grid: nearest edge point
    lon = rng.uniform(-185, 185, n)
    lat[:40] = np.arange(-90, 110, 5)[:40]                      # exactly between two grid latitudes: first one wins
    lon[:40] = np.arange(-180, 20, 5)[:40]
    obs = {"lat": lat, "lon": lon, "sst": rng.standard_normal(n)}
    out = G.map_to_grid(obs, grid, grid_coords=["lat", "lon"], sort=False)
    want, per = _reference_index([grid["lat"], grid["lon"]], [lat, lon])
    assert np.array_equal(out["grid_idx"], want)
    assert np.array_equal(out["grid_lat"], grid["lat"][per[0]]) and np.array_equal(out["grid_lon"], grid["lon"][per[1]])
    srt = G.map_to_grid(obs, grid, grid_coords=["lat", "lon"])
    order = np.argsort(want, kind="stable")
    assert np.array_equal(srt["grid_idx"], want[order]) and np.array_equal(srt["sst"], obs["sst"][order])
    # bounds: grid restricted by inclusive coordinate slices, observations by [lower, upper)
    bnd = [(-42.5, 47.5), (-100.0, 20.0)]
    sub = G.map_to_grid(obs, grid, grid_coords=["lat", "lon"], bounds=bnd, sort=False, add_grid_pts=False)
    keep = (lat >= bnd[0][0]) & (lat < bnd[0][1]) & (lon >= bnd[1][0]) & (lon < bnd[1][1])
    ax = [grid["lat"][(grid["lat"] >= bnd[0][0]) & (grid["lat"] <= bnd[0][1])],
          grid["lon"][(grid["lon"] >= bnd[1][0]) & (grid["lon"] <= bnd[1][1])]]
    want_sub, _ = _reference_index(ax, [lat[keep], lon[keep]])
    assert np.array_equal(sub["grid_idx"], want_sub) and "grid_lat" not in sub
    # pandas in -> pandas out
    import pandas as pd

    df = G.map_to_grid(pd.DataFrame(obs), grid, grid_coords=["lat", "lon"])
    assert list(df.columns) == ["lat", "lon", "sst", "grid_idx", "grid_lat", "grid_lon"]
    assert np.array_equal(df["grid_idx"].to_numpy(), want[order])


def test_find_nearest_vs_live_reference(G, reference):
    rng = np.random.default_rng(1)
    axis = np.sort(rng.uniform(-90, 90, 57))
    vals = np.concatenate([rng.uniform(-100, 100, 400), axis[:5], 0.5 * (axis[3:8] + axis[4:9])])
    idx_ref, _ = reference.utils.find_nearest(axis, vals)
    flat, per = G.nearest_flat_index([axis], [vals])
    assert np.array_equal(flat, np.array(idx_ref)) and np.array_equal(per[0], flat)


def test_map_to_grid_into_kriging(G):
    """The index the mapping produces is the `idx` of the kriging classes (notebook cells 28-31)."""
    from glomargridding_b200 import synthetic as synth
    from glomargridding_b200.ellipse import EllipseCovarianceBuilder
    from glomargridding_b200.kriging import OrdinaryKriging

    lats, lons = synth.global_grid(10.0)
    Lx, Ly, th, sd = synth.ellipse_fields(lats, lons)
    C = EllipseCovarianceBuilder(Lx, Ly, th, sd, lats, lons, v=1.5, delta_x_method="Met_Office",
                                 precision=np.float64).cov_ns_device
    rng = np.random.default_rng(3)
    cells = np.sort(rng.choice(lats.size * lons.size, 120, replace=False))
    obs = {"lat": lats[cells // lons.size] + rng.uniform(-4, 4, 120), "lon": lons[cells % lons.size] + rng.uniform(-4, 4, 120),
           "val": rng.standard_normal(120)}
    m = G.map_to_grid(obs, {"latitude": lats, "longitude": lons})
    assert np.array_equal(m["grid_idx"], cells)
    z = OrdinaryKriging(C, m["grid_idx"], m["val"], np.diag(np.full(120, 0.04))).solve()
    assert z.shape == (lats.size * lons.size,) and np.isfinite(z).all()
