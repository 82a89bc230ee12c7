"""Parity of the CUDA ellipse-covariance path against the oracle and the reference goldens.

Tolerance: 1e-10 entrywise relative (BASELINE.json north_star) with the 1e-290 absolute floor of
SURVEY.md §8(d).  All calls go through the Python classes, i.e. through the C ABI.
"""

import numpy as np
import pytest
from conftest import from_triu, load_golden, max_rel

from glomargridding_b200 import synthetic as synth

pytestmark = pytest.mark.gpu

TOL = 1e-10


@pytest.fixture(scope="module")
def Builder():
    import torch

    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from glomargridding_b200.ellipse import EllipseCovarianceBuilder

    return EllipseCovarianceBuilder


SMALL = load_golden("ellipse_small.npz")


def _case(name):
    parts = name.split("_")
    return (float(parts[0][2:]), "Modified_Met_Office" if parts[1] == "mod" else "Met_Office",
            np.float64 if parts[2] == "f64" else np.float32, "masked" in parts,
            6000.0 if "maxdist" in parts else None)


@pytest.mark.parametrize("name", [str(c) for c in SMALL["cases"]])
def test_small_grid_vs_reference_golden(Builder, name):
    v, meth, prec, masked, max_dist = _case(name)
    lats, lons = SMALL["lats"], SMALL["lons"]
    Lx, Ly, th, sd = synth.ellipse_fields(lats, lons, land_mask=masked)
    b = Builder(Lx, Ly, th, sd, lats, lons, v=v, delta_x_method=meth, max_dist=max_dist, precision=prec)
    got, want = b.cov_ns, from_triu(SMALL[name])
    assert got.dtype == prec and got.shape == want.shape
    assert np.array_equal(got, got.T)
    if prec == np.float64:
        assert max_rel(got, want) <= TOL
    else:
        # the reference's own float32 criterion (test/test_ellipse.py:212)
        assert np.allclose(got, want, rtol=1e-5)
    if max_dist is not None:
        assert np.array_equal(got == 0, want == 0)
    assert np.array_equal(np.diag(got), (b.stdev_compressed ** 2).astype(prec))


@pytest.mark.parametrize("mode", ["table", "direct"])
@pytest.mark.parametrize("v", [0.5, 1.5, 2.5, 2.2, 0.3, 7.25])
def test_matern_modes_agree_with_oracle(Builder, v, mode):
    from oracle import EllipseOracle

    lats, lons = synth.global_grid(15.0)
    Lx, Ly, th, sd = synth.ellipse_fields(lats, lons)
    got = Builder(Lx, Ly, th, sd, lats, lons, v=v, precision=np.float64, matern_mode=mode).cov_ns
    want = EllipseOracle(Lx, Ly, th, sd, lats, lons, v).build()
    assert max_rel(got, want) <= TOL


@pytest.mark.parametrize("v,meth", [(1.5, "Modified_Met_Office"), (2.2, "Modified_Met_Office"),
                                    (0.5, "Met_Office"), (2.2, "Met_Office")])
def test_5deg_full_matrix_vs_oracle(Builder, v, meth):
    from oracle import EllipseOracle

    lats, lons = synth.global_grid(5.0)
    Lx, Ly, th, sd = synth.ellipse_fields(lats, lons)
    got = Builder(Lx, Ly, th, sd, lats, lons, v=v, delta_x_method=meth, precision=np.float64).cov_ns
    want = EllipseOracle(Lx, Ly, th, sd, lats, lons, v, meth).build()
    assert got.shape == (2592, 2592)
    assert max_rel(got, want) <= TOL
    assert np.array_equal(got, got.T)


@pytest.mark.parametrize("tag,res", [("5deg", 5.0), ("1deg", 1.0)])
def test_sampled_pairs_vs_reference_golden(Builder, tag, res):
    """`c_ij_anisotropic_array(i_s, j_s)` on the pairs the reference itself evaluated (incl. +-pi wrap ties)."""
    g = load_golden(f"ellipse_pairs_{tag}.npz")
    lats, lons = synth.global_grid(res)
    Lx, Ly, th, sd = synth.ellipse_fields(lats, lons)
    for name in [k for k in g.files if k.startswith("nu")]:
        v = float(name.split("_")[0][2:])
        meth = "Modified_Met_Office" if name.endswith("_mod") else "Met_Office"
        b = Builder(Lx, Ly, th, sd, lats, lons, v=v, delta_x_method=meth, precision=np.float64, rows=(0, 64))
        got = b.c_ij_anisotropic_array(g["i"], g["j"])
        assert max_rel(got, g[name]) <= TOL, name


def test_1deg_row_blocks_vs_golden_pairs_and_symmetry(Builder):
    """Full-size (N = 64 800) build in row blocks: entries sampled by the reference, block symmetry,
    exact diagonal - the size-independent checks of SURVEY.md §8(d)."""
    import torch

    g = load_golden("ellipse_pairs_1deg.npz")
    lats, lons = synth.global_grid(1.0)
    Lx, Ly, th, sd = synth.ellipse_fields(lats, lons)
    N = 64800
    gi, gj, want = g["i"].astype(np.int64), g["j"].astype(np.int64), g["nu2.2_mod"]
    checked = 0
    for row0, nrows in [(0, 4096), (30016, 4096), (64800 - 4000, 4000)]:
        b = Builder(Lx, Ly, th, sd, lats, lons, v=2.2, precision=np.float64, rows=(row0, nrows))
        blk = b.cov_ns_device
        assert blk.shape == (nrows, N)
        sq = blk[:, row0:row0 + nrows]
        assert torch.equal(sq, sq.T)                       # mirrored tiles are bit-identical
        sd_c = torch.from_numpy(b.stdev_compressed[row0:row0 + nrows]).cuda()
        assert torch.equal(torch.diagonal(sq), sd_c * sd_c)
        for a, c in ((gi, gj), (gj, gi)):                   # both orientations of every sampled pair
            sel = (a >= row0) & (a < row0 + nrows)
            got = blk[torch.from_numpy(a[sel] - row0).cuda(), torch.from_numpy(c[sel]).cuda()].cpu().numpy()
            assert max_rel(got, want[sel]) <= TOL
            checked += int(sel.sum())
    assert checked > 3000


def test_row_block_equals_full_build(Builder):
    lats, lons = synth.global_grid(10.0)
    Lx, Ly, th, sd = synth.ellipse_fields(lats, lons, land_mask=True)
    full = Builder(Lx, Ly, th, sd, lats, lons, v=2.2, precision=np.float64).cov_ns
    N = full.shape[0]
    for row0, nrows in [(0, 64), (64, 100), (128, N - 128), (70, 33), (N - 1, 1), (128, 192), (0, 200), (256, 100),
                        (192, 130)]:
        blk = Builder(Lx, Ly, th, sd, lats, lons, v=2.2, precision=np.float64, rows=(row0, nrows)).cov_ns
        assert np.array_equal(blk, full[row0:row0 + nrows])


def test_api_errors_and_attributes(Builder):
    lats, lons = synth.global_grid(20.0)
    Lx, Ly, th, sd = synth.ellipse_fields(lats, lons, land_mask=True)
    with pytest.raises(ValueError, match="max_dist must be a number"):
        Builder(Lx, Ly, th, sd, lats, lons, v=1.5, max_dist="far")
    with pytest.raises(ValueError, match="Unknown 'delta_x_method' value"):
        Builder(Lx, Ly, th, sd, lats, lons, v=1.5, delta_x_method="nope")
    with pytest.raises(ValueError, match="Unknown covariance_method"):
        Builder(Lx, Ly, th, sd, lats, lons, v=1.5, covariance_method="nope")
    with pytest.raises(ValueError, match="batch_size must be set"):
        Builder(Lx, Ly, th, sd, lats, lons, v=1.5, covariance_method="batched")
    b = Builder(Lx, Ly, th, sd, lats, lons, v=0.5, covariance_method="batched", batch_size=100)
    n = int((~np.ma.getmaskarray(Lx)).sum())
    assert b.cov_ns.dtype == np.float32 and b.cov_ns.shape == (n, n) and b.covar_size == n
    assert b.xy_shape == Lx.shape and b.n_elements == Lx.size
    assert b.sigmas.shape == (n, 4) and b.sqrt_dets.shape == (n,)
    b.calculate_cor()
    assert np.isclose(1, np.max(np.diag(b.cor_ns)))       # reference test/test_ellipse.py:215-217
    b.uncompress_cov()
    assert b.cov_ns.shape == (Lx.size, Lx.size)            # reference test/test_ellipse.py:387-388
    assert np.isnan(b.cov_ns).sum() > 0


def test_haversine_distance_matrix_vs_oracle(Builder):
    """reference test/test_distances.py:33-50 (Halifax - Southampton within 1 km) and the asin-form oracle"""
    from glomargridding_b200.distances import haversine_distance_from_frame, radial_dist
    from oracle.ellipse import haversine_km

    rng = np.random.default_rng(5)
    lat, lon = rng.uniform(-90, 90, 300), rng.uniform(-180, 180, 300)
    got = haversine_distance_from_frame({"lat": lat, "lon": lon})
    la, lo = np.radians(lat), np.radians(lon)
    want = haversine_km(la[:, None], lo[:, None], la[None, :], lo[None, :])
    assert got.shape == (300, 300) and np.max(np.abs(got - want)) <= 1e-9 and np.allclose(np.diag(got), 0)
    d = haversine_distance_from_frame({"lat": [44.6476, 50.9097], "lon": [-63.5728, -1.4043]})
    assert abs(d[0, 1] - 4557) < 1 and abs(radial_dist(44.6476, -63.5728, 50.9097, -1.4043) - d[0, 1]) < 1e-6
    assert np.allclose(haversine_distance_from_frame({"lat": lat[:5], "lon": lon[:5]}, radius=1.0) * 6371, got[:5, :5])
    with pytest.raises(ValueError, match="only contain 'lat' and 'lon'"):
        haversine_distance_from_frame({"lon": lon, "lat": lat})


@pytest.mark.parametrize("prec", [np.float64, np.float32])
def test_calculate_cor_and_uncompress_on_device_vs_oracle(Builder, prec):
    """`calculate_cor` (utils.cov_2_cor, reference utils.py:637-671) and `uncompress_cov` (ellipse/covariance.py:475-509)
    run as CUDA kernels: bit-equal to the numpy restatement (same IEEE operations in the matrix' own precision)."""
    import torch

    from glomargridding_b200.utils import cor_2_cov, cov_2_cor
    from oracle.ellipse import cov_2_cor as cor_oracle
    from oracle.ellipse import uncompress_cov as unc_oracle

    lats, lons = synth.global_grid(20.0)
    Lx, Ly, th, sd = synth.ellipse_fields(lats, lons, land_mask=True)
    b = Builder(Lx, Ly, th, sd, lats, lons, v=1.5, precision=prec, max_dist=9000.0)     # max_dist: exact zeros
    cov = b.cov_ns.copy()
    assert cov.dtype == prec and (cov == 0).sum() > 0
    b.calculate_cor()
    want = cor_oracle(cov)
    assert b.cor_ns.dtype == prec and b.cor_ns_device.is_cuda
    assert np.array_equal(b.cor_ns, want)
    assert np.all(np.diag(b.cor_ns) == 1) and ((b.cor_ns == 0) == (cov == 0)).all()
    # round trip through the functional forms, numpy in / numpy out and tensor in / tensor out
    back = cor_2_cov(b.cor_ns, np.diag(cov))
    assert back.dtype == prec and np.array_equal(back, want * np.outer(np.sqrt(np.diag(cov)), np.sqrt(np.diag(cov))))
    t = cov_2_cor(torch.from_numpy(cov).cuda())
    assert isinstance(t, torch.Tensor) and t.is_cuda and np.array_equal(t.cpu().numpy(), want)
    for fills in ((np.nan, np.nan), (1.0, 0.0)):
        b2 = Builder(Lx, Ly, th, sd, lats, lons, v=1.5, precision=prec, max_dist=9000.0)
        b2.uncompress_cov(*fills)
        got = b2.cov_ns
        ref = unc_oracle(cov, b2.data_mask, *fills)
        assert got.shape == (Lx.size, Lx.size) and got.dtype == prec
        assert np.array_equal(got, ref, equal_nan=True)
    with pytest.raises(ValueError, match="Correlation Diagonal"):
        cov_2_cor(np.diag([1.0, 0.0, 4.0]))
    bad = Builder(Lx, Ly, th, sd, lats, lons, v=1.5, precision=prec)
    bad.data_mask = np.zeros_like(np.asarray(bad.data_mask))
    with pytest.raises(ValueError, match="cannot be aligned"):
        bad.uncompress_cov()


def test_geometry_helpers_vs_reference_golden(Builder):
    """`_sigma_rot_func_multi`, `_det_22_*`, `_haversine_*`, `_mo_disp_*`, `_mod_mo_disp_*` (reference
    `ellipse/covariance.py:512-654`; its own checks: `test/test_matrix.py:74-140`) as device kernels against the
    unmodified reference's outputs.  Products and sums are rounded like numpy's; sin / cos / asin are CUDA's (1-2 ulp),
    hence 1e-13 relative; the wrap of the longitude difference at exactly +-pi must fall on the same side."""
    import glomargridding_b200.ellipse.covariance as ec

    g = load_golden("geometry_small.npz")
    sig = ec._sigma_rot_func_multi(g["Lx"], g["Ly"], g["theta"])
    assert sig.shape == g["sigmas"].shape and sig.dtype == np.float64
    assert np.max(np.abs(sig - g["sigmas"])) <= 1e-13 * np.max(np.abs(g["sigmas"]))
    assert np.array_equal(sig[:, 1], sig[:, 2])
    assert np.array_equal(ec._det_22_multi(g["sigmas"]), g["dets"])                    # exact: two products, one difference
    assert np.array_equal([ec._det_22_single(g["sigmas"][i]) for i in range(8)], g["single_det"])
    pts = (g["lat0"], g["lon0"], g["lat1"], g["lon1"])
    hav = ec._haversine_multi(*pts)
    assert np.allclose(hav, g["haversine"], rtol=1e-12, atol=1e-9)
    for fn, key in ((ec._mo_disp_multi, "mo"), (ec._mod_mo_disp_multi, "mod_mo")):
        dy, dx = fn(*pts)
        assert np.array_equal(dy, g[key][0])                                           # one subtraction, one product
        assert np.allclose(dx, g[key][1], rtol=1e-14, atol=1e-11)
        assert np.array_equal(np.sign(dx), np.sign(g[key][1]))                         # same side of the wrap
    assert np.array_equal(ec._mo_disp_multi(*pts)[1], g["mo"][1])                      # no transcendental: exact
    for k in range(8):
        args = tuple(float(p[k - 8]) for p in pts)
        assert np.isclose(ec._haversine_single(*args), g["single_haversine"][k], rtol=1e-12, atol=1e-9)
        assert np.allclose(ec._mod_mo_disp_single(*args), g["single_mod_mo"][k], rtol=1e-14, atol=1e-11)
        assert np.array_equal(ec._mo_disp_single(*args), g["single_mo"][k])
    # float32 inputs give float32 results like the numpy expressions do
    assert ec._haversine_multi(*[p.astype(np.float32) for p in pts]).dtype == np.float32


def test_strategy_methods_leave_the_off_diagonal_matrix(Builder):
    """`calculate_covariance_array / _loop / _batched` (reference `:250-403`) fill `cov_ns` WITHOUT the diagonal
    (`_calculate_covariance` adds diag(stdev^2) afterwards, `:246`); the three strategies agree exactly (reference
    test/test_ellipse.py:300-353 allows rtol 1e-5 between them)."""
    lats, lons = synth.global_grid(20.0)
    Lx, Ly, th, sd = synth.ellipse_fields(lats, lons, land_mask=True)
    b = Builder(Lx, Ly, th, sd, lats, lons, v=1.5, precision=np.float64, batch_size=100)
    full = b.cov_ns.copy()
    off = full - np.diag(np.diag(full))
    assert np.array_equal(np.diag(full), b.stdev_compressed ** 2)
    for name in ("calculate_covariance_array", "calculate_covariance_loop", "calculate_covariance_batched"):
        getattr(b, name)()
        assert np.array_equal(b.cov_ns, off), name
    b.batch_size = None
    with pytest.raises(ValueError, match="batch_size must be set"):
        b.calculate_covariance_batched()
    b._get_disp_fn()
    assert b.disp_fn.__name__ == "_mod_mo_disp_multi"
    b.covariance_method, b.delta_x_method = "low_memory", "Met_Office"
    b._get_disp_fn()
    assert b.disp_fn.__name__ == "_mo_disp_single"


def test_reference_own_golden_atlantic_window(Builder):
    """The reference's own known-answer test for this path, `test/test_ellipse.py:177-217`: fitted ellipse parameters
    of an Atlantic window, nu = 0.5, DEFAULT precision (float32), expected covariance from `test/data/cov_no_hfix.nc`
    (fixture: tests/golden/make_golden_atlantic.py) - with the reference's own tolerance, then the correlation check
    of `:213-217` and the uncompress check of `test_ellipse_covariance_rescale` (`:356-388`)."""
    g = load_golden("ellipse_atlantic.npz")
    expected = from_triu(g["cov_triu"])
    mask = g["Lx"] > 1e5
    fields = [np.ma.masked_where(mask, g[k]) for k in ("Lx", "Ly", "theta", "stdev")]
    b = Builder(*fields, g["lats"], g["lons"], v=0.5)
    assert b.cov_ns.dtype == np.float32 and b.cov_ns.shape == expected.shape
    assert np.allclose(b.cov_ns, expected, rtol=1e-5)
    b64 = Builder(*fields, g["lats"], g["lons"], v=0.5, precision=np.float64)
    assert np.allclose(b64.cov_ns, expected, rtol=1e-5)
    b.calculate_cor()
    assert hasattr(b, "cor_ns") and np.isclose(1, np.max(np.diag(b.cor_ns)))
    b.uncompress_cov()
    assert b.cov_ns.shape[0] == b.cov_ns.shape[1] == g["Lx"].shape[0] ** 2


# ---- 1 degree, full rows against the oracle (parity at the headline size) -----------------
def _rows_1deg():
    """Rows of the 64 800-point matrix that the tile kernel treats differently: both sides of every row-block
    boundary of a 2 / 4 / 8-rank build (bench.row_block: 128-row macro tiles), macro-tile and tile edges, both poles,
    the equator, the ragged last tile, and one row for EACH of the 360 longitudes (so that every longitude meets its
    180-degree partner column - the 12 pairs whose difference is one ulp away from pi included - in both
    orientations, since full rows hold the entries left and right of the diagonal)."""
    N, T = 64800, 128
    tiles = (N + T - 1) // T
    rows = {0, 1, 63, 64, 127, 128, 359, 360, N - 1, N - 2, N - 64, N - 65, N - 360, N - 361, 32399, 32400, 32401}
    for world in (2, 4, 8):
        for r in range(1, world):
            b = (tiles * r // world) * T
            rows.update({b - 1, b, b + 1})
    for lon in range(360):
        rows.add(((lon * 7) % 180) * 360 + lon)
    return np.array(sorted(rows), dtype=np.int64)


def _oracle_rows(o, rows):
    from concurrent.futures import ThreadPoolExecutor

    with ThreadPoolExecutor(max_workers=8) as ex:       # numpy / scipy ufuncs release the GIL
        return np.concatenate(list(ex.map(lambda r: o.rows(int(r), int(r) + 1), rows)), axis=0)


def test_1deg_full_rows_vs_oracle_f64(Builder):
    """Tile kernel, general nu (the headline configuration): ~400 full rows = 2.6e7 entries at 1e-10."""
    import torch
    from oracle import EllipseOracle

    lats, lons = synth.global_grid(1.0)
    Lx, Ly, th, sd = synth.ellipse_fields(lats, lons)
    rows = _rows_1deg()
    assert rows.size >= 256
    C = Builder(Lx, Ly, th, sd, lats, lons, v=2.2, precision=np.float64).cov_ns_device
    got = C[torch.from_numpy(rows).cuda()].cpu().numpy()
    got_cols = C[:, torch.from_numpy(rows[:32]).cuda()].T.cpu().numpy()      # the mirrored entries of 32 of them
    del C
    torch.cuda.empty_cache()
    want = _oracle_rows(EllipseOracle(Lx, Ly, th, sd, lats, lons, 2.2), rows)
    assert got.size >= 1.6e7
    assert max_rel(got, want) <= TOL
    assert np.array_equal(got_cols, got[:32])               # C[j, i] is bit-equal to C[i, j]


def test_1deg_full_rows_vs_oracle_f32_default_precision(Builder):
    """The builder's DEFAULT precision (float32, like the reference) at 1 degree, the reference's own criterion
    `np.allclose(rtol=1e-5)` (test/test_ellipse.py:212): float32 inputs, float32 wrap decisions."""
    import torch
    from oracle import EllipseOracle

    lats, lons = synth.global_grid(1.0)
    Lx, Ly, th, sd = synth.ellipse_fields(lats, lons)
    rows = _rows_1deg()[::4]
    C = Builder(Lx, Ly, th, sd, lats, lons, v=1.5).cov_ns_device
    assert C.dtype == torch.float32
    got = C[torch.from_numpy(rows).cuda()].cpu().numpy()
    del C
    torch.cuda.empty_cache()
    want = _oracle_rows(EllipseOracle(Lx, Ly, th, sd, lats, lons, 1.5, precision=np.float32), rows)
    assert want.dtype == np.float32 and np.allclose(got, want, rtol=1e-5)


def test_1deg_max_dist_rows_vs_oracle(Builder):
    import torch
    from oracle import EllipseOracle

    lats, lons = synth.global_grid(1.0)
    Lx, Ly, th, sd = synth.ellipse_fields(lats, lons)
    rows = _rows_1deg()[::8]
    for row0 in (0, 32384):                                 # two row blocks that hold some of the rows
        sel = rows[(rows >= row0) & (rows < row0 + 8192)]
        b = Builder(Lx, Ly, th, sd, lats, lons, v=0.5, max_dist=3000.0, precision=np.float64, rows=(row0, 8192))
        got = b.cov_ns_device[torch.from_numpy(sel - row0).cuda()].cpu().numpy()
        want = _oracle_rows(EllipseOracle(Lx, Ly, th, sd, lats, lons, 0.5, max_dist=3000.0), sel)
        assert np.array_equal(got == 0, want == 0)
        assert max_rel(got, want) <= TOL
