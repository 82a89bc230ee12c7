"""The oracle restatement against the committed golden vectors (CPU only).

Goldens come from the unmodified reference (tests/golden/make_golden.py) and from the reference's
own GeoStats.jl known-answer file.  Covariance: bit-exact in float64 (same arithmetic order, same
scipy entry points) - asserted at 1e-13 to allow for a different libm/scipy build on another host.
"""

import warnings

import numpy as np
import pytest
from conftest import from_triu, load_golden, max_norm_rel, max_rel

from glomargridding_b200 import synthetic as synth
from oracle import (EllipseOracle, extended_inverse, ordinary_kriging, simple_kriging,
                    stochastic_kriging)
from oracle.variogram import geostats_setup

SMALL = load_golden("ellipse_small.npz")


def _case(name):
    parts = name.split("_")
    v = float(parts[0][2:])
    meth = "Modified_Met_Office" if parts[1] == "mod" else "Met_Office"
    prec = np.float64 if parts[2] == "f64" else np.float32
    return v, meth, prec, "masked" in parts, (6000.0 if "maxdist" in parts else None)


@pytest.mark.parametrize("name", [str(c) for c in SMALL["cases"]])
def test_ellipse_small_full_matrix(name):
    v, meth, prec, masked, max_dist = _case(name)
    lats, lons = SMALL["lats"], SMALL["lons"]
    Lx, Ly, th, sd = synth.ellipse_fields(lats, lons, land_mask=masked)
    got = EllipseOracle(Lx, Ly, th, sd, lats, lons, v, meth, max_dist, prec).build()
    want = from_triu(SMALL[name])
    assert got.dtype == prec and got.shape == want.shape
    assert max_rel(got, want) <= (1e-13 if prec == np.float64 else 1e-6)
    assert np.array_equal(got, got.T)
    if max_dist is not None:
        assert (want == 0).sum() > 0 and np.array_equal(got == 0, want == 0)


@pytest.mark.parametrize("tag,res", [("5deg", 5.0), ("1deg", 1.0)])
def test_ellipse_sampled_pairs(tag, res):
    g = load_golden(f"ellipse_pairs_{tag}.npz")
    lats, lons = synth.global_grid(res)
    Lx, Ly, th, sd = synth.ellipse_fields(lats, lons)
    for name in [k for k in g.files if k.startswith("nu")]:
        v = float(name.split("_")[0][2:])
        meth = "Modified_Met_Office" if name.endswith("_mod") else "Met_Office"
        o = EllipseOracle(Lx, Ly, th, sd, lats, lons, v, meth)
        assert max_rel(o.pairs(g["i"], g["j"]), g[name]) <= 1e-13, name


def test_geostats_known_answer():
    """Reference test/test_kriging.py:67-192 (OrdinaryKriging vs GeoStats.jl), default allclose tol."""
    g = load_golden("geostats_ok.npz")
    cov, idx, obs = geostats_setup()
    assert np.array_equal(idx, g["idx"])
    with np.errstate(all="ignore"):
        res = ordinary_kriging(cov, idx, obs)
    assert np.allclose(g["expected"], res["field"].reshape(20, 20))


@pytest.mark.parametrize("tag", ["spd", "indef"])
def test_kriging_small(tag):
    g = load_golden("kriging_small.npz")
    cov = g[f"{tag}_cov"]
    assert (g[f"{tag}_min_eig"] > 0) == (tag == "spd")
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        sk = simple_kriging(cov, g["idx"], g["obs"], g["E_full"], mean=0.25)
        ok = ordinary_kriging(cov, g["idx"], g["obs"], g["E_full"])
        ok0 = ordinary_kriging(cov, g["idx"], g["obs"], None)
    assert np.array_equal(ok["idx"], g["kept_idx"]) and len(ok["idx"]) == len(g["idx"]) - 1
    for key, got in [("sk_field", sk["field"]), ("sk_unc", sk["uncertainty"]), ("sk_mask", sk["constraint_mask"]),
                     ("sk_weights", sk["weights"]), ("ok_field", ok["field"]), ("ok_unc", ok["uncertainty"]),
                     ("ok_mask", ok["constraint_mask"]), ("ok_weights", ok["weights"]),
                     ("ok_noE_field", ok0["field"]), ("ok_noE_unc", ok0["uncertainty"])]:
        assert max_norm_rel(got, g[f"{tag}_{key}"]) <= 1e-11, key


def test_kriging_diag_only_variant_matches():
    g = load_golden("kriging_small.npz")
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        a = ordinary_kriging(g["spd_cov"], g["idx"], g["obs"], g["E_full"], big_temporaries=True)
        b = ordinary_kriging(g["spd_cov"], g["idx"], g["obs"], g["E_full"], big_temporaries=False)
    for k in ("uncertainty", "constraint_mask"):
        assert max_norm_rel(a[k], b[k]) <= 1e-12


def test_stochastic_small_injected_draws():
    g = load_golden("kriging_small.npz")
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        st = stochastic_kriging(g["spd_cov"], g["idx"], g["obs"], g["E_full"], simulated_state=g["st_state"],
                                simulated_obs_noise=g["st_noise"])
    for key, got in [("st_result", st["result"]), ("st_epsilon", st["epsilon"]),
                     ("st_gridded_field", st["gridded_field"]), ("st_simulated_grid", st["simulated_grid"]),
                     ("st_unc", st["uncertainty"]), ("st_mask", st["constraint_mask"]),
                     ("st_ok_weights", st["weights"]), ("st_sk_weights", st["simple_weights"])]:
        assert max_norm_rel(got, g[key]) <= 1e-10, key


@pytest.mark.parametrize("n", [10, 25, 100])
def test_extended_inverse(n):
    g = load_golden("kriging_small.npz")
    S = g[f"ext_inv_{n}_S"]
    got = extended_inverse(np.linalg.inv(S))
    assert np.allclose(got, g[f"ext_inv_{n}"])
    S_ext = np.block([[S, np.ones((n, 1))], [np.ones((1, n)), 0]])
    assert np.allclose(got, np.linalg.inv(S_ext))


def test_variogram_oracle_vs_reference_golden():
    """oracle/variogram.py against the four reference variogram models (tests/golden/variogram_small.npz)."""
    import importlib.util
    import os

    from oracle import variogram as ov

    spec = importlib.util.spec_from_file_location("make_golden_cases", os.path.join(os.path.dirname(__file__), "golden",
                                                                                   "variogram_cases.py"))
    cases = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(cases)
    g = load_golden("variogram_small.npz")
    dist = g["dist"]
    for k, (cls, p) in enumerate(cases.VARIOGRAM_CASES):
        rng = p.get("range")
        if rng is None:
            rng = p["effective_range"] / (2 if cls in ("GaussianVariogram", "MaternVariogram") else 3 if cls == "ExponentialVariogram" else 1)
        with np.errstate(all="ignore"):
            if cls == "MaternVariogram":
                got = ov.matern_variogram(dist, p["psill"], p["nugget"], rng, p["nu"], p["method"])
            elif cls == "GaussianVariogram":
                got = ov.gaussian_variogram(dist, p["psill"], p["nugget"], rng)
            elif cls == "ExponentialVariogram":
                got = ov.exponential_variogram(dist, p["psill"], p["nugget"], rng)
            else:
                got = ov.spherical_variogram(dist, p["psill"], p["nugget"], rng)
        assert np.allclose(got, g[f"case{k}"], rtol=1e-13, atol=1e-15), (k, cls)


def test_notebook_2014_03_field_chain():
    """BASELINE configs[0]: 5-degree ellipse covariance -> eigenvalue clip -> simple / ordinary kriging of the real
    HadSST.4 2014-03 field with the real HadCRUT.5 error covariance; golden = the unmodified reference's outputs."""
    from conftest import load_notebook_case
    from oracle.covariance_tools import explained_variance_clip

    g, E = load_notebook_case()
    lats, lons = synth.global_grid(5.0)
    Lx, Ly, th, sd = synth.ellipse_fields(lats, lons)
    raw = EllipseOracle(Lx, Ly, th, sd, lats, lons, 1.5, "Modified_Met_Office").build()
    cov, kept, w = explained_variance_clip(raw)
    assert abs(w.min() - float(g["cov_min_eig_raw"])) < 1e-9          # indefinite before the clip
    assert max_norm_rel(np.diag(cov), g["cov_diag_clipped"]) <= 1e-10
    idx, obs = g["idx"], g["obs"]
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        sk = simple_kriging(cov, idx, obs, E)
        ok = ordinary_kriging(cov, idx, obs, E)
    for got, tag in ((sk, "sk"), (ok, "ok")):
        assert max_norm_rel(got["field"], g[f"{tag}_field"]) <= 1e-8
        assert max_norm_rel(got["uncertainty"] ** 2, g[f"{tag}_unc"] ** 2) <= 1e-8
        assert max_norm_rel(got["constraint_mask"], g[f"{tag}_mask"]) <= 1e-8


def test_eigenvalue_clip_oracle_vs_reference_golden():
    """oracle/covariance_tools.py against the reference's `eigenvalue_clip` outputs (tests/golden/clip_small.npz)."""
    from oracle.covariance_tools import explained_variance_clip, laloux_clip

    g = load_golden("clip_small.npz")
    cov = from_triu(g["cov"])
    scale = np.abs(cov).max()
    assert np.max(np.abs(explained_variance_clip(cov)[0] - from_triu(g["ev95"]))) <= 1e-11 * scale
    assert np.max(np.abs(explained_variance_clip(cov, 0.8)[0] - from_triu(g["ev80"]))) <= 1e-11 * scale
    assert np.max(np.abs(laloux_clip(cov, None, 40) - from_triu(g["laloux_t40"]))) <= 1e-11 * scale
    assert np.max(np.abs(laloux_clip(cov, cov.shape[0], 400) - from_triu(g["laloux_t400"]))) <= 1e-11 * scale


def test_geometry_helpers_oracle_vs_reference_golden():
    """Module-level geometry helpers (`ellipse/covariance.py:512-654`): the oracle's restatements are the same numpy
    expressions, so they must reproduce the reference's outputs bit for bit - including the strict +-pi wrap."""
    from oracle import ellipse as oe

    g = load_golden("geometry_small.npz")
    sig = oe.sigma_rot(g["Lx"], g["Ly"], g["theta"])
    assert np.array_equal(sig, g["sigmas"])
    assert np.array_equal(oe.det22(sig), g["dets"])
    pts = (g["lat0"], g["lon0"], g["lat1"], g["lon1"])
    assert np.array_equal(oe.haversine_km(*pts), g["haversine"])
    assert np.array_equal(np.stack(oe.displacements_km(*pts, "Met_Office")), g["mo"])
    assert np.array_equal(np.stack(oe.displacements_km(*pts, "Modified_Met_Office")), g["mod_mo"])
    # the scalar variants agree with the vector ones except for the coincident-point shortcut (`:561-562`)
    assert np.array_equal(g["single_mo"], g["mo"][:, -8:].T) and np.array_equal(g["single_mod_mo"], g["mod_mo"][:, -8:].T)
    assert np.allclose(g["single_haversine"], g["haversine"][-8:], rtol=1e-15, atol=0)


def _atlantic_case():
    from conftest import from_triu

    g = load_golden("ellipse_atlantic.npz")
    mask = g["Lx"] > 1e5                                       # test/test_ellipse.py:196
    fields = [np.ma.masked_where(mask, g[k]) for k in ("Lx", "Ly", "theta", "stdev")]
    return g, fields, from_triu(g["cov_triu"])


def test_reference_own_ellipse_golden_atlantic():
    """The reference's own known-answer test for the covariance path (`test/test_ellipse.py:177-217`): fitted ellipse
    parameters of an Atlantic window (`test/data/Atlantic_Ocean_07.nc`), nu = 0.5, default float32 precision, expected
    covariance `test/data/cov_no_hfix.nc` - same tolerance as the reference (`np.allclose(rtol=1e-5)`), in float32 like
    the reference's default and in float64."""
    from oracle.ellipse import cov_2_cor, ellipse_covariance

    g, fields, expected = _atlantic_case()
    assert float(g["matern_nu"]) == 0.5 and expected.shape == (262, 262)
    for precision in (np.float32, np.float64):
        cov = ellipse_covariance(*fields, g["lats"], g["lons"], 0.5, precision=precision)
        assert cov.shape == expected.shape and cov.dtype == precision
        assert np.allclose(cov, expected, rtol=1e-5)
        cor = cov_2_cor(np.asarray(cov, dtype=np.float64))
        assert np.isclose(1, np.max(np.diag(cor)))              # :215-217
    assert np.allclose(g["cor_diag"], 1.0)


# ---- SURVEY 8f row N4 (ellipse-parameter fit): the oracle is in place and pinned; the device port is the next step ----
def _fit_cases():
    from oracle.ellipse_fit import EllipseModelOracle

    g = load_golden("ellipse_fit.npz")
    for c in range(int(g["n_cases"])):
        aniso, rot, pd, unit = (bool(x) for x in g[f"c{c}_meta"])
        yield c, g, EllipseModelOracle(aniso, rot, pd, float(g[f"c{c}_v"]), unit_sigma=unit)


def test_ellipse_fit_correlation_and_likelihood_vs_reference_golden():
    """`cov_ij_anisotropic` / `cov_ij_isotropic` (model.py:582-686) and `negative_log_likelihood` (model.py:266-384)
    against the unmodified reference's outputs on the same displacements / correlations."""
    for c, g, m in _fit_cases():
        X, y, true = g[f"c{c}_X"], g[f"c{c}_y"], list(g[f"c{c}_true"])
        assert np.allclose(m.cov_ij(X, true), g[f"c{c}_cov"], rtol=1e-13, atol=0)
        p_true = true + ([] if m.unit_sigma else [0.05])
        p_guess = list(m.default_guesses) + ([] if m.unit_sigma else [0.1])
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            got = [m.negative_log_likelihood(X, y.copy(), p_true), m.negative_log_likelihood(X, y.copy(), p_guess)]
        assert np.allclose(got, g[f"c{c}_nll"], rtol=1e-12, atol=0), (c, got, g[f"c{c}_nll"])


def test_ellipse_fit_nelder_mead_vs_reference_golden():
    """`EllipseModel.fit(estimate_SE=None)` (model.py:402-500): same optimiser, same likelihood - same optimum."""
    for c, g, m in _fit_cases():
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            res, se, bounds = m.fit(g[f"c{c}_X"], g[f"c{c}_y"].copy())
        assert se is None and np.array_equal(np.array(bounds), g[f"c{c}_bounds"])
        assert np.allclose(res.x, g[f"c{c}_fit_x"], rtol=1e-6, atol=1e-9), (c, res.x, g[f"c{c}_fit_x"])
        assert np.isclose(res.fun, g[f"c{c}_fit"][0], rtol=1e-9)
        assert bool(res.success) == bool(g[f"c{c}_fit"][3])


def test_ellipse_fit_oracle_vs_live_reference(reference):
    """Random parameters and displacements against the live reference (skipped where its sources are absent)."""
    import importlib

    from oracle import ellipse_fit as ef

    model = importlib.import_module("glomar_gridding.ellipse.model")
    dist = importlib.import_module("glomar_gridding.distances")
    rng = np.random.default_rng(8)
    for _ in range(50):
        v = float(rng.choice([0.5, 1.5, 2.2, 3.7]))
        Lx, Ly, th = rng.uniform(300, 3000), rng.uniform(300, 3000), rng.uniform(-np.pi, np.pi)
        dx, dy = rng.uniform(-2000, 2000, 40), rng.uniform(-2000, 2000, 40)
        for theta in (th, None):
            assert np.array_equal(ef.mahalanobis(dx, dy, Lx, Ly, theta), dist.mahal_dist_func(dx, dy, Lx, Ly, theta=theta))
            assert np.array_equal(ef.cov_ij_anisotropic(v, 1, dx, dy, Lx, Ly, theta=theta),
                                  model.cov_ij_anisotropic(v, 1, dx, dy, Lx, Ly, theta=theta))
        assert np.array_equal(ef.cov_ij_isotropic(v, 1, np.hypot(dx, dy), Lx), model.cov_ij_isotropic(v, 1, np.hypot(dx, dy), Lx))
        assert np.array_equal(ef.sigma_rot(Lx, Ly, th), dist.sigma_rot_func(Lx, Ly, th))
    with pytest.raises(ValueError, match="isotropic rotated"):
        ef.EllipseModelOracle(False, True, True, 1.5)
    with pytest.raises(ValueError, match="must be > 0"):
        ef.EllipseModelOracle(True, True, True, 0.0)
