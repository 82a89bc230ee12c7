"""Generate the golden fixtures in this directory from the UNMODIFIED reference.

Run in the build container (needs `/root/reference`; numpy 2.3.5 / scipy 1.18.1 there):

    python tests/golden/make_golden.py

Every array is produced by the reference's own classes (imported from source through
`oracle/reference_shim.py`) on the deterministic synthetic inputs of
`glomargridding_b200/synthetic.py`; the GeoStats.jl vector is the reference's own known-answer
file (`test/data/geostatsjl_ord_krig_results.dat`, 400 floats) re-packed as .npz.  The fixtures are
what pins the oracle (and through it the CUDA path) on the GPU box, where the reference is absent.
"""

from __future__ import annotations

import contextlib
import io
import os
import sys
import warnings

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from glomargridding_b200 import synthetic as synth  # noqa: E402
from oracle.reference_shim import REFERENCE_ROOT, load_reference  # noqa: E402

ref = load_reference()


def quiet(fn, *a, **k):
    with contextlib.redirect_stdout(io.StringIO()), warnings.catch_warnings():
        warnings.simplefilter("ignore")
        return fn(*a, **k)


class _PairsOnly(ref.EllipseCovarianceBuilder):
    """Reference builder state without the full build: only the pre-terms of
    `_calculate_covariance` (ellipse/covariance.py:218-224), so that the reference's own
    `c_ij_anisotropic_array` can be called on sampled index pairs at 1 degree."""

    def _calculate_covariance(self):
        cov_mod = ref.ellipse_covariance
        self._get_disp_fn()
        self.sigmas = cov_mod._sigma_rot_func_multi(
            self.Lx_compressed, self.Ly_compressed, self.theta_compressed
        ).astype(self.precision)
        self.sqrt_dets = np.sqrt(cov_mod._det_22_multi(self.sigmas))
        self.gamma_v_term = cov_mod.gamma(self.v) * (2 ** (self.v - 1))
        self.sqrt_v_term = np.sqrt(self.v) * 2


def triu(a):
    iu = np.triu_indices(a.shape[0], k=0)
    return a[iu]


def make_geostats():
    with open(os.path.join(REFERENCE_ROOT, "test", "data", "geostatsjl_ord_krig_results.dat")) as f:
        vals = np.array([float(line) for line in f])
    expected = np.reshape(vals, (20, 20), "F")  # Julia is column-major (test/test_kriging.py:53-61)
    np.savez_compressed(os.path.join(HERE, "geostats_ok.npz"), expected=expected,
                        idx=np.array([84, 289, 194]), obs=np.array([1.0, 0.0, 1.0]))


def make_ellipse_small():
    """Full matrices (upper triangle incl. diagonal) on a 20-degree global grid (9 x 18 cells)."""
    lats, lons = synth.global_grid(20.0)
    out = {"lats": lats, "lons": lons}
    cases = [
        ("nu0.5_mod_f64", 0.5, "Modified_Met_Office", np.float64, False, None),
        ("nu1.5_mod_f64", 1.5, "Modified_Met_Office", np.float64, False, None),
        ("nu2.5_mod_f64", 2.5, "Modified_Met_Office", np.float64, False, None),
        ("nu2.2_mod_f64", 2.2, "Modified_Met_Office", np.float64, False, None),
        ("nu2.2_mo_f64", 2.2, "Met_Office", np.float64, False, None),
        ("nu0.8_mod_f64", 0.8, "Modified_Met_Office", np.float64, False, None),
        ("nu3.7_mo_f64", 3.7, "Met_Office", np.float64, False, None),
        ("nu1.0_mod_f64", 1.0, "Modified_Met_Office", np.float64, False, None),
        ("nu2.0_mo_f64", 2.0, "Met_Office", np.float64, False, None),
        ("nu2.2_mod_f64_masked", 2.2, "Modified_Met_Office", np.float64, True, None),
        ("nu1.5_mod_f64_maxdist", 1.5, "Modified_Met_Office", np.float64, False, 6000.0),
        ("nu0.5_mod_f32", 0.5, "Modified_Met_Office", np.float32, False, None),
        ("nu2.2_mod_f32_masked", 2.2, "Modified_Met_Office", np.float32, True, None),
    ]
    names = []
    for name, v, meth, prec, masked, max_dist in cases:
        Lx, Ly, th, sd = synth.ellipse_fields(lats, lons, land_mask=masked)
        b = quiet(ref.EllipseCovarianceBuilder, Lx, Ly, th, sd, lats, lons, v=v, delta_x_method=meth,
                  max_dist=max_dist, precision=prec)
        out[name] = triu(b.cov_ns)
        names.append(name)
    out["cases"] = np.array(names)
    np.savez_compressed(os.path.join(HERE, "ellipse_small.npz"), **out)


def sample_pairs(N, nlon, rng, n_random, n_near, n_tie):
    i = rng.integers(0, N, n_random)
    j = rng.integers(0, N, n_random)
    # near-diagonal / adjacent cells (small Bessel arguments)
    a = rng.integers(0, N - nlon - 1, n_near)
    off = rng.choice([1, 2, nlon, nlon + 1, nlon - 1], n_near)
    # longitudes exactly 180 degrees apart (wrap ties, SURVEY.md §7 "hard parts")
    t = rng.integers(0, N, n_tie)
    ti = (t // nlon) * nlon + (t % nlon)
    tj = (rng.integers(0, N // nlon, n_tie)) * nlon + ((t % nlon) + nlon // 2) % nlon
    i = np.concatenate([i, a, ti])
    j = np.concatenate([j, a + off, tj])
    keep = i != j
    i, j = i[keep], j[keep]
    lo, hi = np.minimum(i, j), np.maximum(i, j)
    return lo.astype(np.int64), hi.astype(np.int64)


def make_ellipse_pairs():
    """Sampled (i<j) pairs from the reference's own `c_ij_anisotropic_array` at 5 and 1 degree."""
    for res, tag, n_random in [(5.0, "5deg", 12000), (1.0, "1deg", 20000)]:
        lats, lons = synth.global_grid(res)
        Lx, Ly, th, sd = synth.ellipse_fields(lats, lons)
        rng = np.random.default_rng(int(res * 100) + 17)
        i_s, j_s = sample_pairs(lats.size * lons.size, lons.size, rng, n_random, 3000, 3000)
        out = {"i": i_s.astype(np.int32), "j": j_s.astype(np.int32)}
        for name, v, meth in [("nu2.2_mod", 2.2, "Modified_Met_Office"), ("nu1.5_mod", 1.5, "Modified_Met_Office"),
                              ("nu0.5_mo", 0.5, "Met_Office"), ("nu2.5_mo", 2.5, "Met_Office"),
                              ("nu0.3_mod", 0.3, "Modified_Met_Office"), ("nu3.0_mo", 3.0, "Met_Office")]:
            b = quiet(_PairsOnly, Lx, Ly, th, sd, lats, lons, v=v, delta_x_method=meth, precision=np.float64)
            out[name] = b.c_ij_anisotropic_array(i_s, j_s)
        np.savez_compressed(os.path.join(HERE, f"ellipse_pairs_{tag}.npz"), **out)


def make_kriging_small():
    """SimpleKriging / OrdinaryKriging / StochasticKriging outputs of the reference on a 20-degree grid,
    for an SPD covariance ("Met_Office") and for the indefinite default one ("Modified_Met_Office")."""
    lats, lons = synth.global_grid(20.0)
    M = lats.size * lons.size
    Lx, Ly, th, sd = synth.ellipse_fields(lats, lons)
    idx, obs = synth.fixed_observations(seed=11, n_cells=M, n_obs=60)
    E_full = synth.full_grid_error_cov(idx, M, lons.size, seed=7)
    # make one observed cell carry a zero error variance so the filter path (kriging.py:99-118) runs
    E_full[idx[5], idx[5]] = 0.0
    out = {"idx": idx, "obs": obs, "E_full": E_full}
    for tag, meth in [("spd", "Met_Office"), ("indef", "Modified_Met_Office")]:
        cov = quiet(ref.EllipseCovarianceBuilder, Lx, Ly, th, sd, lats, lons, v=1.5, delta_x_method=meth,
                    precision=np.float64).cov_ns
        out[f"{tag}_cov"] = cov
        out[f"{tag}_min_eig"] = np.linalg.eigvalsh(cov).min()
        sk = quiet(ref.SimpleKriging, cov, idx.copy(), obs.copy(), E_full.copy())
        out[f"{tag}_sk_field"] = quiet(sk.solve, 0.25)
        out[f"{tag}_sk_unc"] = quiet(sk.get_uncertainty)
        out[f"{tag}_sk_mask"] = quiet(sk.constraint_mask)
        out[f"{tag}_sk_weights"] = sk.kriging_weights
        ok = quiet(ref.OrdinaryKriging, cov, idx.copy(), obs.copy(), E_full.copy())
        out[f"{tag}_ok_field"] = quiet(ok.solve)
        out[f"{tag}_ok_unc"] = quiet(ok.get_uncertainty)
        out[f"{tag}_ok_mask"] = quiet(ok.constraint_mask)
        out[f"{tag}_ok_weights"] = ok.kriging_weights
        out["kept_idx"] = ok.idx
        # no error covariance at all
        ok0 = quiet(ref.OrdinaryKriging, cov, idx.copy(), obs.copy())
        out[f"{tag}_ok_noE_field"] = quiet(ok0.solve)
        out[f"{tag}_ok_noE_unc"] = quiet(ok0.get_uncertainty)
        if tag == "spd":
            st = quiet(ref.StochasticKriging, cov, idx.copy(), obs.copy(), E_full.copy())
            rng = np.random.default_rng(99)
            state = np.linalg.cholesky(cov) @ rng.standard_normal(M)
            np.random.seed(20140371)
            res = quiet(st.solve, simulated_state=state)
            out["st_state"] = state
            out["st_noise"] = st.simulated_obs - state[st.idx]
            out["st_result"] = res
            out["st_epsilon"] = st.epsilon
            out["st_gridded_field"] = st.gridded_field
            out["st_simulated_grid"] = st.simulated_grid
            out["st_unc"] = quiet(st.get_uncertainty)
            out["st_mask"] = quiet(st.constraint_mask)
            out["st_ok_weights"] = st.kriging_weights
            out["st_sk_weights"] = st.simple_kriging_weights
    for n in (10, 25, 100):  # test_inverse_trick, test/test_kriging.py:297-316
        np.random.seed(31900)
        A = np.random.rand(n, n)
        S = A @ A.T
        out[f"ext_inv_{n}_S"] = S
        out[f"ext_inv_{n}"] = ref.kriging._extended_inverse(np.linalg.inv(S))
    np.savez_compressed(os.path.join(HERE, "kriging_small.npz"), **out)


def make_clip_small():
    """`eigenvalue_clip` of the reference (both methods) on the indefinite default-method 20-degree covariance."""
    lats, lons = synth.global_grid(20.0)
    Lx, Ly, th, sd = synth.ellipse_fields(lats, lons)
    cov = quiet(ref.EllipseCovarianceBuilder, Lx, Ly, th, sd, lats, lons, v=1.5, precision=np.float64).cov_ns
    out = {"cov": triu(cov), "min_eig": np.linalg.eigvalsh(cov).min()}
    clip = ref.covariance_tools.eigenvalue_clip
    out["ev95"] = triu(quiet(clip, cov))
    out["ev80"] = triu(quiet(clip, cov, method="explained_variance", target_variance_fraction=0.8))
    out["laloux_t40"] = triu(quiet(clip, cov, method="Laloux_2000", num_time_pts=40))
    out["laloux_t400"] = triu(quiet(clip, cov, method="Laloux_2000", num_time_pts=400, num_grid_pts=cov.shape[0]))
    w = np.linalg.eigvalsh(cov / np.sqrt(np.outer(np.diag(cov), np.diag(cov))))
    out["laloux_kept"] = np.array([int((w > (1 + np.sqrt(cov.shape[0] / t)) ** 2).sum()) for t in (40, 400)])
    np.savez_compressed(os.path.join(HERE, "clip_small.npz"), **out)


from variogram_cases import VARIOGRAM_CASES  # noqa: E402  (tests/golden/variogram_cases.py)


def variogram_points():
    """10 x 10 degree grid of reference test/test_variogram.py (lat 0..9, lon 0..9), row-major (lat, lon)."""
    lat, lon = np.meshgrid(np.arange(10.0), np.arange(10.0), indexing="ij")
    return lat.ravel(), lon.ravel()


def make_variogram_small():
    """The four variogram models of the reference on the haversine distance matrix of a 10 x 10 grid - the
    parameter sets of reference test/test_variogram.py:22-110 plus one general-nu Matern."""
    from sklearn.metrics.pairwise import haversine_distances

    lat, lon = variogram_points()
    dist = haversine_distances(np.radians(np.column_stack([lat, lon]))) * 6371.0   # reference distances.py:218
    out = {"dist": dist}
    for k, (cls, params) in enumerate(VARIOGRAM_CASES):
        with np.errstate(all="ignore"):
            out[f"case{k}"] = quiet(getattr(ref.variogram, cls)(**params).fit, dist.copy())
    np.savez_compressed(os.path.join(HERE, "variogram_small.npz"), **out)


if __name__ == "__main__":
    if len(sys.argv) > 1:                      # e.g. `make_golden.py make_clip_small`: regenerate single fixtures
        for name in sys.argv[1:]:
            globals()[name]()
        sys.exit(0)
    make_clip_small()
    make_variogram_small()
    make_geostats()
    make_ellipse_small()
    make_ellipse_pairs()
    make_kriging_small()
    for f in sorted(os.listdir(HERE)):
        if f.endswith(".npz"):
            print(f, os.path.getsize(os.path.join(HERE, f)) // 1024, "KiB")
