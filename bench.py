#!/usr/bin/env python
"""Benchmark of the hot path (BASELINE.json metric: "ellipse-cov entries/s @1deg grid; kriged months/s").

Primary line value: 1-degree global (N = 64 800) non-stationary ellipse Matern covariance build, general nu via
Bessel-K (BASELINE.json configs[2], the configuration "ellipse-cov entries/s @1deg grid" is quoted on), row-block
sharded across the GPUs of one box.  The second half of the metric, "kriged months/s" (configs[1]: 480 synthetic
months of 5-degree ordinary kriging, sharded by month), is measured in the same run and reported under
`"kriging"` in the same JSON line, with its own roofline, e2e and cpu_baseline objects.  So are configs[3] -
1-degree ordinary kriging of 20 000 observations against the row-block sharded covariance (`"kriging_1deg"`:
distributed Cholesky + column-sharded forward solve) - and configs[4] - the 200-member 5-degree stochastic
ensemble (`"stochastic"`).  The FP64 peaks the fractions are taken against (DFMA pipe, DMMA) are measured in the
run by `gmg_microbench_fp64` (`"fp64_peaks"`).

    python bench.py --gpus N --steps K --warmup W            # this implementation
    python bench.py --impl reference --gpus N --steps K ...   # the reference's CPU algorithm (oracle port)

One JSON line on stdout (rank 0).  See DESIGN.md "Measurement" for what each key means.
"""

from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time
import warnings

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "ellipse-cov entries/s @1deg grid; kriged months/s"
UNIT = "entries/s"
TILE = 128                  # row blocks start on a macro-tile (2 x 2 tiles of 64) boundary
# Fallbacks only: the run measures both with gmg_microbench_fp64 ("fp64_peaks" in the line) and divides by those.
DMMA_PEAK_TFLOPS = 37.1     # mma.sync m8n8k4 f64, measured on this pool's B200 (profiles/r01_microbench_fp64.json)
DFMA_PEAK_TFLOPS = 34.2     # plain DFMA, same file
# ALGORITHMIC FP64 work per evaluated pair of the tile kernel ("the ncu-measured dfma / dadd / dmul
# instruction count of the shipped kernel, which is then frozen as the algorithmic figure"): frozen at the round-1
# kernel's count - 50 FP64 pipe instructions per pair for general nu (degree-7 Matern table), 43 for the closed forms
# (profiles/r01_kernels.md).  Each holds one FP64 issue slot and is counted as one FMA slot (2 flop)
# against the DFMA rate measured in the run.  The kernel shipped now executes K1_FP64_EXECUTED of them (degree-5
# table with 32 sub-intervals per octave, round 2; 94 per two pairs in the SASS of the row loop): doing the same job with fewer instructions shows up as time, not as a smaller numerator.
K1_FP64_INSTR_PER_PAIR = {"table": 50.0, "closed": 43.0}
K1_FP64_EXECUTED = {"table": 47.0, "closed": 43.0}


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--res", type=float, default=1.0, help="grid resolution in degrees (1.0 = headline config)")
    ap.add_argument("--nu", type=float, default=2.2)
    ap.add_argument("--e2e-steps", type=int, default=2)
    ap.add_argument("--cpu-rows", type=int, default=768, help="rows of the 1-degree matrix the CPU baseline evaluates")
    ap.add_argument("--months", type=int, default=480)
    ap.add_argument("--month-steps", type=int, default=3)
    ap.add_argument("--cpu-months", type=int, default=12)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-kriging", action="store_true")
    ap.add_argument("--no-kriging-1deg", action="store_true")
    ap.add_argument("--no-stochastic", action="store_true")
    ap.add_argument("--no-ellipse-fit", action="store_true")
    ap.add_argument("--no-notebook", action="store_true", help="skip the configs[0] leg (HadSST 2014-03 notebook field)")
    ap.add_argument("--notebook-steps", type=int, default=5)
    ap.add_argument("--no-fp64-peaks", action="store_true", help="skip the in-run DFMA / DMMA microbenchmark (ncu passes)")
    ap.add_argument("--fit-res", type=float, default=5.0, help="grid of the training cube of the ellipse-fit workload")
    ap.add_argument("--fit-steps", type=int, default=5)
    ap.add_argument("--cpu-fit-cells", type=int, default=48, help="cells the CPU arm of the ellipse fit fits")
    ap.add_argument("--nobs-1deg", type=int, default=20000)
    ap.add_argument("--krige-1deg-steps", type=int, default=3)
    ap.add_argument("--members", type=int, default=200)
    ap.add_argument("--stochastic-steps", type=int, default=20)
    ap.add_argument("--cpu-cols-1deg", type=int, default=1024,
                    help="grid columns the CPU arm of the 1-degree kriging solves for (scaled to 64 800)")
    ap.add_argument("--ref-batches", type=int, default=2, help="2M-pair batches of the reference's own batched loop per step")
    return ap.parse_args()


def workload(res, nu):
    from glomargridding_b200 import synthetic as synth

    lats, lons = synth.global_grid(res)
    Lx, Ly, theta, stdev = synth.ellipse_fields(lats, lons)
    return {"lats": lats, "lons": lons, "Lx": Lx, "Ly": Ly, "theta": theta, "stdev": stdev, "nu": nu,
            "N": lats.size * lons.size}


def month_problems(months, n_cells, nlon):
    """Synthetic months of SURVEY.md §8(d): month m has default_rng(m) observations and a HadCRUT-style n x n
    error covariance (12 longitude-band blocks + diagonal)."""
    from glomargridding_b200 import synthetic as synth

    out = []
    for m in months:
        idx, obs = synth.month_observations(m, n_cells)
        out.append((idx, obs, synth.hadcrut_style_error_cov(idx, nlon, seed=m)))
    return out


def row_block(N, rank, world):
    """Contiguous row block of rank `rank`, aligned to the 128-row macro tile so mirrored tiles stay inside it."""
    tiles = (N + TILE - 1) // TILE
    t0 = tiles * rank // world
    t1 = tiles * (rank + 1) // world
    return t0 * TILE, min(N, t1 * TILE) - t0 * TILE


def config_dict(args, wl, world):
    return {"workload": f"{args.res:g}deg global grid ({wl['N']} cells) ellipse Matern covariance, nu={args.nu:g} "
                        f"(general nu via Bessel-K), Modified_Met_Office, FP64 [BASELINE configs[2]]; secondary: "
                        f"5deg grid ordinary kriging of {args.months} synthetic months [configs[1]]; 1deg ordinary "
                        f"kriging of {args.nobs_1deg} observations against the sharded covariance [configs[3]]; "
                        f"5deg stochastic kriging, {args.members} members [configs[4]]; {args.fit_res:g}deg ellipse-"
                        f"parameter fit of every cell [SURVEY 8f N4]",
            "N": wl["N"], "nu": args.nu, "sharding": f"row blocks x{world}; months round-robin x{world}",
            "l2": "output (8*N^2/world bytes per GPU) is far larger than the 126 MB L2; no flush needed"}


# ------------------------------------------------------------------------------------------------
# CPU arm: the oracle port of the reference's algorithm on a bounded sample of the same workload
# ------------------------------------------------------------------------------------------------
def cpu_cov_sample(wl, n_rows):
    """Time the reference's pair formula (oracle port of `c_ij_anisotropic_array`) on every i<j pair whose row
    index lies in a contiguous block of `n_rows` rows in the middle of the matrix.  The reference evaluates
    N(N-1)/2 pairs to fill N^2 entries, so entries/s = 2 * pairs/s."""
    from oracle import EllipseOracle

    N = wl["N"]
    o = EllipseOracle(wl["Lx"], wl["Ly"], wl["theta"], wl["stdev"], wl["lats"], wl["lons"], wl["nu"])
    r0 = (N // 2 // TILE) * TILE
    rows = np.arange(r0, min(N, r0 + n_rows))
    t0 = time.perf_counter()
    pairs = 0
    for r in rows:  # strips keep the temporaries small, like the reference's "batched" strategy
        j = np.arange(r + 1, N)
        if j.size:
            o.pairs(np.full(j.size, r), j)
            pairs += j.size
    dt = time.perf_counter() - t0
    return 2.0 * pairs / dt, pairs, dt


def reference_ns():
    """The UNMODIFIED reference package (staged under oracle/_ref by __graft_entry__.build(), or /root/reference in
    the build container) imported through the stub shim - or None, in which case the CPU legs time the oracle port."""
    try:
        from oracle.reference_shim import load_reference, reference_available

        return load_reference() if reference_available() else None
    except Exception:
        return None


_REF_BUILDER = {}


def ref_cov_sample(ref, wl, n_batches, batch_size=2_000_000):
    """The reference's OWN covariance build, bounded: the statements of `calculate_covariance_batched`
    (ellipse/covariance.py:382-401) on the unmodified class - `itertools.batched(combinations(range(N), 2),
    batch_size)` pair generation, `np.asarray(batch).T`, `c_ij_anisotropic_array`, scatter - for the first
    `n_batches` batches.  (The method itself cannot be called at 1 degree: it allocates the 33.6 GB matrix and mirrors
    it with a second one.)  `combinations` yields rows 0, 1, ... first, so the sample is the leading rows of the
    matrix and the scatter goes into a buffer of just those rows.  entries/s = 2 * pairs/s like the port."""
    import contextlib
    import io
    from itertools import combinations, islice

    cov_mod = ref.ellipse_covariance
    key = (wl["N"], wl["nu"])
    if key not in _REF_BUILDER:
        class PreTermsOnly(ref.EllipseCovarianceBuilder):
            def _calculate_covariance(self):              # pre-terms of :218-224, nothing else
                self._get_disp_fn()
                self.sigmas = cov_mod._sigma_rot_func_multi(
                    self.Lx_compressed, self.Ly_compressed, self.theta_compressed).astype(self.precision)
                self.sqrt_dets = np.sqrt(cov_mod._det_22_multi(self.sigmas))
                self.gamma_v_term = cov_mod.gamma(self.v) * (2 ** (self.v - 1))
                self.sqrt_v_term = np.sqrt(self.v) * 2

        with contextlib.redirect_stdout(io.StringIO()):
            _REF_BUILDER[key] = PreTermsOnly(wl["Lx"], wl["Ly"], wl["theta"], wl["stdev"], wl["lats"], wl["lons"],
                                             v=wl["nu"], precision=np.float64, covariance_method="batched",
                                             batch_size=batch_size)
    b = _REF_BUILDER[key]
    N = wl["N"]
    want = n_batches * batch_size
    per_row = N - 1 - np.arange(N)
    rows = int(np.searchsorted(np.cumsum(per_row), want) + 1)
    block = np.zeros((rows, N), dtype=b.precision)
    pairs = 0
    t0 = time.perf_counter()
    for batch in islice(cov_mod.batched(combinations(range(N), 2), b.batch_size), n_batches):
        i_s, j_s = np.asarray(batch).T
        loop_c_ij = b.c_ij_anisotropic_array(i_s, j_s)
        block[i_s, j_s] = loop_c_ij.astype(b.precision)
        pairs += i_s.size
    dt = time.perf_counter() - t0
    return 2.0 * pairs / dt, pairs, dt


def cpu_kriging_sample(n_months, ref=None):
    """The reference's per-month workflow: OrdinaryKriging construct + solve + get_uncertainty on the first `n_months`
    synthetic months - the unmodified classes when `ref` is given, the oracle port (M x M temporaries included like
    the reference) otherwise."""
    from glomargridding_b200 import synthetic as synth
    from oracle import EllipseOracle, ordinary_kriging

    use_all_host_threads()
    lats, lons = synth.global_grid(5.0)
    Lx, Ly, th, sd = synth.ellipse_fields(lats, lons)
    cov = EllipseOracle(Lx, Ly, th, sd, lats, lons, 1.5, "Met_Office").build()
    probs = month_problems(range(n_months), cov.shape[0], lons.size)
    t0 = time.perf_counter()
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        for idx, obs, E in probs:
            if ref is not None:
                k = ref.OrdinaryKriging(cov, idx, obs, E)
                k.solve()
                k.get_uncertainty()
            else:
                ordinary_kriging(cov, idx, obs, E, big_temporaries=True)
    dt = time.perf_counter() - t0
    return n_months / dt, dt


def kriging_1deg_inputs(n_obs):
    from glomargridding_b200 import synthetic as synth

    lats, lons = synth.global_grid(1.0)
    N = lats.size * lons.size
    idx, obs = synth.fixed_observations(4, N, n_obs)
    return lats, lons, N, idx, obs


def cpu_kriging_1deg_sample(n_obs, n_cols):
    """configs[3] on the host cores, bounded.  The reference solves the (n+1) x (n+1)
    Lagrange-extended system for all 64 800 grid columns with LAPACK dgesv (kriging.py:571-586): ~5.3e12 flop of LU
    and ~5.2e13 flop of triangular solves.  Timed here: the SAME two LAPACK stages - dgetrf on the full extended
    matrix, dgetrs on `n_cols` of the grid columns - the second scaled by 64 800 / n_cols, plus the field and a
    DIAGONAL-ONLY variance on those columns (the reference's `get_uncertainty` forms two 64 800 x 64 800 temporaries,
    67 GB, and cannot run at this size).  The matrix is assembled (untimed) from the oracle's pair formula."""
    import scipy.linalg as sla
    from concurrent.futures import ThreadPoolExecutor

    from glomargridding_b200 import synthetic as synth
    from oracle import EllipseOracle

    use_all_host_threads()
    lats, lons, N, idx, obs = kriging_1deg_inputs(n_obs)
    Lx, Ly, th, sd = synth.ellipse_fields(lats, lons)
    o = EllipseOracle(Lx, Ly, th, sd, lats, lons, 1.5, "Met_Office")
    n = idx.size
    t_setup = time.perf_counter()
    S_ext = np.zeros((n + 1, n + 1))
    S_ext[:n, :n] = synth.hadcrut_style_error_cov(idx, lons.size)

    def strip(a0):
        a1 = min(n, a0 + 64)
        ii, jj = np.meshgrid(np.arange(a0, a1), np.arange(n), indexing="ij")
        up = jj > ii
        vals = o.pairs(idx[ii[up]], idx[jj[up]])
        blk = np.zeros((a1 - a0, n))
        blk[up] = vals
        return a0, a1, blk

    with ThreadPoolExecutor(max_workers=min(16, os.cpu_count() or 1)) as ex:
        for a0, a1, blk in ex.map(strip, range(0, n, 64)):
            S_ext[a0:a1, :n] += blk
            S_ext[:n, a0:a1] += blk.T
    S_ext[np.arange(n), np.arange(n)] += o.stdev[idx] ** 2
    S_ext[n, :n] = 1.0
    S_ext[:n, n] = 1.0
    cols = np.linspace(0, N - 1, n_cols).astype(np.int64)
    rhs = np.ones((n + 1, n_cols))
    lo, hi = np.minimum(idx[:, None], cols[None, :]), np.maximum(idx[:, None], cols[None, :])
    same = lo == hi
    vals = o.pairs(np.where(same, 0, lo).ravel(), np.where(same, 1, hi).ravel()).reshape(n, n_cols)
    rhs[:n] = np.where(same, (o.stdev[idx] ** 2)[:, None], vals)
    t_setup = time.perf_counter() - t_setup
    t0 = time.perf_counter()
    lu, piv = sla.lu_factor(S_ext, overwrite_a=True, check_finite=False)
    t_factor = time.perf_counter() - t0
    t0 = time.perf_counter()
    W = sla.lu_solve((lu, piv), rhs, check_finite=False)                  # (n + 1, n_cols): weights^T incl. multiplier
    field = W[:n].T @ obs
    u2 = o.stdev[cols] ** 2 - np.einsum("ak,ak->k", W, rhs) - W[n]       # diag(C - W Cx_ext) - alpha, diagonal only
    t_solve = time.perf_counter() - t0
    total = t_factor + t_solve * (N / n_cols)
    assert np.isfinite(field).all() and np.isfinite(u2).all()
    return {"seconds_scaled": total, "t_factor": t_factor, "t_solve_sample": t_solve, "t_setup": t_setup,
            "n_cols": n_cols}


def cpu_stochastic_sample(n_members, ref=None):
    """configs[4] on the host cores: StochasticKriging weights, 200 states from the
    covariance and 200 observation-noise vectors (numpy's SVD-based multivariate normal, stochastic.py:460), the
    epsilon GEMM and the ordinary-kriging field - unmodified reference classes when `ref` is given."""
    from glomargridding_b200 import synthetic as synth
    from oracle import EllipseOracle, mv_normal_draw, stochastic_kriging

    use_all_host_threads()
    lats, lons = synth.global_grid(5.0)
    Lx, Ly, th, sd = synth.ellipse_fields(lats, lons)
    cov = EllipseOracle(Lx, Ly, th, sd, lats, lons, 1.5, "Met_Office").build()
    M = cov.shape[0]
    idx, obs = synth.fixed_observations(201403, M, 1341)
    E = synth.hadcrut_style_error_cov(idx, lons.size)
    np.random.seed(20140371)
    t0 = time.perf_counter()
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        if ref is not None:
            sk = ref.StochasticKriging(cov, idx, obs, E)
            sk.get_kriging_weights()
            W_ok, W_sk = sk.kriging_weights, sk.simple_kriging_weights
            X = ref.scipy_mv_normal_draw(np.zeros(M), cov, ndraws=n_members)
            noise = ref.scipy_mv_normal_draw(np.zeros(idx.size), E, ndraws=n_members)
        else:
            r = stochastic_kriging(cov, idx, obs, E, simulated_state=np.zeros(M), simulated_obs_noise=np.zeros(idx.size),
                                   big_temporaries=False)
            W_ok, W_sk = r["weights"], r["simple_weights"]
            X = mv_normal_draw(np.zeros(M), cov, n_members)
            noise = mv_normal_draw(np.zeros(idx.size), E, n_members)
        eps = (X[:, idx] + noise) @ W_sk.T - X
        result = (W_ok @ np.append(obs, 0.0))[None, :] + eps
    dt = time.perf_counter() - t0
    assert result.shape == (n_members, M)
    return n_members / dt, dt


def notebook_case():
    """configs[0] inputs (`tests/golden/notebook_2014_03.npz`, decoded from the reference notebook's netCDF files by
    `tests/golden/make_golden_notebook.py`): observation cells and values of the HadSST.4.0.1.0 2014-03 ensemble-member
    field, the full-grid float32 HadCRUT.5 error covariance of that month, and the unmodified reference's outputs."""
    g = np.load(os.path.join(ROOT, "tests", "golden", "notebook_2014_03.npz"))
    E = np.zeros((2592, 2592), dtype=np.float32)
    E[g["E_rows"], g["E_cols"]] = g["E_vals"]
    return g, E


def cpu_notebook_sample(ref=None):
    """configs[0] on the host cores, whole (it is the reference's own CPU-runnable case): 5-degree ellipse covariance
    build -> eigenvalue_clip -> SimpleKriging construct + solve + get_uncertainty (`notebooks/Ellipse_to_grid.ipynb`
    cells 19-34) - the unmodified classes when `ref` is given, the oracle port otherwise."""
    import contextlib
    import io

    from glomargridding_b200 import synthetic as synth

    use_all_host_threads()
    g, E = notebook_case()
    idx, obs = g["idx"], g["obs"]
    lats, lons = synth.global_grid(5.0)
    Lx, Ly, th, sd = synth.ellipse_fields(lats, lons)
    t0 = time.perf_counter()
    with warnings.catch_warnings(), contextlib.redirect_stdout(io.StringIO()):
        warnings.simplefilter("ignore")
        if ref is not None:
            cov = ref.EllipseCovarianceBuilder(Lx, Ly, th, sd, lats, lons, v=1.5, precision=np.float64).cov_ns
            t1 = time.perf_counter()
            cov = ref.covariance_tools.eigenvalue_clip(cov)
            t2 = time.perf_counter()
            k = ref.SimpleKriging(cov, idx.copy(), obs.copy(), E.copy())
            z, u = k.solve(), k.get_uncertainty()
        else:
            from oracle import EllipseOracle, simple_kriging
            from oracle.covariance_tools import explained_variance_clip

            cov = EllipseOracle(Lx, Ly, th, sd, lats, lons, 1.5, "Modified_Met_Office").build()
            t1 = time.perf_counter()
            cov = explained_variance_clip(cov)[0]
            t2 = time.perf_counter()
            r = simple_kriging(cov, idx, obs, E)
            z, u = r["field"], r["uncertainty"]
    t3 = time.perf_counter()
    err = float(np.abs(z - g["sk_field"]).max() / np.abs(g["sk_field"]).max())
    return {"seconds": t3 - t0, "t_build": t1 - t0, "t_clip": t2 - t1, "t_krige": t3 - t2, "field_err": err}


def host_threads():
    try:
        from threadpoolctl import threadpool_info

        return max([p.get("num_threads", 1) for p in threadpool_info()] + [1])
    except Exception:
        return os.cpu_count() or 1


def use_all_host_threads():
    """torchrun exports OMP_NUM_THREADS=1; the CPU arm is meant to use every host core for LAPACK/BLAS."""
    try:
        from threadpoolctl import threadpool_limits

        threadpool_limits(limits=os.cpu_count() or 1)
    except Exception:
        pass


_FIT_KW = dict(max_distance=3000.0, min_distance=0.3, delta_x_method="Modified_Met_Office", tol=1e-4)
_FIT_MODEL = dict(anisotropic=True, rotated=True, physical_distance=True, v=1.5, unit_sigma=False)


class _Coord:
    def __init__(self, values):
        self.values = np.asarray(values)


def fit_cube(res):
    from glomargridding_b200 import synthetic as synth

    data, lats, lons = synth.training_cube(res, 48, 11)
    coords = {"time": _Coord(np.arange(data.shape[0])), "latitude": _Coord(lats), "longitude": _Coord(lons)}
    return data, coords


def cpu_ellipse_fit_sample(res, n_cells, ref=None):
    """The reference's per-cell fit (`EllipseBuilder.fit_ellipse_model`, estimate.py:187-400: training-pair selection +
    scipy Nelder-Mead on the Fisher-transformed likelihood) for `n_cells` cells spread over the unmasked cells of the
    training cube - the unmodified classes when `ref` is given, else the oracle's likelihood + scipy's optimiser on
    the same training pairs.  cells/s."""
    import importlib

    data, coords = fit_cube(res)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        if ref is not None:
            model = importlib.import_module("glomar_gridding.ellipse.model")
            sys.modules["glomar_gridding.ellipse"].EllipseModel = model.EllipseModel
            estimate = importlib.import_module("glomar_gridding.ellipse.estimate")
            import contextlib
            import io

            with contextlib.redirect_stdout(io.StringIO()):
                b = estimate.EllipseBuilder(data, coords)
            m = model.EllipseModel(**_FIT_MODEL)
            cells = np.linspace(0, int(b.small_covar_size) - 1, n_cells).astype(int)
            t0 = time.perf_counter()
            for k in cells:
                b.fit_ellipse_model(int(k), m, estimate_SE=None, **_FIT_KW)
            dt = time.perf_counter() - t0
            return n_cells / dt, dt, int(b.small_covar_size)
        from oracle.ellipse_fit import EllipseModelOracle

        flat = np.ma.getdata(data).reshape(data.shape[0], -1)
        keep = ~np.any(np.ma.getmaskarray(data), axis=0).ravel()
        flat = flat[:, keep] - flat[:, keep].mean(0, keepdims=True)
        cov = flat.T @ flat / (data.shape[0] - 1)
        sd = np.sqrt(np.diag(cov))
        cor = cov / np.outer(sd, sd)
        lon_g, lat_g = np.meshgrid(coords["longitude"].values, coords["latitude"].values)
        lat, lon = lat_g.ravel()[keep], lon_g.ravel()[keep]
        m = EllipseModelOracle(True, True, True, 1.5, unit_sigma=False)
        cells = np.linspace(0, lat.size - 1, n_cells).astype(int)
        t0 = time.perf_counter()
        for k in cells:
            dy, dx = lat - lat[k], lon - lon[k]
            dx = np.where(dx > 180, dx - 360, np.where(dx < -180, dx + 360, dx))
            la, lk = np.radians(lat), np.radians(lat[k])
            hav = 2 * np.arcsin(np.sqrt(np.sin(0.5 * (lk - la)) ** 2 + np.cos(lk) * np.cos(la)
                                        * np.sin(0.5 * np.radians(lon[k] - lon)) ** 2)) * 6371.0
            ok = (hav <= 3000.0) & (hav >= 0.3) & (hav != 0)
            X = 6371.0 * np.column_stack([np.radians(dx) * 0.5 * (np.cos(la) + np.cos(lk)), np.radians(dy)])[ok]
            m.fit(X, cor[k, ok].copy(), tol=1e-4)
        dt = time.perf_counter() - t0
        return n_cells / dt, dt, int(lat.size)


def cpu_sub_baselines(args, ref):
    """cpu_baseline objects of the three secondary workloads (rank 0, N = 1 / the reference arm)."""
    kind = "reference" if ref is not None else "port"
    out = {}
    if not args.no_kriging:
        mps, dt = cpu_kriging_sample(args.cpu_months, ref)
        out["kriging"] = {"value": mps, "unit": "months/s", "cores": host_threads(), "kind": kind,
                          "sample": f"{args.cpu_months} of the {args.months} months ({dt:.1f} s): OrdinaryKriging "
                                    "construct + solve + get_uncertainty, numpy LAPACK/BLAS on all host threads"}
    if not args.no_kriging_1deg:
        r = cpu_kriging_1deg_sample(args.nobs_1deg, args.cpu_cols_1deg)
        out["kriging_1deg"] = {
            "value": 1.0 / r["seconds_scaled"], "unit": "solves/s", "cores": host_threads(), "kind": "port",
            "seconds_scaled": r["seconds_scaled"],
            "sample": f"LAPACK dgetrf of the full ({args.nobs_1deg}+1)-extended system ({r['t_factor']:.1f} s) + dgetrs, "
                      f"field and DIAGONAL-ONLY variance on {r['n_cols']} of the 64800 grid columns "
                      f"({r['t_solve_sample']:.2f} s, scaled x{64800 / r['n_cols']:.1f}); what np.linalg.solve does "
                      "at kriging.py:584-586; the reference's get_uncertainty needs two 64800^2 temporaries (67 GB) "
                      f"and cannot run; matrix assembly by the oracle untimed ({r['t_setup']:.0f} s)"}
    if not args.no_ellipse_fit:
        cps, dt, n_all = cpu_ellipse_fit_sample(args.fit_res, args.cpu_fit_cells, ref)
        out["ellipse_fit"] = {"value": cps, "unit": "cells/s", "cores": 1, "kind": kind,
                              "sample": f"{args.cpu_fit_cells} of the {n_all} unmasked cells of the {args.fit_res:g}deg "
                                        f"training cube ({dt:.1f} s): EllipseBuilder.fit_ellipse_model (training-pair "
                                        "selection + scipy Nelder-Mead, anisotropic_rotated_pd, v = 1.5, sigma "
                                        "fitted), a serial Python loop like the reference's compute_params"}
    if not args.no_notebook:
        try:
            r = cpu_notebook_sample(ref)
            out["notebook"] = {"value": 1.0 / r["seconds"], "unit": "fields/s", "cores": host_threads(), "kind": kind,
                               "seconds": r["seconds"],
                               "sample": f"the whole case once ({r['seconds']:.1f} s): EllipseCovarianceBuilder 5deg build "
                                         f"{r['t_build']:.1f} s (single-threaded numpy), eigenvalue_clip {r['t_clip']:.1f} s, "
                                         f"SimpleKriging construct + solve + get_uncertainty {r['t_krige']:.1f} s (LAPACK / "
                                         f"BLAS on all host threads); field {r['field_err']:.1e} from the committed golden"}
        except Exception as e:  # a missing oracle symbol must not take the other legs down
            out["notebook"] = {"error": repr(e)}
    if not args.no_stochastic:
        mps, dt = cpu_stochastic_sample(args.members, ref)
        out["stochastic"] = {"value": mps, "unit": "members/s", "cores": host_threads(), "kind": kind,
                             "sample": f"all {args.members} members once ({dt:.1f} s): StochasticKriging weights "
                                       "(explicit inverse), SVD-based draws of 200 states and 200 noise vectors "
                                       "(np.random.multivariate_normal), epsilon GEMM, ordinary-kriging field"}
    return out


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    wl = workload(args.res, args.nu)
    ref = reference_ns()
    kind = "reference" if ref is not None else "port"

    def sample(full):
        if ref is not None:
            return ref_cov_sample(ref, wl, args.ref_batches if full else 1, 2_000_000 if full else 100_000)
        return cpu_cov_sample(wl, args.cpu_rows if full else 8)

    for _ in range(min(args.warmup, 1)):
        sample(False)
    vals, times = [], []
    for _ in range(args.steps):
        v, pairs, dt = sample(True)
        vals.append(v)
        times.append(dt)
    value = float(np.mean(vals))
    if ref is not None:
        what = (f"the unmodified reference's batched build (ellipse/covariance.py:382-401: itertools pair generation, "
                f"np.asarray(batch), c_ij_anisotropic_array, scatter) for its first {args.ref_batches} batches of "
                f"2 000 000 pairs per step ({pairs} pairs), precision float64; single-threaded like the reference; "
                f"entries/s = 2*pairs/s")
    else:
        what = (f"all i<j pairs with i in a {args.cpu_rows}-row block of the {wl['N']}-cell matrix per step "
                f"({pairs} pairs), numpy/scipy oracle port (the reference's covariance build is single-threaded); "
                f"entries/s = 2*pairs/s")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": float(np.mean(times) * 1e3),
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": config_dict(args, wl, args.gpus),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": 1, "kind": kind, "sample": what},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    names = {"kriging": "kriged months/s", "kriging_1deg": "1deg ordinary kriging solves/s (n=20000)",
             "stochastic": "stochastic kriging members/s", "ellipse_fit": "ellipse fits (grid cells)/s",
             "notebook": "5deg build + clip + simple kriging of the HadSST 2014-03 field, fields/s"}
    for key, cb in cpu_sub_baselines(args, ref).items():
        line[key] = {"metric": names[key], "value": cb["value"], "unit": cb["unit"], "cpu_baseline": cb,
                     "e2e": {"value": cb["value"], "unit": cb["unit"], "h2d_bytes_per_step": 0,
                             "d2h_bytes_per_step": 0}}
    emit(line)


# ------------------------------------------------------------------------------------------------
# clocks sampling during the timed region
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    def __init__(self, index):
        self.samples, self.reasons, self._stop = [], set(), threading.Event()
        self.max_mhz = None
        try:
            import pynvml

            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None
        self.t = threading.Thread(target=self._run, daemon=True)

    def _run(self):
        nv = self.nv
        names = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "hw_thermal_slowdown": 0x40,
                 "sw_thermal_slowdown": 0x20, "hw_power_brake": 0x80}
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            time.sleep(0.02)

    def __enter__(self):
        if self.nv:
            self.t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        if self.nv:
            self.t.join()

    def summary(self):
        return {"sm_mhz": float(np.median(self.samples)) if self.samples else None,
                "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}


def load_json(path, default):
    try:
        with open(os.path.join(ROOT, path)) as f:
            return json.load(f)
    except Exception:
        return default


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------
def run_b200(args):
    import torch
    import torch.distributed as dist

    from glomargridding_b200._lib import check, lib
    from glomargridding_b200.ellipse import EllipseCovarianceBuilder

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        # NCCL prints its version banner on stdout when the first communicator is created: keep stdout to the one
        # JSON line by pointing fd 1 at stderr until the communicator exists
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=torch.device("cuda", local))
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)
    device = torch.device("cuda", local)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        t = torch.tensor([x], dtype=torch.float64, device=device)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    wl = workload(args.res, args.nu)
    N = wl["N"]
    row0, nrows = row_block(N, rank, world)
    L = lib()
    warmup = max(args.warmup, 3)

    # ---- device-resident inputs (the builder's own host prep, done once) ------------------------
    probe = EllipseCovarianceBuilder(wl["Lx"], wl["Ly"], wl["theta"], wl["stdev"], wl["lats"], wl["lons"],
                                     v=args.nu, precision=np.float64, rows=(row0, min(nrows, TILE)))
    ins = probe._inputs_device
    ws = probe._ws
    ws_bytes = ws.numel()
    out = torch.empty((nrows, N), dtype=torch.float64, device=device)
    stream = torch.cuda.current_stream()

    def cov_kernel():
        check(L.gmg_ellipse_cov_f64(ws.data_ptr(), N, row0, nrows, args.nu, -1.0, 0, out.data_ptr(), N, 0,
                                    stream.cuda_stream), "cov")

    def step():
        check(L.gmg_ellipse_prepare_f64(*[t.data_ptr() for t in ins], N, args.nu, 1, ws.data_ptr(), ws_bytes,
                                        stream.cuda_stream), "prepare")
        cov_kernel()

    for _ in range(warmup):
        step()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    # the dominant kernel's launches are bracketed by their own events INSIDE the timed region (events do not
    # serialise anything: the stream is in order already)
    kev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    with ClockSampler(local) as clocks:
        e0.record(stream)
        for k0, k1 in kev:
            check(L.gmg_ellipse_prepare_f64(*[t.data_ptr() for t in ins], N, args.nu, 1, ws.data_ptr(), ws_bytes,
                                            stream.cuda_stream), "prepare")
            k0.record(stream)
            cov_kernel()
            k1.record(stream)
        e1.record(stream)
        barrier()
        total_ms = e0.elapsed_time(e1)
    kernel_ms = [k0.elapsed_time(k1) for k0, k1 in kev]
    ms_per_step = max_over_ranks(total_ms) / args.steps
    value = float(N) * float(N) / (ms_per_step * 1e-3)
    fp64_peaks = measure_fp64_peaks(L, device, local, total_ms) if rank == 0 and not args.no_fp64_peaks else None

    # ---- end to end through the public API: host inputs -> numpy cov_ns (pinned) ------------------
    e2e = None
    if not args.no_e2e:
        def e2e_step():
            b = EllipseCovarianceBuilder(wl["Lx"], wl["Ly"], wl["theta"], wl["stdev"], wl["lats"], wl["lons"],
                                         v=args.nu, precision=np.float64, rows=(row0, nrows))
            c = b.cov_ns  # D2H of the rank's row block into pinned host memory
            return c[0, 0] + c[-1, -1]

        e2e_step()  # warm-up: pins the host block (torch's caching host allocator reuses it afterwards)
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.e2e_steps):
            e2e_step()
        barrier()
        dt = max_over_ranks((time.perf_counter() - t0) / args.e2e_steps)
        e2e = {"value": float(N) * float(N) / dt, "unit": UNIT, "h2d_bytes_per_step": 6 * N * 8,
               "d2h_bytes_per_step": int(nrows) * N * 8, "steps": args.e2e_steps,
               "note": "EllipseCovarianceBuilder(host arrays).cov_ns -> numpy in pinned host memory, per rank row "
                       "block; the 8*N^2/world-byte device->host copy of the matrix dominates"}
    del out
    torch.cuda.empty_cache()

    # the DMMA rate the secondary workloads divide by: measured by rank 0 in this run, shared with every rank
    dmma_peak = max_over_ranks(fp64_peaks["dmma_tflops"] if fp64_peaks else 0.0) or DMMA_PEAK_TFLOPS
    def leg(fn, *a):
        # on one GPU a failing secondary leg is reported in its own object and the headline line still prints; with
        # several ranks the legs hold collectives, so an exception must take the whole job down instead of hanging it
        if world > 1:
            return fn(*a)
        try:
            return fn(*a)
        except Exception as e:
            torch.cuda.synchronize()
            return {"error": repr(e)}

    common_args = (args, rank, world, device, barrier, max_over_ranks, local)
    kriging = None if args.no_kriging else leg(run_months, *common_args, dmma_peak)
    kriging_1deg = None if args.no_kriging_1deg else leg(run_kriging_1deg, *common_args, dmma_peak)
    stochastic = None if args.no_stochastic else leg(run_stochastic, *common_args, dmma_peak)
    ellipse_fit = None if args.no_ellipse_fit else leg(run_ellipse_fit, *common_args)
    notebook = None
    if not args.no_notebook:
        try:                       # rank 0 only and no collective inside: a failure here cannot hang the other ranks
            notebook = run_notebook(args, rank, world, device, barrier, local)
        except Exception as e:
            notebook = {"error": repr(e)}
        barrier()                  # the other ranks wait here, so that every rank leaves the process group together

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peaks = load_json("MEASURED_PEAKS.json", {})
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    k_ms = float(np.mean(kernel_ms))
    hbm_achieved = 8.0 * nrows * N / (k_ms * 1e-3) / 1e9
    general_nu = args.nu not in (0.5, 1.5, 2.5)
    # pairs this rank's launch evaluates: 64 x 64 tiles, upper triangle only inside the square block on the diagonal
    tr, tc = -(-nrows // 64), -(-N // 64)
    pairs_eval = 4096.0 * (tr * (tr + 1) // 2 + tr * (tc - tr))
    fp64_instr = K1_FP64_INSTR_PER_PAIR["table" if general_nu else "closed"] * pairs_eval
    # the kernel is timed inside a long step (args.steps builds back to back): the SUSTAINED rate is its ceiling
    dfma_peak = fp64_peaks["dfma_tflops_sustained"] if fp64_peaks else DFMA_PEAK_TFLOPS
    fp64_achieved = 2.0 * fp64_instr / (k_ms * 1e-3) / 1e12
    prof = load_json("profiles/r02_ellipse_full_traffic.json", None) or load_json("profiles/r01_ellipse_full_traffic.json", {})
    replay = world == 1 and abs(args.nu - prof.get("nu", -1)) < 1e-12
    common = {"kernel": "ellipse_cov_tile_kernel<%s, f64> (persistent, 2 CTAs/SM)" % ("TABLE" if general_nu else "closed form"),
              "kernel_ms": k_ms, "traffic": prof.get("dram_bytes_per_launch") if replay else None,
              "traffic_replayed_from": prof.get("source") if replay else None,
              "algorithmic_bytes_per_launch": 8 * int(nrows) * N, "pairs_evaluated_per_launch": pairs_eval,
              "hbm": {"achieved": hbm_achieved, "peak": hbm_peak, "unit": "GB/s", "frac": hbm_achieved / hbm_peak,
                      "peak_source": "MEASURED_PEAKS.json hbm_gbs (burst copy)" if peaks else "fallback 6650 GB/s"},
              "fp64": {"achieved": fp64_achieved, "peak": dfma_peak, "unit": "TFLOP/s", "frac": fp64_achieved / dfma_peak,
                       "fp64_instructions_per_pair": fp64_instr / pairs_eval,
                       "fp64_instructions_per_pair_executed": K1_FP64_EXECUTED["table" if general_nu else "closed"],
                       "peak_source": "gmg_microbench_fp64 DFMA, sustained over a window as long as the timed region, in this "
                                      "run (burst: fp64_peaks.dfma_tflops)" if fp64_peaks else "fallback constant",
                       "frac_of_burst_peak": fp64_achieved / fp64_peaks["dfma_tflops"] if fp64_peaks else None,
                       "ncu_fp64_pipe_active_replayed": prof.get("fp64_pipe_pct_of_peak") if replay else None}}
    if general_nu:
        roofline = dict(bound="fp64", achieved=fp64_achieved, peak=dfma_peak, unit="TFLOP/s", frac=fp64_achieved / dfma_peak,
                        note="general nu (SURVEY 8d): FP64-pipe bound.  achieved = FP64 pipe instructions per launch "
                             "(frozen SASS count x evaluated pairs), each counted as one FMA slot = 2 flop, / kernel "
                             "time (events around every launch of the timed region); peak = sustained DFMA rate measured in "
                             "this run.  The HBM-write view of the same launch (8 B "
                             "per entry against the measured copy bandwidth) is under 'hbm'.", **common)
    else:
        roofline = dict(bound="hbm", achieved=hbm_achieved, peak=hbm_peak, unit="GB/s", frac=hbm_achieved / hbm_peak,
                        note="closed-form nu (SURVEY 8d): HBM-write bound, 8 B per entry", **common)
    cpu = None
    if not args.no_cpu_baseline and world == 1:      # reported at N = 1 only (the reference arm covers N > 1)
        ref = reference_ns()
        if ref is not None:
            v, pairs, dt = ref_cov_sample(ref, wl, args.ref_batches)
            cpu = {"value": v, "unit": UNIT, "cores": 1, "kind": "reference",
                   "sample": f"the unmodified reference's batched build loop (ellipse/covariance.py:382-401) for its "
                             f"first {args.ref_batches} batches of 2 000 000 pairs ({pairs} pairs, {dt:.1f} s), "
                             "single-threaded like the reference; entries/s = 2*pairs/s"}
        else:
            v, pairs, dt = cpu_cov_sample(wl, args.cpu_rows)
            cpu = {"value": v, "unit": UNIT, "cores": 1, "kind": "port",
                   "sample": f"all i<j pairs with i in a {args.cpu_rows}-row block ({pairs} pairs, {dt:.1f} s), "
                             f"numpy/scipy oracle port (the reference's build is single-threaded); entries/s = 2*pairs/s"}
        subs = cpu_sub_baselines(args, ref)
        for key, obj in (("kriging", kriging), ("kriging_1deg", kriging_1deg), ("stochastic", stochastic),
                         ("ellipse_fit", ellipse_fit), ("notebook", notebook)):
            if obj is not None and key in subs:
                obj["cpu_baseline"] = subs[key]
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config_dict(args, wl, world),
        "clocks": clocks.summary(), "e2e": e2e, "gpu_launches": 3 * args.steps, "roofline": roofline,
        "cpu_baseline": cpu, "fp64_peaks": fp64_peaks, "kriging": kriging, "kriging_1deg": kriging_1deg,
        "stochastic": stochastic, "ellipse_fit": ellipse_fit, "notebook": notebook,
    }
    emit(line)
    if world > 1:
        dist.destroy_process_group()


def measure_fp64_peaks(L, device, local, window_ms=25.0):
    """DFMA-pipe and DMMA (mma.sync m8n8k4 f64) rates of THIS GPU in THIS run (`gmg_microbench_fp64`), with the
    clocks sampled while they run: the denominators of the FP64 roofline fractions (MEASURED_PEAKS.json has none).
    Burst = best single launch; sustained = launches back to back for as long as the bench's own timed region lasted
    (`window_ms`) - the rate the pipe holds under the power cap a long FP64 step runs into."""
    import ctypes

    import torch

    from glomargridding_b200._lib import check

    scratch = torch.zeros(16, dtype=torch.float64, device=device)
    out = {}
    with ClockSampler(local) as clocks:
        for name, what in (("dfma", 0), ("dmma", 1)):
            best, sus = ctypes.c_double(), ctypes.c_double()
            reps = int(min(2000, max(40, window_ms / 0.6)))      # one launch of 4096 iterations is ~0.6 ms
            check(L.gmg_microbench_fp64(what, 1 << 12, reps, scratch.data_ptr(), ctypes.byref(best), ctypes.byref(sus),
                                        torch.cuda.current_stream().cuda_stream), "gmg_microbench_fp64")
            out[f"{name}_tflops"] = best.value
            out[f"{name}_tflops_sustained"] = sus.value
    out["clocks"] = clocks.summary()
    out["how"] = ("gmg_microbench_fp64: register-only kernels, 8 CTAs x 256 threads per SM, 8 independent chains per "
                  "thread; best single launch (burst) and launches back to back over a window as long as the timed "
                  "region of the covariance steps (sustained, %d ms); DFMA = 2 flop per " % round(window_ms) +
                  "lane-instruction, DMMA m8n8k4 = 512 flop per warp-instruction")
    return out


def run_months(args, rank, world, device, barrier, max_over_ranks, local, dmma_peak=DMMA_PEAK_TFLOPS):
    """kriged months/s: `--months` synthetic 5-degree months, ordinary kriging, months round-robin over ranks."""
    import torch

    from glomargridding_b200 import _dense as dn
    from glomargridding_b200 import synthetic as synth
    from glomargridding_b200.batched import krige_batch, shard
    from glomargridding_b200.ellipse import EllipseCovarianceBuilder

    lats, lons = synth.global_grid(5.0)
    Lx, Ly, th, sd = synth.ellipse_fields(lats, lons)
    cov = EllipseCovarianceBuilder(Lx, Ly, th, sd, lats, lons, v=1.5, delta_x_method="Met_Office",
                                   precision=np.float64).cov_ns_device
    M = cov.shape[0]
    mine = shard(args.months, rank, world)
    probs = month_problems(mine, M, lons.size)
    kb = dn.KrigeBatch(cov, probs, ordinary=True)      # inputs resident in HBM before the timed region
    for _ in range(2):
        kb.run()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local) as clocks:
        e0.record()
        for _ in range(args.month_steps):
            kb.run()
        e1.record()
        barrier()
    ms = max_over_ranks(e0.elapsed_time(e1)) / args.month_steps
    ns = np.array(kb.ns, dtype=np.float64)
    flops = float((ns ** 3 / 3 + ns ** 2 * (M + 2)).sum())     # Cholesky + the forward triangular solve
    tflops_local = flops / (ms * 1e-3) / 1e12
    nblk = int(np.ceil(kb.n_max / 128))
    launches = 1 + (3 * nblk - 2) + 2 * nblk + 2
    del kb
    torch.cuda.empty_cache()
    e2e = None
    if not args.no_e2e:
        def e2e_step():
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                r = krige_batch(cov, probs, "ordinary")
            return r["field"][0, 0]

        for _ in range(2):                         # warm-up: pinned staging buffers, thread pools, allocator cache
            e2e_step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.e2e_steps):
            e2e_step()
        barrier()
        dt = max_over_ranks(time.perf_counter() - t0) / args.e2e_steps
        # the same months with the error covariances as float32 - what HadCRUT.5 ships (tests/conftest.py) and what the
        # reference's notebook feeds: they travel as float32 and are widened by the gather kernel (same bits out)
        probs64 = probs
        probs = [(i, o, E.astype(np.float32)) for i, o, E in probs64]
        for _ in range(2):
            e2e_step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.e2e_steps):
            e2e_step()
        barrier()
        dt32 = max_over_ranks(time.perf_counter() - t0) / args.e2e_steps
        probs = probs64
        e2e = {"value": args.months / dt, "unit": "months/s",
               "float32_error_cov": {"value": args.months / dt32, "unit": "months/s",
                                     "note": "the same call with float32 error covariances (HadCRUT.5's dtype): "
                                             "uploaded as float32, half the host-to-device bytes, identical results"},
               "h2d_bytes_per_step": int(sum(8 * (dn.packed_error_cov_doubles(len(p[0])) + 2 * len(p[0])) for p in probs)),
               "d2h_bytes_per_step": int(len(probs) * 5 * M * 8), "steps": args.e2e_steps,
               "note": "krige_batch(cov on device, host idx/obs/error_cov per month) -> numpy field, uncertainty, "
                       "constraint mask; per-rank share of the months; symmetric error covariances are uploaded "
                       "as their lower block trapezoid (about half of n^2)"}
    return {"metric": "kriged months/s", "value": args.months / (ms * 1e-3), "unit": "months/s",
            "ms_per_step": ms, "steps": args.month_steps, "months": args.months,
            "config": {"workload": f"5deg grid (M={M}) ordinary kriging, {args.months} synthetic months, n~U(600,1600) "
                                   "obs, HadCRUT-style error covariance, Met_Office ellipse covariance (SPD), "
                                   "construct + solve + uncertainty + constraint mask per month"},
            "clocks": clocks.summary(), "gpu_launches": launches * args.month_steps,
            "roofline": {"bound": "tensor", "kernel": "gemm_tile (DMMA m8n8k4 f64) inside trsm_step / potrf_update",
                         "achieved": tflops_local, "peak": dmma_peak, "unit": "TFLOP/s",
                         "frac": tflops_local / dmma_peak, "traffic": None,
                         "peak_source": "mma.sync m8n8k4 f64 rate measured in this run (fp64_peaks.dmma_tflops; "
                                        "MEASURED_PEAKS.json has no FP64 figure)",
                         "note": "flops actually executed, n^3/3 + n^2 (M+2) per month, over the whole "
                                 "gather+factor+solve+apply step of this rank.  SURVEY 8d counts 2 n^2 M for two "
                                 "triangular solves; field / variance / mask only need the forward one "
                                 "(d = ||L^-1 c||^2), so half of that work is not done at all"},
            "e2e": e2e}


def run_kriging_1deg(args, rank, world, device, barrier, max_over_ranks, local, dmma_peak):
    """configs[3]: the 1-degree covariance is built in row blocks (one per rank, never gathered) and
    `--nobs-1deg` observations are kriged against it: `distributed.krige_sharded` = assemble S (one exchange),
    distributed Cholesky, column-sharded forward solve, apply, gather of the M-vectors.  One "solve" = one complete
    ordinary kriging (field + uncertainty + constraint mask on all 64 800 cells)."""
    import torch

    from glomargridding_b200 import synthetic as synth
    from glomargridding_b200.distributed import krige_sharded
    from glomargridding_b200.ellipse import EllipseCovarianceBuilder

    lats, lons, N, idx, obs = kriging_1deg_inputs(args.nobs_1deg)
    Lx, Ly, th, sd = synth.ellipse_fields(lats, lons)
    r0, nrows = row_block(N, rank, world)
    n = idx.size

    def build():
        return EllipseCovarianceBuilder(Lx, Ly, th, sd, lats, lons, v=1.5, delta_x_method="Met_Office",
                                        precision=np.float64, rows=(r0, nrows))

    b = build()
    del b
    barrier()
    t0 = time.perf_counter()
    b = build()
    barrier()
    build_s = max_over_ranks(time.perf_counter() - t0)
    cov_rows = b.cov_ns_device
    E = synth.hadcrut_style_error_cov(idx, lons.size)            # host, n x n, identical on every rank
    E_dev = torch.from_numpy(E).to(device)

    def solve(err):
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            return krige_sharded(cov_rows, r0, idx, obs, err, "ordinary")

    res = solve(E_dev)                                            # warm-up: communicators, workspaces, pinned buffers
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local) as clocks:
        e0.record()
        for _ in range(args.krige_1deg_steps):
            res = solve(E_dev)
        e1.record()
        barrier()
    sec = max_over_ranks(e0.elapsed_time(e1)) * 1e-3 / args.krige_1deg_steps
    flops = n ** 3 / 3.0 + float(n) ** 2 * (N + 2)
    tflops_per_gpu = flops / sec / world / 1e12
    e2e = None
    if not args.no_e2e:
        solve(E)
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.e2e_steps):
            res_h = solve(E)
        barrier()
        dt = max_over_ranks(time.perf_counter() - t0) / args.e2e_steps
        mine = int(((idx >= r0) & (idx < r0 + nrows)).sum())
        e2e = {"value": 1.0 / dt, "unit": "solves/s", "seconds": dt, "h2d_bytes_per_step": 8 * (mine * n + 2 * n),
               "d2h_bytes_per_step": 5 * N * 8, "steps": args.e2e_steps,
               "note": "krige_sharded(row block on device, host idx / obs / n x n error covariance) -> numpy field, "
                       "uncertainty, constraint mask on every rank; each rank uploads only the rows of the error "
                       "covariance that belong to its own observations (bytes are per rank)"}
        assert np.array_equal(res_h["field"], res["field"])
    golden = None
    try:                                                         # parity evidence in the driver's own run
        g = np.load(os.path.join(ROOT, "tests", "golden", "krige_1deg_n20000.npz"))
        if int(g["n_obs"]) == n:
            rel = lambda a, w: float(np.max(np.abs(a - w)) / np.max(np.abs(w)))  # noqa: E731
            c = g["cells"]
            golden = {"field": rel(res["field"][c], g["field"]),
                      "variance": rel(res["uncertainty"][c] ** 2, g["uncertainty"] ** 2),
                      "constraint_mask": rel(res["constraint_mask"][c], g["constraint_mask"]),
                      "what": "max-norm relative difference to the unmodified reference's OrdinaryKriging on the "
                              "reduced grid (tests/golden/make_golden_1deg_krige.py), 20 064 cells"}
    except Exception as exc:  # pragma: no cover
        golden = {"error": repr(exc)}
    del cov_rows, b, E_dev
    torch.cuda.empty_cache()
    nblk = -(-n // 128)
    return {"metric": "1deg ordinary kriging solves/s (n=%d)" % n, "value": 1.0 / sec, "unit": "solves/s",
            "seconds": sec, "higher_is_better": True, "steps": args.krige_1deg_steps,
            "covariance_build_seconds": build_s,
            "config": {"workload": f"1deg grid (M={N}) ordinary kriging of {n} observations (default_rng(4)), "
                                   "HadCRUT-style n x n error covariance, Met_Office ellipse covariance nu=1.5 built in "
                                   f"row blocks x{world} and never gathered; field + uncertainty + constraint mask"},
            "clocks": clocks.summary(), "gpu_launches": (1 + 3 * nblk + 2 * nblk + 2) * args.krige_1deg_steps,
            "roofline": {"bound": "tensor", "kernel": "gemm_tile (DMMA m8n8k4 f64) inside potrf_update / trsm_step_fwd",
                         "achieved": tflops_per_gpu, "peak": dmma_peak, "unit": "TFLOP/s",
                         "frac": tflops_per_gpu / dmma_peak, "traffic": None,
                         "peak_source": "fp64_peaks.dmma_tflops of this run",
                         "note": "n^3/3 + n^2 (M+2) flop of the whole job / seconds / n_gpus, against one GPU's DMMA "
                                 "rate: the fraction covers EVERYTHING between host inputs on the device and numpy "
                                 "results (assembly, exchanges, factorisation, solve, apply, gather)"},
            "e2e": e2e, "golden_parity": golden}


def run_stochastic(args, rank, world, device, barrier, max_over_ranks, local, dmma_peak):
    """configs[4]: `--members` perturbed ordinary-kriging fields on the 5-degree grid (n = 1 341 observations,
    HadCRUT-style error covariance), members sharded over the ranks (`batched.stochastic_ensemble`)."""
    import torch

    from glomargridding_b200 import synthetic as synth
    from glomargridding_b200.batched import shard, stochastic_ensemble
    from glomargridding_b200.ellipse import EllipseCovarianceBuilder

    lats, lons = synth.global_grid(5.0)
    Lx, Ly, th, sd = synth.ellipse_fields(lats, lons)
    cov = EllipseCovarianceBuilder(Lx, Ly, th, sd, lats, lons, v=1.5, delta_x_method="Met_Office",
                                   precision=np.float64).cov_ns_device
    M = cov.shape[0]
    idx, obs = synth.fixed_observations(201403, M, 1341)
    E = synth.hadcrut_style_error_cov(idx, lons.size)
    mine = len(shard(args.members, rank, world))
    n = idx.size

    def step(dev_out):
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            return stochastic_ensemble(cov, idx, obs, E, mine, seed=20140371 + rank, return_device=dev_out)

    # (the clock sampler starts BEFORE the warm-up calls: this leg is bound by host launch latency - ~12 ms of small
    # kernels per call - and the sampler's start-up, a subprocess spawn, otherwise lands inside the timed calls)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local) as clocks:
        for _ in range(3):
            step(True)
        barrier()
        e0.record()
        for _ in range(args.stochastic_steps):
            r = step(True)
        e1.record()
        barrier()
    sec = max_over_ranks(e0.elapsed_time(e1)) * 1e-3 / args.stochastic_steps
    assert r["result"].shape == (mine, M)
    flops = (M ** 3 / 3.0 + 2 * n ** 3 / 3.0 + float(n) ** 2 * (M + 2) + 2.0 * mine * M * M + 2.0 * mine * n * n
             + 2.0 * n * n * mine + 2.0 * mine * n * M)
    e2e = None
    if not args.no_e2e:
        step(False)
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.e2e_steps):
            rh = step(False)
        barrier()
        dt = max_over_ranks(time.perf_counter() - t0) / args.e2e_steps
        e2e = {"value": args.members / dt, "unit": "members/s", "seconds": dt,
               "h2d_bytes_per_step": 8 * (n * n + 2 * n), "d2h_bytes_per_step": 8 * (2 * mine * M + 2 * M),
               "steps": args.e2e_steps,
               "note": "stochastic_ensemble(cov on device, host idx / obs / error_cov) -> numpy result, epsilon, "
                       "gridded field, uncertainty; per-rank share of the members"}
        assert np.isfinite(rh["result"]).all()
    return {"metric": "stochastic kriging members/s", "value": args.members / sec, "unit": "members/s",
            "seconds": sec, "higher_is_better": True, "steps": args.stochastic_steps, "members": args.members,
            "config": {"workload": f"5deg grid (M={M}) stochastic kriging: {args.members} ensemble members, n={n} "
                                   "observations, HadCRUT-style error covariance; per call three Cholesky factorisations "
                                   "(C, E, C[idx,idx]+E), Philox normals, draws and epsilon as GEMMs; members sharded "
                                   f"x{world} (every rank repeats the factorisations)"},
            "clocks": clocks.summary(),
            "roofline": {"bound": "tensor", "kernel": "gemm_tile (DMMA m8n8k4 f64)", "achieved": flops / sec / 1e12,
                         "peak": dmma_peak, "unit": "TFLOP/s", "frac": flops / sec / 1e12 / dmma_peak, "traffic": None,
                         "peak_source": "fp64_peaks.dmma_tflops of this run",
                         "note": "flop of one rank's call / its seconds; the error-covariance upload (14 MB) is "
                                 "inside the timed region (the class API takes host arrays), results stay on the device"},
            "e2e": e2e}


def run_notebook(args, rank, world, device, barrier, local):
    """configs[0], the reference's own CPU-runnable case (`notebooks/Ellipse_to_grid.ipynb` cells 19-34): the
    5-degree ellipse covariance build -> `eigenvalue_clip` -> `SimpleKriging` construct + solve + get_uncertainty of the
    real HadSST.4.0.1.0 2014-03 ensemble-member field (1 341 observations) with the real HadCRUT.5 float32 error
    covariance, through the reference-facing classes from HOST arrays to numpy results.  One case, not a batch: every
    rank would repeat it, so rank 0 alone runs it ("replicas only").  `value` = the whole chain with the covariance
    kept on the device between the classes (`cov_ns_device` -> clip -> kriging class), E uploaded and results read
    back every step; `e2e` = the same with the clipped covariance brought to the host as numpy between the steps, the
    way notebook code written against the reference holds it."""
    import torch

    import glomargridding_b200.covariance_tools as ct
    from glomargridding_b200 import synthetic as synth
    from glomargridding_b200.ellipse import EllipseCovarianceBuilder
    from glomargridding_b200.kriging import SimpleKriging

    if rank != 0:
        return None
    g, E = notebook_case()
    idx, obs = g["idx"], g["obs"]
    lats, lons = synth.global_grid(5.0)
    Lx, Ly, th, sd = synth.ellipse_fields(lats, lons)
    M = lats.size * lons.size
    n = idx.size

    def chain(host_cov):
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            b = EllipseCovarianceBuilder(Lx, Ly, th, sd, lats, lons, v=1.5, precision=np.float64)
            cov = ct.eigenvalue_clip(b.cov_ns if host_cov else b.cov_ns_device)
            k = SimpleKriging(cov, idx.copy(), obs.copy(), E)
            return k.solve(), k.get_uncertainty()

    def timed(host_cov, steps):
        for _ in range(2):
            chain(host_cov)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(steps):
            z, u = chain(host_cov)
        torch.cuda.synchronize()
        return (time.perf_counter() - t0) / steps, z, u

    with ClockSampler(local) as clocks:
        sec, z, u = timed(False, args.notebook_steps)
    rel = lambda a, ref: float(np.abs(np.asarray(a) - ref).max() / np.abs(ref).max())  # noqa: E731
    parity = {"field": rel(z, g["sk_field"]), "variance": rel(np.asarray(u) ** 2, g["sk_unc"] ** 2),
              "golden": "tests/golden/notebook_2014_03.npz (the UNMODIFIED reference's SimpleKriging outputs)",
              "tolerance": 1e-8}
    e2e = None
    if not args.no_e2e:
        dt, z2, _ = timed(True, args.notebook_steps)
        e2e = {"value": 1.0 / dt, "unit": "fields/s", "seconds": dt, "steps": args.notebook_steps,
               "h2d_bytes_per_step": 6 * M * 8 + M * M * 8 + n * n * 4 + 2 * n * 8,
               "d2h_bytes_per_step": 2 * M * M * 8 + 2 * M * 8,
               "note": "host Lx / Ly / theta / stdev -> EllipseCovarianceBuilder.cov_ns (numpy) -> eigenvalue_clip (numpy in, "
                       "numpy out) -> SimpleKriging(numpy cov, idx, obs, float32 E).solve() / get_uncertainty() -> numpy"}
        parity["field_e2e"] = rel(z2, g["sk_field"])
    return {"metric": "5deg build + clip + simple kriging of the HadSST 2014-03 field, fields/s", "value": 1.0 / sec,
            "unit": "fields/s", "seconds": sec, "higher_is_better": True, "steps": args.notebook_steps,
            "config": {"workload": f"BASELINE configs[0]: 5deg global grid (M={M}) ellipse covariance (nu=1.5, default "
                                   f"displacement method) -> eigenvalue_clip -> SimpleKriging of the HadSST.4.0.1.0 2014-03 "
                                   f"ensemble-member field (n={n}) with the HadCRUT.5 error covariance; rank 0 only"},
            "golden_parity": parity, "clocks": clocks.summary(),
            "roofline": None,
            "note": "a chain of five small calls (build 0.1 ms of kernel, tridiagonalisation + Sturm + projector of the "
                    "clip, gather + Cholesky + forward solve of one n = 1 341 system): bound by host launch latency and "
                    "the clip's spectrum kernels, no single-kernel roofline applies",
            "e2e": e2e}


def run_ellipse_fit(args, rank, world, device, barrier, max_over_ranks, local):
    """the ellipse parameters of every unmasked cell of a training cube (`EllipseBuilder.compute_params`:
    covariance over time, training-pair selection, batched Nelder-Mead) - cells sharded over the ranks."""
    import torch

    from glomargridding_b200.ellipse import EllipseBuilder, EllipseModel
    from glomargridding_b200.ellipse.model import _run_fit

    data, coords = fit_cube(args.fit_res)
    b = EllipseBuilder(data, coords)
    m = EllipseModel(**_FIT_MODEL)
    n_all = int(b.small_covar_size)
    cells = np.arange(rank, n_all, world, dtype=np.int64)
    g, bnd = m._guesses_bounds(None, None)

    def step(e2e):
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            if e2e:                                  # host cube in -> parameter fields out (one rank's share of cells)
                bb = EllipseBuilder(data, coords)
                offs, x0, x1, y, total, _ = bb._train_device(cells, True, True, _FIT_KW["delta_x_method"], True,
                                                             _FIT_KW["min_distance"], _FIT_KW["max_distance"])
                return _run_fit(m, offs, x0, x1, y, total, cells.size, g, bnd, _FIT_KW["tol"])
            offs, x0, x1, y, total, _ = b._train_device(cells, True, True, _FIT_KW["delta_x_method"], True,
                                                        _FIT_KW["min_distance"], _FIT_KW["max_distance"])
            return _run_fit(m, offs, x0, x1, y, total, cells.size, g, bnd, _FIT_KW["tol"])

    for _ in range(2):
        res = step(False)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local) as clocks:
        e0.record()
        for _ in range(args.fit_steps):
            res = step(False)
        e1.record()
        barrier()
    sec = max_over_ranks(e0.elapsed_time(e1)) * 1e-3 / args.fit_steps
    nfev = int(sum(r.nfev for r in res))
    e2e = None
    if not args.no_e2e:
        step(True)
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.e2e_steps):
            step(True)
        barrier()
        dt = max_over_ranks(time.perf_counter() - t0) / args.e2e_steps
        e2e = {"value": n_all / dt, "unit": "cells/s", "seconds": dt,
               "h2d_bytes_per_step": int(8 * data.shape[0] * n_all), "d2h_bytes_per_step": int(cells.size * 5 * 8 + 8 * n_all * n_all * 2),
               "steps": args.e2e_steps,
               "note": "EllipseBuilder(host training cube) - covariance and correlation over time (one GEMM), mirrored "
                       "to the host like the reference's attributes - then training pairs + fit of this rank's cells"}
    return {"metric": "ellipse fits (grid cells)/s", "value": n_all / sec, "unit": "cells/s", "seconds": sec,
            "higher_is_better": True, "steps": args.fit_steps, "cells": n_all,
            "config": {"workload": f"{args.fit_res:g}deg training cube (48 time steps, {n_all} unmasked cells): "
                                   "anisotropic_rotated_pd ellipse (Lx, Ly, theta, sigma), v = 1.5, neighbours within "
                                   f"3000 km, Nelder-Mead tol 1e-4; cells sharded x{world}"},
            "clocks": clocks.summary(), "gpu_launches": 3 * args.fit_steps,
            "likelihood_evaluations_per_step": nfev,
            "roofline": None, "e2e": e2e}


def emit(line):
    """The one JSON line, written to the process's ORIGINAL stdout (see `main`)."""
    data = (json.dumps(line) + "\n").encode()
    while data:                                   # a pipe may take a 12 KB line in pieces
        data = data[os.write(_JSON_FD, data):]


_JSON_FD = 1


def main():
    # stdout carries the JSON line and nothing else: file descriptor 1 points at stderr for the whole run (library
    # banners, the reference's own print() calls - e.g. cov_2_cor's "Numerical error correction applied" - and
    # anything a child writes land there), and `emit` writes the line to a duplicate of the original descriptor
    global _JSON_FD
    a = parse_args()
    sys.stdout.flush()
    _JSON_FD = os.dup(1)
    os.dup2(2, 1)
    if a.impl == "reference":
        run_reference(a)
    else:
        run_b200(a)


if __name__ == "__main__":
    main()
