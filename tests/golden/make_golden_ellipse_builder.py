"""Golden vectors for `EllipseBuilder` (training-set selection + per-cell ellipse fit),
produced by the UNMODIFIED reference (`glomar_gridding/ellipse/estimate.py`, `model.py`) through the import shim.

    python tests/golden/make_golden_ellipse_builder.py        # build container only (needs /root/reference), ~3 min

Training cube: 48 time steps on a 10-degree global grid (18 x 36) with a land mask, drawn (seeded, numpy only) from a
smooth non-stationary covariance; cells are masked for all times (land) and, for a few cells, at single times (the
"time-varying mask" the reference discards, estimate.py:126-132).  Stored: the cube, its mask, the reference's
covariance / correlation, `_get_train_data` of a few cells for every selection mode, and `fit_ellipse_model(k)` of
EVERY unmasked cell (what `compute_params` loops over; the method itself needs xarray, which is absent) for two
models: anisotropic_rotated_pd (v = 1.5, sigma fitted) and isotropic (degrees, v = 0.5, unit sigma).
"""

from __future__ import annotations

import importlib
import os
import sys
import time
import types
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle.reference_shim import load_reference  # noqa: E402

load_reference()
model = importlib.import_module("glomar_gridding.ellipse.model")
sys.modules["glomar_gridding.ellipse"].EllipseModel = model.EllipseModel        # estimate.py:31 imports it from the package
estimate = importlib.import_module("glomar_gridding.ellipse.estimate")


class _Coord:
    def __init__(self, values):
        self.values = np.asarray(values)


def cube(seed=11, nt=48):
    from glomargridding_b200 import synthetic as synth

    return synth.training_cube(10.0, nt, seed)          # numpy only, seeded: the cube is also stored in the fixture


def main():
    t0 = time.time()
    data, lats, lons = cube()
    coords = {"time": _Coord(np.arange(data.shape[0])), "latitude": _Coord(lats), "longitude": _Coord(lons)}
    b = estimate.EllipseBuilder(data, coords)
    n = int(b.small_covar_size)
    out = {"data": np.ma.getdata(data), "mask": np.ma.getmaskarray(data), "lats": lats, "lons": lons,
           "cov_triu": b.cov[np.triu_indices(n)], "n_cells": np.int64(n)}        # cor: the rows of four cells, below
    # _get_train_data: (anisotropic, physical_distance, delta_x_method, physical_distance_selection, min, max)
    modes = [(True, True, "Modified_Met_Office", True, 0.3, 3000.0), (True, True, "Met_Office", True, 0.3, 3000.0),
             (True, True, "Modified_Met_Office", False, 0.3, 35.0), (False, True, "Modified_Met_Office", True, 0.3, 3000.0),
             (False, True, "Met_Office", False, 0.3, 35.0), (False, False, None, True, 0.3, 35.0),
             (True, False, None, True, 0.3, 35.0), (True, False, "Met_Office", True, 0.3, 3000.0)]
    cells = [0, 7, n // 2, n - 1]
    for mi, (an, pd, meth, sel, dmin, dmax) in enumerate(modes):
        for c in cells:
            X, y = b._get_train_data(c, dmin, dmax, an, meth, physical_distance=pd, physical_distance_selection=sel)
            out[f"train_m{mi}_c{c}_X"], out[f"train_m{mi}_c{c}_y"] = np.asarray(X), np.asarray(y)
    out["train_cells"] = np.array(cells)
    out["cor_rows"] = b.cor[cells]
    # fits of every cell
    for tag, kw, fitkw in (("ard_pd", dict(anisotropic=True, rotated=True, physical_distance=True, v=1.5, unit_sigma=False),
                            dict(max_distance=3000.0, min_distance=0.3, delta_x_method="Modified_Met_Office", tol=1e-4)),
                           ("iso_deg", dict(anisotropic=False, rotated=False, physical_distance=False, v=0.5, unit_sigma=True),
                            dict(max_distance=35.0, min_distance=0.3, delta_x_method=None, tol=1e-4))):
        m = model.EllipseModel(**kw)
        rows, nfev = [], []
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            for k in range(n):
                r = b.fit_ellipse_model(k, m, estimate_SE=None, **fitkw)
                rows.append(r["ModelParams"])
                nfev.append([r["Results"].nfev, r["Results"].status, r["Results"].fun])
        out[f"fit_{tag}"] = np.array(rows, dtype=np.float64)
        out[f"fit_{tag}_aux"] = np.array(nfev, dtype=np.float64)
        print(tag, n, "cells", f"{time.time() - t0:.0f} s", "qc codes", np.unique(out[f'fit_{tag}'][:, -2], return_counts=True))
    np.savez_compressed(os.path.join(HERE, "ellipse_builder.npz"), **out)


if __name__ == "__main__":
    main()
