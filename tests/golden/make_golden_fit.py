"""Golden vectors for the ellipse-parameter fit, produced by the UNMODIFIED reference
(`glomar_gridding/ellipse/model.py`, imported through `oracle/reference_shim.py`).

    python tests/golden/make_golden_fit.py        # build container only (needs /root/reference)

Every case is one grid cell's fit: displacements X (km) of its neighbours on a 1-degree patch, "observed" correlations
y drawn from a known ellipse plus Fisher-space noise (seeded), then the reference's own
  * `cov_ij_anisotropic` / `cov_ij_isotropic` on X for the true parameters,
  * `negative_log_likelihood` at the true parameters and at the default guesses,
  * `fit(..., estimate_SE=None)` with scipy's Nelder-Mead: x, fun, nit.
"""

from __future__ import annotations

import importlib
import os
import sys
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle.reference_shim import load_reference  # noqa: E402

load_reference()
model = importlib.import_module("glomar_gridding.ellipse.model")

# (anisotropic, rotated, physical_distance, v, unit_sigma, true parameters)
CASES = [
    (True, True, True, 0.5, True, (900.0, 450.0, 0.4)),
    (True, True, True, 1.5, True, (1400.0, 600.0, -0.7)),
    (True, True, True, 2.2, False, (700.0, 500.0, 1.1)),
    (True, False, True, 1.5, True, (1000.0, 400.0)),
    (False, False, True, 0.5, True, (800.0,)),
    (True, True, False, 1.5, True, (9.0, 4.0, 0.3)),
]


def patch(physical: bool, rng):
    """Displacements of the 120 neighbours nearest to the centre of a 15 x 15 one-degree patch at 40 N."""
    jj, ii = np.meshgrid(np.arange(-7, 8), np.arange(-7, 8))
    dy, dx = ii.ravel().astype(float), jj.ravel().astype(float) * np.cos(np.deg2rad(40.0))
    keep = (dy != 0) | (dx != 0)
    dx, dy = dx[keep], dy[keep]
    order = np.argsort(np.hypot(dx, dy), kind="stable")[:120]
    dx, dy = dx[order], dy[order]
    if physical:
        dx, dy = dx * 111.12, dy * 111.12
    return dx, dy


def main():
    out = {"n_cases": np.int64(len(CASES))}
    for c, (aniso, rot, pd, v, unit, true) in enumerate(CASES):
        rng = np.random.default_rng(1000 + c)
        m = model.EllipseModel(anisotropic=aniso, rotated=rot, physical_distance=pd, v=v, unit_sigma=unit)
        dx, dy = patch(pd, rng)
        X = np.column_stack([dx, dy]) if aniso else np.hypot(dx, dy)
        kw = dict(zip(("Lx", "Ly", "theta"), true)) if aniso else {"R": true[0]}
        clean = m.cov_ij(X, **kw)
        y = np.tanh(np.arctanh(np.clip(clean, -0.999, 0.999)) + 0.05 * rng.standard_normal(clean.size))
        p_true = list(true) + ([] if unit else [0.05])
        p_guess = list(m.default_guesses) + ([] if unit else [0.1])
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            nll_true = m.negative_log_likelihood(X, y.copy(), p_true)
            nll_guess = m.negative_log_likelihood(X, y.copy(), p_guess)
            res, se, bounds = m.fit(X, y.copy(), estimate_SE=None)
        out.update({f"c{c}_meta": np.array([aniso, rot, pd, unit], dtype=np.int64), f"c{c}_v": np.float64(v),
                    f"c{c}_true": np.array(true), f"c{c}_X": X, f"c{c}_y": y, f"c{c}_cov": clean,
                    f"c{c}_nll": np.array([nll_true, nll_guess]), f"c{c}_fit_x": res.x,
                    f"c{c}_fit": np.array([res.fun, res.nit, res.nfev, float(res.success)]),
                    f"c{c}_bounds": np.array(bounds)})
        print(c, m.fform, "true", true, "fit", np.round(res.x, 4), "nit", res.nit, "nLL", round(float(res.fun), 6))
    np.savez_compressed(os.path.join(HERE, "ellipse_fit.npz"), **out)


if __name__ == "__main__":
    main()
