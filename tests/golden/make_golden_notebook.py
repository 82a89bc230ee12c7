"""configs[0] fixture: the reference's own CPU case - 5-degree global grid, simple / ordinary kriging of the
HadSST.4.0.1.0 2014-03 ensemble-member field of `notebooks/Ellipse_to_grid.ipynb` (cells 19-34) with the
HadCRUT.5.0.2.0 error covariance of the same month.

Run in the build container (needs `/root/reference`):

    python tests/golden/make_golden_notebook.py

Inputs (decoded from the notebook's netCDF-4 files by `decode_notebook_nc.py`, no HDF5 library needed):
  * `tos` (36 x 72 float32, latitude ascending from -87.5): 1 341 valid cells -> `obs`, `idx` = row-major flat index
    (first three 227 / 228 / 229 = -0.280791 / -0.704167 / 0.754848, the frame printed at notebook line 1601);
  * `tas_cov` (2592 x 2592 float32, fill > 1e6 -> 0, cell 17) + diag(`tas_unc`^2) (fill -> 0, cells 18-19)
    -> the full-grid float32 error covariance, stored here as its non-zero entries.
The spatial covariance is the 5-degree ellipse covariance of the synthetic smooth fields (`synthetic.ellipse_fields`,
nu = 1.5, default displacement method) passed through the reference's `eigenvalue_clip` exactly as notebook cell 20
does - fitting the ellipses to the ESA-CCI cube is outside the hot path.
Outputs: the unmodified reference's SimpleKriging / OrdinaryKriging field, uncertainty and constraint mask.
"""

from __future__ import annotations

import contextlib
import io
import os
import sys
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, HERE)
sys.path.insert(0, ROOT)

from decode_notebook_nc import read_chunked_f32  # noqa: E402
from glomargridding_b200 import synthetic as synth  # noqa: E402
from oracle.reference_shim import REFERENCE_ROOT, load_reference  # noqa: E402

ref = load_reference()
NB = os.path.join(REFERENCE_ROOT, "notebooks")


def quiet(fn, *a, **k):
    with contextlib.redirect_stdout(io.StringIO()), warnings.catch_warnings():
        warnings.simplefilter("ignore")
        return fn(*a, **k)


def notebook_inputs():
    tos = read_chunked_f32(os.path.join(NB, "HadSST.4.0.1.0_ensemble_member_71_2014_03.nc"), (36, 72), (36, 72))
    flat = tos.reshape(-1)
    idx = np.flatnonzero(np.abs(flat) < 1e4).astype(np.int64)
    obs = flat[idx].astype(np.float32)
    E = read_chunked_f32(os.path.join(NB, "HadCRUT.5.0.2.0.error_covariance.2014_03.nc"),
                         (1, 2592, 2592), (1, 1296, 1296))[0]
    E[E > 1e6] = 0.0                                                      # notebook cell 17
    unc = read_chunked_f32(os.path.join(NB, "HadCRUT.5.0.2.0.uncorrelated_2014_03.nc"), (36, 72), (36, 72))
    unc = unc.reshape(2592)
    unc[unc > 1e6] = 0.0                                                  # cell 18
    E += np.diag(np.power(unc, 2))                                        # cell 19 (float32)
    return idx, obs, E


def main():
    idx, obs, E = notebook_inputs()
    assert idx.size == 1341 and tuple(idx[:3]) == (227, 228, 229)
    assert np.allclose(obs[:3], [-0.280791, -0.704167, 0.754848], atol=1e-6)
    lats, lons = synth.global_grid(5.0)
    Lx, Ly, th, sd = synth.ellipse_fields(lats, lons)
    cov_raw = quiet(ref.EllipseCovarianceBuilder, Lx, Ly, th, sd, lats, lons, v=1.5, precision=np.float64).cov_ns
    cov = quiet(ref.covariance_tools.eigenvalue_clip, cov_raw)            # notebook cell 20
    out = {"idx": idx, "obs": obs}
    r, c = np.nonzero(E)
    out["E_rows"], out["E_cols"], out["E_vals"] = r.astype(np.int32), c.astype(np.int32), E[r, c]
    out["cov_min_eig_raw"] = np.linalg.eigvalsh(cov_raw).min()
    out["cov_min_eig_clipped"] = np.linalg.eigvalsh(cov).min()
    out["cov_diag_clipped"] = np.diag(cov).copy()
    sk = quiet(ref.SimpleKriging, cov, idx.copy(), obs.copy(), E.copy())
    out["sk_field"] = quiet(sk.solve)
    out["sk_unc"] = quiet(sk.get_uncertainty)
    out["sk_mask"] = quiet(sk.constraint_mask)
    ok = quiet(ref.OrdinaryKriging, cov, idx.copy(), obs.copy(), E.copy())
    out["ok_field"] = quiet(ok.solve)
    out["ok_unc"] = quiet(ok.get_uncertainty)
    out["ok_mask"] = quiet(ok.constraint_mask)
    out["kept_idx"] = ok.idx
    np.savez_compressed(os.path.join(HERE, "notebook_2014_03.npz"), **out)
    print({k: (v.shape, v.dtype) for k, v in out.items() if hasattr(v, "shape")})
    print("min eig raw / clipped:", out["cov_min_eig_raw"], out["cov_min_eig_clipped"])


if __name__ == "__main__":
    main()
