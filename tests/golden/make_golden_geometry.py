"""Golden vectors for the module-level geometry helpers of the ellipse covariance path, produced by the UNMODIFIED
reference (`glomar_gridding/ellipse/covariance.py:512-654`, imported through `oracle/reference_shim.py`).

    python tests/golden/make_golden_geometry.py        # build container only (needs /root/reference)

Inputs are seeded; the point pairs include the cases the reference's own tests probe (random pairs over the globe,
`test/test_matrix.py:74-140`) plus the edges that decide parity: longitude differences just either side of +-pi (the
strict wrap), exactly +-pi, coincident points and the poles.
"""

from __future__ import annotations

import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle.reference_shim import load_reference  # noqa: E402

cov = load_reference().ellipse_covariance


def inputs():
    rng = np.random.default_rng(512)
    n = 1024
    lat0 = np.deg2rad(90 - 180 * rng.random(n))
    lon0 = np.deg2rad(180 - 360 * rng.random(n))
    lat1 = np.deg2rad(90 - 180 * rng.random(n))
    lon1 = np.deg2rad(180 - 360 * rng.random(n))
    # edges: dlon = +-pi exactly, one ulp either side, coincident points, poles
    e_lon0 = np.array([np.pi / 2, -np.pi / 2, np.nextafter(np.pi / 2, 4), np.nextafter(np.pi / 2, 0), 0.3, 0.0, 3.0, -3.0])
    e_lon1 = np.array([-np.pi / 2, np.pi / 2, -np.pi / 2, -np.pi / 2, 0.3, 0.0, -3.0, 3.0])
    e_lat0 = np.array([0.1, -0.2, 0.3, 0.4, 0.5, np.pi / 2, -1.0, 1.2])
    e_lat1 = np.array([0.2, 0.2, -0.3, 0.4, 0.5, -np.pi / 2, 1.0, 1.2])
    Lx = 600 + 2400 * rng.random(n)
    Ly = 300 + 900 * rng.random(n)
    theta = np.pi * (rng.random(n) - 0.5)
    return (np.concatenate([lat0, e_lat0]), np.concatenate([lon0, e_lon0]), np.concatenate([lat1, e_lat1]),
            np.concatenate([lon1, e_lon1]), Lx, Ly, theta)


def main():
    lat0, lon0, lat1, lon1, Lx, Ly, theta = inputs()
    sig = cov._sigma_rot_func_multi(Lx, Ly, theta)
    out = {
        "lat0": lat0, "lon0": lon0, "lat1": lat1, "lon1": lon1, "Lx": Lx, "Ly": Ly, "theta": theta,
        "sigmas": sig, "dets": cov._det_22_multi(sig),
        "haversine": cov._haversine_multi(lat0, lon0, lat1, lon1),
        "mo": np.stack(cov._mo_disp_multi(lat0, lon0, lat1, lon1)),
        "mod_mo": np.stack(cov._mod_mo_disp_multi(lat0, lon0, lat1, lon1)),
    }
    k = 8
    out["single_haversine"] = np.array([cov._haversine_single(*map(float, (lat0[i], lon0[i], lat1[i], lon1[i])))
                                        for i in range(-k, 0)], dtype=np.float64)
    out["single_mod_mo"] = np.array([cov._mod_mo_disp_single(*map(float, (lat0[i], lon0[i], lat1[i], lon1[i])))
                                     for i in range(-k, 0)])
    out["single_mo"] = np.array([cov._mo_disp_single(*map(float, (lat0[i], lon0[i], lat1[i], lon1[i])))
                                 for i in range(-k, 0)])
    out["single_det"] = np.array([cov._det_22_single(sig[i]) for i in range(k)])
    np.savez_compressed(os.path.join(HERE, "geometry_small.npz"), **out)
    print({k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
