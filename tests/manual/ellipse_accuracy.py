"""Max entrywise relative error of the tile kernel against (a) the reference's own sampled pairs at 1 and 5 degrees
(tests/golden/ellipse_pairs_*.npz) and (b) the oracle's full 5-degree matrix - the margin to the 1e-10 tolerance.
Used to judge kernel variants (GLOMAR_B200_LIB=<variant .so>)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from conftest import load_golden, max_rel
from glomargridding_b200 import synthetic as synth
from glomargridding_b200.ellipse import EllipseCovarianceBuilder
from oracle import EllipseOracle
out = {}
for res, tag in [(5.0, "5deg"), (1.0, "1deg")]:
    g = load_golden(f"ellipse_pairs_{tag}.npz")
    lats, lons = synth.global_grid(res)
    Lx, Ly, th, sd = synth.ellipse_fields(lats, lons)
    gi, gj = g["i"].astype(np.int64), g["j"].astype(np.int64)
    for name, v, meth in [("nu2.2_mod", 2.2, "Modified_Met_Office"), ("nu1.5_mod", 1.5, "Modified_Met_Office"),
                          ("nu0.5_mo", 0.5, "Met_Office"), ("nu2.5_mo", 2.5, "Met_Office"),
                          ("nu0.3_mod", 0.3, "Modified_Met_Office"), ("nu3.0_mo", 3.0, "Met_Office")]:
        if res == 1.0:
            b = EllipseCovarianceBuilder(Lx, Ly, th, sd, lats, lons, v=v, delta_x_method=meth, precision=np.float64,
                                         rows=(0, 64))
            got = b.c_ij_anisotropic_array(gi, gj)
        else:
            C = EllipseCovarianceBuilder(Lx, Ly, th, sd, lats, lons, v=v, delta_x_method=meth,
                                         precision=np.float64).cov_ns
            got = C[gi, gj]
        out[f"{tag}_{name}"] = max_rel(got, g[name])
lats, lons = synth.global_grid(5.0)
Lx, Ly, th, sd = synth.ellipse_fields(lats, lons)
for v in (1.5, 2.2):
    want = EllipseOracle(Lx, Ly, th, sd, lats, lons, v, "Modified_Met_Office").build()
    got = EllipseCovarianceBuilder(Lx, Ly, th, sd, lats, lons, v=v, precision=np.float64).cov_ns
    out[f"5deg_full_nu{v}"] = max_rel(got, want)
print(os.environ.get("GLOMAR_B200_LIB", "default"), {k: float(f"{v:.2e}") for k, v in out.items()}, flush=True)
