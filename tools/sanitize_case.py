"""Small end-to-end case for compute-sanitizer (tools/sanitize.sh): the tile kernel with its barrier-free mbarrier
record pipeline (macro tiles, mirrored stores, general nu and closed form, float64 and float32 stores), and the
kriging chain (gather -> overlapped Cholesky + forward solve on three streams -> apply, plus the backward solve).
Sizes are tiny: the sanitizer serialises and instruments every access."""
import os
import sys
import warnings

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from glomargridding_b200 import synthetic as synth  # noqa: E402
from glomargridding_b200.ellipse import EllipseCovarianceBuilder  # noqa: E402
from glomargridding_b200.kriging import OrdinaryKriging  # noqa: E402

res = float(sys.argv[1]) if len(sys.argv) > 1 else 10.0
lats, lons = synth.global_grid(res)
Lx, Ly, th, sd = synth.ellipse_fields(lats, lons)
for v, prec in ((2.2, np.float64), (1.5, np.float64), (0.5, np.float32)):
    b = EllipseCovarianceBuilder(Lx, Ly, th, sd, lats, lons, v=v, precision=prec)
    c = b.cov_ns_device
    assert torch.equal(c, c.T)
    b2 = EllipseCovarianceBuilder(Lx, Ly, th, sd, lats, lons, v=v, precision=prec, rows=(128, 200))
    assert torch.equal(b2.cov_ns_device, c[128:328])
b = EllipseCovarianceBuilder(Lx, Ly, th, sd, lats, lons, v=1.5, delta_x_method="Met_Office", precision=np.float64)
C = b.cov_ns_device
idx, obs = synth.fixed_observations(3, C.shape[0], min(300, C.shape[0] // 2))
E = synth.hadcrut_style_error_cov(idx, lons.size)
with warnings.catch_warnings():
    warnings.simplefilter("ignore")
    k = OrdinaryKriging(C, idx, obs, E)
    z, u, m = k.solve(), k.get_uncertainty(), k.constraint_mask()
    W = k.kriging_weights
assert np.isfinite(z).all() and np.isfinite(u).all() and W.shape == (C.shape[0], idx.size + 1)
torch.cuda.synchronize()
print("sanitize_case ok", C.shape, idx.size)
