"""Small driver for ncu: one 1-degree row block of the ellipse covariance (nu and rows selectable)."""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from glomargridding_b200 import synthetic as synth  # noqa: E402
from glomargridding_b200.ellipse import EllipseCovarianceBuilder  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--nu", type=float, default=2.2)
ap.add_argument("--rows", type=int, default=8192)
ap.add_argument("--res", type=float, default=1.0)
ap.add_argument("--reps", type=int, default=3)
a = ap.parse_args()
lats, lons = synth.global_grid(a.res)
Lx, Ly, th, sd = synth.ellipse_fields(lats, lons)
import torch  # noqa: E402

for _ in range(a.reps):
    b = EllipseCovarianceBuilder(Lx, Ly, th, sd, lats, lons, v=a.nu, precision=np.float64, rows=(0, a.rows))
torch.cuda.synchronize()
print("ok", b.cov_ns_device.shape, float(b.cov_ns_device[0, 1]))
