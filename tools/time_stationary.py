"""Stationary covariance route at 1 degree (64 800 cells would be 33.6 GB; 2 degrees = 16 200 cells here by default):
three passes (distance matrix, variogram, variance - variogram) against the fused single pass."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from glomargridding_b200 import synthetic as synth, variogram as V
from glomargridding_b200.distances import haversine_distance
res = float(sys.argv[1]) if len(sys.argv) > 1 else 2.0
lats, lons = synth.global_grid(res)
lon_g, lat_g = np.meshgrid(lons, lats)
la, lo = lat_g.ravel(), lon_g.ravel()
N = la.size
def timed(f, reps=3):
    best = 1e9
    for _ in range(reps):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); out = f(); e1.record(); e1.synchronize()
        best = min(best, e0.elapsed_time(e1)); del out
    return best
for nu in (1.5, 2.2):
    m = V.MaternVariogram(psill=1.2, nugget=0.05, range=1500, nu=nu)
    t3 = timed(lambda: V.variogram_to_covariance(m.fit(haversine_distance(la, lo, return_device=True)), 1.25))
    t1 = timed(lambda: V.stationary_covariance(la, lo, m, 1.25, return_device=True))
    print(f"N={N} nu={nu}: three passes {t3:.2f} ms, fused {t1:.2f} ms ({N * N * 8 / t1 / 1e6:.0f} GB/s of output)", flush=True)
