"""Latency of ONE drop-in class call at 5 degrees (BASELINE configs[0] shape: 2592 cells, ~1341 observations):
OrdinaryKriging(cov on device, host idx / obs / error_cov) -> solve() + get_uncertainty() + constraint_mask()."""
import os, sys, time, warnings
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from glomargridding_b200 import synthetic as synth
from glomargridding_b200.ellipse import EllipseCovarianceBuilder
from glomargridding_b200.kriging import OrdinaryKriging, SimpleKriging

lats, lons = synth.global_grid(5.0)
Lx, Ly, th, sd = synth.ellipse_fields(lats, lons)
cov = EllipseCovarianceBuilder(Lx, Ly, th, sd, lats, lons, v=1.5, delta_x_method="Met_Office", precision=np.float64).cov_ns_device
idx, obs = synth.fixed_observations(2014, cov.shape[0], 1341)
E = synth.hadcrut_style_error_cov(idx, lons.size, seed=3)
for cls in (OrdinaryKriging, SimpleKriging):
    ts = []
    for rep in range(8):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            k = cls(cov, idx, obs, E)
            t1 = time.perf_counter()
            z = k.solve()
            t2 = time.perf_counter()
            u = k.get_uncertainty()
            m = k.constraint_mask()
        torch.cuda.synchronize(); t3 = time.perf_counter()
        ts.append((t1 - t0, t2 - t1, t3 - t2))
    a = np.array(ts[3:]) * 1e3
    print(f"{cls.__name__}: construct {a[:, 0].mean():.2f} ms, solve {a[:, 1].mean():.2f} ms, uncertainty + mask {a[:, 2].mean():.2f} ms, "
          f"total {a.sum(1).mean():.2f} ms (first call {1e3 * sum(ts[0]):.1f} ms)", flush=True)
