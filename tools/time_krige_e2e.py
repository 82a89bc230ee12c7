"""End-to-end timing of krige_batch (host idx / obs / error_cov per month -> numpy results) for 480 months,
the kriging `e2e` leg of bench.py, with a stage breakdown."""
import os, sys, time, warnings
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from glomargridding_b200 import synthetic as synth
from glomargridding_b200.batched import krige_batch
from glomargridding_b200.ellipse import EllipseCovarianceBuilder
months = int(sys.argv[1]) if len(sys.argv) > 1 else 480
lats, lons = synth.global_grid(5.0)
Lx, Ly, th, sd = synth.ellipse_fields(lats, lons)
cov = EllipseCovarianceBuilder(Lx, Ly, th, sd, lats, lons, v=1.5, delta_x_method="Met_Office",
                               precision=np.float64).cov_ns_device
probs = bench.month_problems(list(range(months)), cov.shape[0], lons.size)
ts = []
for _ in range(4):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        r = krige_batch(cov, probs, "ordinary")
    ts.append(time.perf_counter() - t0)
print("krige_batch e2e s:", [round(t, 4) for t in ts], "months/s:", round(months / min(ts[1:]), 1), "cores", os.cpu_count())
