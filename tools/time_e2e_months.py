"""End-to-end months/s of krige_batch (host idx / obs / E in, numpy out) with the stage timeline of the last call:
GLOMAR_B200_TIMING=1 python tools/time_e2e_months.py [months] [float32]"""
import os, sys, time, warnings
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from glomargridding_b200 import synthetic as synth
from glomargridding_b200.batched import krige_batch
from glomargridding_b200.ellipse import EllipseCovarianceBuilder
from bench import month_problems
n_months = int(sys.argv[1]) if len(sys.argv) > 1 else 480
f32 = len(sys.argv) > 2 and sys.argv[2] == "float32"
lats, lons = synth.global_grid(5.0)
Lx, Ly, th, sd = synth.ellipse_fields(lats, lons)
cov = EllipseCovarianceBuilder(Lx, Ly, th, sd, lats, lons, v=1.5, delta_x_method="Met_Office", precision=np.float64).cov_ns_device
probs = month_problems(range(n_months), cov.shape[0], lons.size)
if f32:
    probs = [(i, o, E.astype(np.float32)) for i, o, E in probs]
timing = os.environ.pop("GLOMAR_B200_TIMING", None)
with warnings.catch_warnings():
    warnings.simplefilter("ignore")
    for rep in range(5):
        if rep == 4 and timing:
            os.environ["GLOMAR_B200_TIMING"] = timing
        torch.cuda.synchronize(); t0 = time.perf_counter()
        out = krige_batch(cov, probs, "ordinary")
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
        print(f"rep {rep}: {n_months / dt:.0f} months/s ({dt * 1e3:.1f} ms)", flush=True)
