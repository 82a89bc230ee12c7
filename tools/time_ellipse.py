"""Time the ellipse tile kernel alone (CUDA events) for a few nu; used to compare kernel build variants
(GLOMAR_B200_LIB=<variant .so>)."""
import argparse, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from glomargridding_b200 import synthetic as synth
from glomargridding_b200._lib import check, lib
from glomargridding_b200.ellipse import EllipseCovarianceBuilder
ap = argparse.ArgumentParser()
ap.add_argument("--res", type=float, default=1.0)
ap.add_argument("--rows", type=int, default=0)
ap.add_argument("--row0", type=int, default=0, help="first row; not a multiple of 64 -> no symmetry, every entry computed")
ap.add_argument("--reps", type=int, default=5)
ap.add_argument("--nus", default="1.5,2.2")
a = ap.parse_args()
lats, lons = synth.global_grid(a.res)
Lx, Ly, th, sd = synth.ellipse_fields(lats, lons)
N = lats.size * lons.size
rows = a.rows or (N - a.row0)
out = torch.empty((rows, N), dtype=torch.float64, device="cuda")
L = lib()
res = {}
for nu in [float(x) for x in a.nus.split(",")]:
    b = EllipseCovarianceBuilder(Lx, Ly, th, sd, lats, lons, v=nu, precision=np.float64, rows=(0, 64))
    s = torch.cuda.current_stream().cuda_stream
    ts = []
    for i in range(a.reps + 2):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        check(L.gmg_ellipse_cov_f64(b._ws.data_ptr(), N, a.row0, rows, nu, -1.0, 0, out.data_ptr(), N, 0, s), "cov")
        e1.record(); e1.synchronize()
        ts.append(e0.elapsed_time(e1))
    t = float(np.median(ts[2:]))
    res[nu] = (t, rows * N / t / 1e6, 8 * rows * N / t / 1e6)
print(os.environ.get("GLOMAR_B200_LIB", "default"), " ".join(f"nu={k}: {v[0]:.2f} ms {v[1]:.0f} Gentries/s {v[2]:.0f} GB/s" for k, v in res.items()), flush=True)
