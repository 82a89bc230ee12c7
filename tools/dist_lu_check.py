"""torchrun script: `krige_sharded` against the row blocks of an INDEFINITE covariance (the default displacement method,
5 degrees) - every rank must take the LU route together and agree with the single-GPU class (which also falls to LU).

    torchrun --nproc-per-node G tools/dist_lu_check.py
"""
import os, sys, warnings
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from glomargridding_b200 import synthetic as synth
from glomargridding_b200.distributed import krige_sharded
from glomargridding_b200.ellipse import EllipseCovarianceBuilder
from glomargridding_b200.kriging import OrdinaryKriging

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
lats, lons = synth.global_grid(5.0)
Lx, Ly, th, sd = synth.ellipse_fields(lats, lons)
N = lats.size * lons.size
tiles = (N + 127) // 128
t0, t1 = tiles * rank // world, tiles * (rank + 1) // world
r0, nrows = t0 * 128, min(N, t1 * 128) - t0 * 128
b = EllipseCovarianceBuilder(Lx, Ly, th, sd, lats, lons, v=1.5, precision=np.float64, rows=(r0, nrows))
idx, obs = synth.fixed_observations(7, N, 700)
E = synth.hadcrut_style_error_cov(idx, lons.size)
with warnings.catch_warnings(record=True) as w:
    warnings.simplefilter("always")
    res = krige_sharded(b.cov_ns_device, r0, idx, obs, E, "ordinary")
took_lu = any("LU with partial pivoting" in str(x.message) for x in w)
full = EllipseCovarianceBuilder(Lx, Ly, th, sd, lats, lons, v=1.5, precision=np.float64)
with warnings.catch_warnings():
    warnings.simplefilter("ignore")
    ok = OrdinaryKriging(full.cov_ns_device, idx, obs, E.copy())
    z, m = ok.solve(), ok.constraint_mask()
rel = lambda x, y: float(np.max(np.abs(x - y)) / np.max(np.abs(y)))
errs = {"field": rel(res["field"], z), "mask": rel(res["constraint_mask"], m)}
print(f"rank {rank}/{world}: LU route {took_lu}, vs the single-GPU class {errs}", flush=True)
assert took_lu and max(errs.values()) <= 1e-8, errs
if world > 1:
    dist.destroy_process_group()
