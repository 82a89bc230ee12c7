"""torchrun script: row-block sharded covariance build + column-sharded ordinary kriging (distributed.py).

  1. parity at 5 degrees: every rank's result equals the single-GPU OrdinaryKriging class (<= 1e-10 max-norm rel)
  2. timing at 1 degree, n = 20 000 (BASELINE configs[2] + configs[3]) with the matrix never leaving its row blocks

    torchrun --nproc-per-node G tools/dist_kriging_check.py [--res 1 --nobs 20000]
"""
import argparse, json, os, sys, time, warnings
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from glomargridding_b200 import synthetic as synth
from glomargridding_b200.distributed import krige_sharded
from glomargridding_b200.ellipse import EllipseCovarianceBuilder
from glomargridding_b200.kriging import OrdinaryKriging

ap = argparse.ArgumentParser()
ap.add_argument("--res", type=float, default=1.0)
ap.add_argument("--nobs", type=int, default=20000)
a = ap.parse_args()
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))


def block(N):
    tiles = (N + 127) // 128           # 128-row macro-tile alignment: the symmetric 2 x 2-tile enumeration applies
    t0, t1 = tiles * rank // world, tiles * (rank + 1) // world
    return t0 * 128, min(N, t1 * 128) - t0 * 128


def build(res):
    lats, lons = synth.global_grid(res)
    Lx, Ly, th, sd = synth.ellipse_fields(lats, lons)
    N = lats.size * lons.size
    r0, nrows = block(N)
    b = EllipseCovarianceBuilder(Lx, Ly, th, sd, lats, lons, v=1.5, delta_x_method="Met_Office",
                                 precision=np.float64, rows=(r0, nrows))
    return b, (Lx, Ly, th, sd, lats, lons), N, r0, nrows


# ---- 1. parity at 5 degrees -----------------------------------------------------------------------------------
b, fields, N, r0, nrows = build(5.0)
idx, obs = synth.fixed_observations(5, N, 900)
E = synth.hadcrut_style_error_cov(idx, fields[5].size)
res = krige_sharded(b.cov_ns_device, r0, idx, obs, E, "ordinary")
full = EllipseCovarianceBuilder(*fields, v=1.5, delta_x_method="Met_Office", precision=np.float64)
with warnings.catch_warnings():
    warnings.simplefilter("ignore")
    ok = OrdinaryKriging(full.cov_ns_device, idx, obs, E.copy())
    z, u, m = ok.solve(), ok.get_uncertainty(), ok.constraint_mask()
rel = lambda x, y: float(np.max(np.abs(x - y)) / np.max(np.abs(y)))
errs = {"field": rel(res["field"], z), "variance": rel(res["uncertainty"] ** 2, u ** 2), "mask": rel(res["constraint_mask"], m)}
assert max(errs.values()) <= 1e-10, errs
del full, ok, b
torch.cuda.empty_cache()

# ---- 2. timing at the requested size ----------------------------------------------------------------------------
t0 = time.perf_counter()
b, fields, N, r0, nrows = build(a.res)
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
t_build = time.perf_counter() - t0
idx, obs = synth.fixed_observations(4, N, a.nobs)
Ediag = np.random.default_rng(7).uniform(0.2, 0.4, a.nobs) ** 2
E = np.diag(Ediag)
times, times_repl = [], []
for rep in range(3):
    for flag, acc in ((True, times), (False, times_repl)):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        res_ = krige_sharded(b.cov_ns_device, r0, idx, obs, E, "ordinary", distributed_cholesky=flag)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        acc.append(time.perf_counter() - t0)
        if flag:
            res = res_
        else:
            assert np.array_equal(res["field"], res_["field"]) and np.array_equal(res["uncertainty"], res_["uncertainty"])
if rank == 0:
    print(json.dumps({"world": world, "parity_5deg_max_norm_rel": errs, "res": a.res, "N": N, "n_obs": a.nobs,
                      "cov_build_s (incl. host prep)": t_build, "krige_sharded_s": times, "krige_sharded_replicated_cholesky_s": times_repl,
                      "field_checksum": float(np.sum(res["field"])), "unc_mean": float(res["uncertainty"].mean())}))
if world > 1:
    dist.destroy_process_group()
