"""Time the kriging pipeline stages (CUDA events): 5-degree batched months and one large system."""
import argparse, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from glomargridding_b200 import synthetic as synth, _dense as dn, _device as dev
from glomargridding_b200._lib import check, lib
from glomargridding_b200.ellipse import EllipseCovarianceBuilder

ap = argparse.ArgumentParser()
ap.add_argument("--res", type=float, default=5.0)
ap.add_argument("--months", type=int, default=120)
ap.add_argument("--nobs", type=int, default=0, help=">0: one system with this many observations instead of months")
ap.add_argument("--reps", type=int, default=3)
a = ap.parse_args()
lats, lons = synth.global_grid(a.res)
Lx, Ly, th, sd = synth.ellipse_fields(lats, lons)
b = EllipseCovarianceBuilder(Lx, Ly, th, sd, lats, lons, v=1.5, delta_x_method="Met_Office", precision=np.float64)
cov = b.cov_ns_device
M = cov.shape[0]
if a.nobs:
    idx, obs = synth.fixed_observations(4, M, a.nobs)
    rng = np.random.default_rng(7)
    E = np.diag(rng.uniform(0.2, 0.4, a.nobs) ** 2)
    problems = [(idx, obs, E)]
else:
    problems = []
    for m in range(a.months):
        idx, obs = synth.month_observations(m, M)
        problems.append((idx, obs, synth.hadcrut_style_error_cov(idx, lons.size, seed=m)))
t0 = time.perf_counter()
kb = dn.KrigeBatch(cov, problems, ordinary=True)
torch.cuda.synchronize()
print(f"setup (alloc + H2D of idx/obs/E): {time.perf_counter() - t0:.3f} s for {len(problems)} problems, "
      f"n_max={kb.n_max}, mean n={np.mean(kb.ns):.0f}")
L, s = lib(), dev.stream_ptr()
C, ldc = cov.data_ptr(), cov.stride(0)
nrhs = M + 2
def ev():
    e = torch.cuda.Event(enable_timing=True); e.record(); return e
for rep in range(a.reps):
    kb.info.zero_()
    e0 = ev()
    check(L.gmg_krige_gather_batched_f64(C, M, ldc, kb._kd.data_ptr(), kb.B, kb.n_max, s), "gather")
    e1 = ev()
    check(L.gmg_potrf_batched_f64(kb._sd.data_ptr(), kb.B, kb.n_max, s), "potrf")
    e2 = ev()
    check(L.gmg_potrs_batched_f64(kb._sd.data_ptr(), kb.B, kb.n_max, nrhs, 1, s), "potrs")
    e3 = ev()
    check(L.gmg_krige_apply_batched_f64(C, M, ldc, kb._kd.data_ptr(), kb.B, kb.n_max, 1, s), "apply")
    e4 = ev()
    torch.cuda.synchronize()
    ns = np.array(kb.ns, dtype=np.float64)
    f_potrf, f_potrs = (ns ** 3 / 3).sum(), (ns ** 2 * nrhs).sum()   # forward solve only
    t = [e0.elapsed_time(e1), e1.elapsed_time(e2), e2.elapsed_time(e3), e3.elapsed_time(e4)]
    tot = sum(t)
    print(f"rep {rep}: gather {t[0]:.2f} ms | potrf {t[1]:.2f} ms ({f_potrf / t[1] / 1e9:.2f} TFLOP/s) | "
          f"fwd-trsm {t[2]:.2f} ms ({f_potrs / t[2] / 1e9:.2f} TFLOP/s) | apply {t[3]:.2f} ms | total {tot:.2f} ms "
          f"-> {len(problems) / tot * 1e3:.1f} problems/s, info max {int(kb.info.max())}", flush=True)
# the overlapped driver (gmg_potrf_solve_f64): factorisation + forward solve on three streams
for rep in range(a.reps):
    kb.info.zero_()
    check(L.gmg_krige_gather_batched_f64(C, M, ldc, kb._kd.data_ptr(), kb.B, kb.n_max, s), "gather")
    e1 = ev()
    check(L.gmg_potrf_solve_f64(None, kb._sd.data_ptr(), kb.B, kb.n_max, nrhs, None, 0, s), "potrf_solve")
    e2 = ev()
    torch.cuda.synchronize()
    t = e1.elapsed_time(e2)
    print(f"overlapped rep {rep}: factor + fwd solve {t:.2f} ms ({(f_potrf + f_potrs) / t / 1e9:.2f} TFLOP/s), "
          f"GMG_SMALL_K={os.environ.get('GMG_SMALL_K', 'default')}", flush=True)
