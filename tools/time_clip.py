"""Time eigenvalue_clip at 5 degrees (device) next to the reference algorithm (numpy eigh) on the host."""
import json, os, sys, time, warnings
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from glomargridding_b200 import synthetic as synth
from glomargridding_b200.covariance_tools import eigenvalue_clip
from glomargridding_b200.ellipse import EllipseCovarianceBuilder
lats, lons = synth.global_grid(5.0)
Lx, Ly, th, sd = synth.ellipse_fields(lats, lons)
b = EllipseCovarianceBuilder(Lx, Ly, th, sd, lats, lons, v=1.5, precision=np.float64)
C = b.cov_ns_device
ts = []
for _ in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    out = eigenvalue_clip(C)
    torch.cuda.synchronize(); ts.append(time.perf_counter() - t0)
covh = b.cov_ns
t0 = time.perf_counter()
w, v = np.linalg.eigh(covh)
csum = np.cumsum(w[::-1]); k = int(np.argmax(csum > 0.95 * w.sum())) + 1
new = w.copy(); new[: len(w) - k] = (w.sum() - w[len(w) - k:].sum()) / (len(w) - k)
ref = (v * new) @ v.T
_ = np.linalg.eigvalsh(ref)          # the reference's positive-definiteness check (:779)
t_ref = time.perf_counter() - t0
print(json.dumps({"n": int(C.shape[0]), "kept": k, "device_s": ts, "reference_algorithm_host_s": t_ref,
                  "max_abs_diff": float(np.max(np.abs(out.cpu().numpy() - ref)))}))
# phase breakdown of the spectral route
import glomargridding_b200.covariance_tools as ct
def timed(fn, reps=3):
    best = 1e9
    for _ in range(reps):
        torch.cuda.synchronize(); t0 = time.perf_counter(); r = fn(); torch.cuda.synchronize()
        best = min(best, time.perf_counter() - t0)
    return r, best
lam, t_eig = timed(lambda: ct.sym_eigvals(C))
we = np.linalg.eigvalsh(covh)
theta = lam.flip(0); csum = torch.cumsum(theta, 0); total = float(torch.diagonal(C).sum())
kk = int(torch.nonzero(csum > 0.95 * total)[0]) + 1
tau = float(0.5 * (theta[kk - 1] + theta[kk])); c = (total - float(csum[kk - 1])) / (C.shape[0] - kk)
(res, steps), t_proj = timed(lambda: ct._projector_clip(C, lam, tau, c))
print(json.dumps({"eigvals_s": t_eig, "eigvals_max_abs_err": float(np.max(np.abs(lam.cpu().numpy() - we))),
                  "k": kk, "rel_gap": float((theta[kk - 1] - theta[kk]) / theta[0]), "newton_schulz_steps": steps,
                  "projector_s": t_proj}))
