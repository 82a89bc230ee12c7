"""VERDICT r01 task 10: the 5-degree months kriged against the DEFAULT displacement method's covariance, which is
indefinite - every month's system is rejected by Cholesky and goes through the blocked LU
(`gmg_gesv_blocked_f64`), like the reference's np.linalg.solve.  Device-resident and end-to-end months/s, the share
of systems that took the LU route, parity of one month against numpy."""
import os, sys, time, warnings
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from glomargridding_b200 import synthetic as synth
from glomargridding_b200.batched import krige_batch
from glomargridding_b200.ellipse import EllipseCovarianceBuilder
from bench import month_problems

n_months = int(sys.argv[1]) if len(sys.argv) > 1 else 60
lats, lons = synth.global_grid(5.0)
Lx, Ly, th, sd = synth.ellipse_fields(lats, lons)
b = EllipseCovarianceBuilder(Lx, Ly, th, sd, lats, lons, v=1.5, precision=np.float64)      # Modified_Met_Office
cov = b.cov_ns_device
ev = torch.linalg.eigvalsh(cov)
print(f"covariance {tuple(cov.shape)}: min eigenvalue {ev.min().item():.3e}, max {ev.max().item():.3e}", flush=True)
probs = month_problems(range(n_months), cov.shape[0], lons.size)
with warnings.catch_warnings():
    warnings.simplefilter("ignore")
    for rep in range(3):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        out = krige_batch(cov, probs, "ordinary")
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        print(f"end to end (host in, numpy out): {n_months / dt:.0f} months/s ({dt * 1e3:.1f} ms)", flush=True)
    # parity of month 0 against numpy's LU on the same system
    idx, obs, E = probs[0]
    C = cov.cpu().numpy()
    n = idx.size
    S = np.ones((n + 1, n + 1)); S[:n, :n] = C[np.ix_(idx, idx)] + E; S[n, n] = 0.0
    R = np.ones((n + 1, C.shape[0])); R[:n] = C[idx]
    W = np.linalg.solve(S, R)
    z = W[:n].T @ obs
    print("month 0 field vs numpy.linalg.solve: max rel", float(np.max(np.abs(out["field"][0] - z)) / np.max(np.abs(z))))
    t0 = time.perf_counter()
    for idx, obs, E in probs[:6]:
        n = idx.size
        S = np.ones((n + 1, n + 1)); S[:n, :n] = C[np.ix_(idx, idx)] + E; S[n, n] = 0.0
        R = np.ones((n + 1, C.shape[0])); R[:n] = C[idx]
        np.linalg.solve(S, R)
    print(f"numpy.linalg.solve alone on this host: {6 / (time.perf_counter() - t0):.1f} months/s")
