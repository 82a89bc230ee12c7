"""GEMM throughput of gmg_gemm_f64 (DMMA tiles) for square and kriging-shaped problems."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from glomargridding_b200 import _dense as dn

def t(m, n, k, ta=False, tb=False, reps=5):
    A = torch.randn((k, m) if ta else (m, k), dtype=torch.float64, device="cuda")
    B = torch.randn((n, k) if tb else (k, n), dtype=torch.float64, device="cuda")
    C = torch.zeros((m, n), dtype=torch.float64, device="cuda")
    for _ in range(2):
        dn.gemm(A, B, ta, tb, out=C)
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); dn.gemm(A, B, ta, tb, out=C); e1.record(); e1.synchronize()
        ts.append(e0.elapsed_time(e1))
    ms = float(np.median(ts))
    print(f"m={m:6d} n={n:6d} k={k:6d} ta={int(ta)} tb={int(tb)}: {ms:8.3f} ms  {2.0 * m * n * k / ms / 1e9:6.2f} TFLOP/s", flush=True)

for s in (2048, 4096, 8192):
    t(s, s, s)
t(4096, 4096, 4096, False, True)
t(4096, 4096, 4096, True, False)
for k in (128, 256, 512, 1024, 2048, 8192):
    t(128, 64800, k)          # one trsm step at 1 degree
for k in (128, 512, 1536):
    t(128, 2592 * 48, k)      # ~48 months' worth of 5-degree right-hand sides in one launch
t(18944, 18944, 128, False, True)   # potrf trailing update shape at n = 20 000
