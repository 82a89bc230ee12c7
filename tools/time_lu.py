"""Blocked LU (gmg_gesv_blocked_f64) timing: one large system, and a batch of month-sized systems."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from glomargridding_b200 import _dense as dn, _device as dev
from glomargridding_b200._lib import check, lib
def run(sizes, nrhs):
    B, n_max = len(sizes), max(sizes)
    lin, pdb, ldp = dn.gesv_ws(n_max)
    g = torch.Generator(device="cuda").manual_seed(1)
    keep, d = [], np.zeros(B, dtype=dn.SOLVE_DESC)
    info = torch.zeros(B, dtype=torch.int32, device="cuda")
    for b, n in enumerate(sizes):
        S = dn.padded(torch.randn((n, n), dtype=torch.float64, device="cuda", generator=g))
        R = dn.padded(torch.randn((n, nrhs), dtype=torch.float64, device="cuda", generator=g))
        li = torch.empty(lin, dtype=torch.float64, device="cuda"); st = torch.empty(pdb, dtype=torch.float64, device="cuda")
        keep += [S, R, li, st]
        d[b] = (S.data_ptr(), S.stride(0), li.data_ptr(), R.data_ptr(), R.stride(0), st.data_ptr(), ldp, info.data_ptr() + 4 * b, n, nrhs)
    dt = dn._desc_tensor(d, "cuda")
    for what, nr in (("factor", 0), ("factor + solve", nrhs)):
        for rep in range(2):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            check(lib().gmg_gesv_blocked_f64(dt.data_ptr(), B, n_max, nr, dev.stream_ptr()), "gesv")
            e1.record(); e1.synchronize()
        ns = np.array(sizes, dtype=np.float64)
        fl = (2 * ns ** 3 / 3).sum() + (2 * ns ** 2 * nr).sum()
        print(f"{B} system(s), n_max={n_max}, nrhs={nr}: {what} {e0.elapsed_time(e1):.1f} ms = {fl / e0.elapsed_time(e1) / 1e9:.2f} TFLOP/s", flush=True)
if len(sys.argv) > 1:                                   # e.g.  time_lu.py 8000 1 2594
    run([int(sys.argv[1])] * int(sys.argv[2]), int(sys.argv[3]))
else:
    run([1341] * 60, 2594)
    run([8000], 2594)
    run([20000], 1024)
