"""One rank's share of the column-sharded 1-degree solve on ONE GPU: the forward solve of n = 20 000 against `--cols`
right-hand sides (8 104 = one rank of 8) with the strip width forced by GMG_STRIP (128 / 64 / 56) or chosen by the
library.  Times gmg_potrs_batched_f64(FORWARD) alone (CUDA events)."""
import argparse, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from glomargridding_b200 import _dense as dn, _device as dev
from glomargridding_b200._lib import check, lib
ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=20000)
ap.add_argument("--cols", type=int, default=8104)
a = ap.parse_args()
n = a.n
g = torch.Generator(device="cuda").manual_seed(1)
A = torch.randn((n, 256), dtype=torch.float64, device="cuda", generator=g)
S = dn.padded(A @ A.T / 256 + 4 * torch.eye(n, dtype=torch.float64, device="cuda"))
ws = dn.factor_inplace(S, n)
R0 = torch.randn((n, a.cols + (a.cols & 1)), dtype=torch.float64, device="cuda", generator=g)
L, s = lib(), dev.stream_ptr()
ts = []
for rep in range(3):
    R = R0.clone()
    d = np.zeros(1, dtype=dn.SOLVE_DESC)
    d[0] = (S.data_ptr(), S.stride(0), ws["Linv"].data_ptr(), R.data_ptr(), R.stride(0), ws["P"].data_ptr(), ws["ldp"],
            ws["info"].data_ptr(), n, a.cols)
    dt = dn._desc_tensor(d, "cuda")
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    check(L.gmg_potrs_batched_f64(dt.data_ptr(), 1, n, a.cols, 1, s), "potrs")
    e1.record(); e1.synchronize()
    ts.append(e0.elapsed_time(e1))
t = float(np.median(ts))
print(f"GMG_STRIP={os.environ.get('GMG_STRIP', 'auto')}: n={n} cols={a.cols}: {t:.1f} ms, {float(n) * n * a.cols / t / 1e9:.2f} TFLOP/s, checksum {float(R[:, :a.cols].sum()):.6e}", flush=True)
