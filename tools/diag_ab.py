"""Diagonal-block kernel A/B against another build (GLOMAR_B200_LIB=<old .so> python tools/diag_ab.py old; then
python tools/diag_ab.py new): factors + inverted diagonal blocks of a ragged batch compared bit for bit, the
factorisation of 60 month-sized systems and of one n = 8 000 system timed."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from glomargridding_b200 import _dense as dn, _device as dev
from glomargridding_b200._lib import check, lib
tag = sys.argv[1]
L, s = lib(), dev.stream_ptr()
def batch(sizes, seed):
    g = torch.Generator(device="cuda").manual_seed(seed)
    keep = []
    d = np.zeros(len(sizes), dtype=dn.SOLVE_DESC)
    info = torch.zeros(len(sizes), dtype=torch.int32, device="cuda")
    for b, n in enumerate(sizes):
        A = torch.randn((n, n + 3), dtype=torch.float64, device="cuda", generator=g)
        S = dn.padded(A @ A.T / n + torch.eye(n, dtype=torch.float64, device="cuda"))
        lin, pdb, ldp = dn.solve_ws(n)
        li = torch.zeros(lin, dtype=torch.float64, device="cuda"); st = torch.zeros(pdb, dtype=torch.float64, device="cuda")
        keep += [S, li, st]
        d[b] = (S.data_ptr(), S.stride(0), li.data_ptr(), 0, 2, st.data_ptr(), ldp, info.data_ptr() + 4 * b, n, 0)
    dt = dn._desc_tensor(d, "cuda")
    orig = [keep[3 * b].clone() for b in range(len(sizes))]
    ts = []
    for rep in range(4):
        for b in range(len(sizes)):
            keep[3 * b].copy_(orig[b])
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        check(L.gmg_potrf_solve_f64(None, dt.data_ptr(), len(sizes), max(sizes), 0, None, 0, s), "potrf")
        e1.record(); e1.synchronize()
        ts.append(e0.elapsed_time(e1))
    assert int(info.abs().sum()) == 0
    return keep, min(ts[1:])
sizes = [1, 17, 128, 129, 300, 1341, 1000] + [1341] * 53
keep, t = batch(sizes, 7)
out = {f"S{b}": np.tril(keep[3 * b][:, :sizes[b]].cpu().numpy()) for b in range(len(sizes))}
out.update({f"L{b}": keep[3 * b + 1].cpu().numpy() for b in range(8)})
np.savez(f"gpurun_out/diag_{tag}.npz", **out)
print(f"{tag}: factorisation of {len(sizes)} systems: {t:.3f} ms", flush=True)
_, t1 = batch([8000], 3)
print(f"{tag}: factorisation of one n = 8000 system: {t1:.3f} ms", flush=True)
if os.path.exists("gpurun_out/diag_old.npz") and tag != "old":
    o = np.load("gpurun_out/diag_old.npz")
    print("bit-identical to the previous build:", all(np.array_equal(o[k], out[k]) for k in out))
