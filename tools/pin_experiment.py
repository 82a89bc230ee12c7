"""Host -> device upload of many per-month error covariances (pageable numpy arrays, ~14 MB each): what does it cost to
(a) pack their lower block trapezoids into a pinned staging buffer with a thread pool (the shipped path),
(b) register each array in place (cudaHostRegister) and copy the blocks straight from user memory (cudaMemcpy2DAsync),
(c) copy from pageable memory directly?   One GPU; prints GB/s of E bytes moved."""
import ctypes, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from concurrent.futures import ThreadPoolExecutor

rt = torch.cuda.cudart()
n, B, NB = 1341, 60, 128
Es = [np.random.default_rng(i).standard_normal((n, n)) for i in range(B)]
blocks = [(r0, min(n, r0 + NB)) for r0 in range(0, n, NB)]
trap = sum((r1 - r0) * r1 for r0, r1 in blocks)
dev = torch.empty(B * trap, dtype=torch.float64, device="cuda")
stage = torch.empty(B * trap, dtype=torch.float64, pin_memory=True)
sn = stage.numpy()
torch.cuda.synchronize()

def pack(b):
    off = b * trap
    for r0, r1 in blocks:
        sz = (r1 - r0) * r1
        np.copyto(sn[off:off + sz].reshape(r1 - r0, r1), Es[b][r0:r1, :r1])
        off += sz

for threads in (1, 2, 4, 8, 15):
    with ThreadPoolExecutor(threads) as ex:
        list(ex.map(pack, range(B)))
        t0 = time.perf_counter()
        list(ex.map(pack, range(B)))
        t1 = time.perf_counter()
        dev.copy_(stage, non_blocking=True); torch.cuda.synchronize()
        t2 = time.perf_counter()
    print(f"pack {threads:2d} threads: {8e-9 * B * trap / (t1 - t0):6.1f} GB/s pack, H2D {8e-9 * B * trap / (t2 - t1):6.1f} GB/s", flush=True)

from glomargridding_b200._lib import lib as _lib
_L = _lib()
def pack_native(b):
    off = b * trap
    for r0, r1 in blocks:
        blk = Es[b][r0:r1, :r1]
        _L.gmg_host_pack_f64(blk.ctypes.data, 0, r1 - r0, r1, n, sn.ctypes.data + 8 * off)
        off += (r1 - r0) * r1
for threads in (1, 2, 4, 8, 15):
    with ThreadPoolExecutor(threads) as ex:
        list(ex.map(pack_native, range(B)))
        t0 = time.perf_counter()
        list(ex.map(pack_native, range(B)))
        t1 = time.perf_counter()
    print(f"native pack {threads:2d} threads: {8e-9 * B * trap / (t1 - t0):6.1f} GB/s", flush=True)
sys.exit(0)
# (b) register in place + 2D copies
lib = ctypes.CDLL("libcudart.so.12") if False else None
def reg_copy(b):
    E = Es[b]
    rc = rt.cudaHostRegister(E.ctypes.data, E.nbytes, 0)
    assert int(rc) == 0, rc
    off = b * trap
    for r0, r1 in blocks:
        rows, cols = r1 - r0, r1
        src = torch.from_numpy(E)[r0:r1, :r1]
        dev[off:off + rows * cols].view(rows, cols).copy_(src, non_blocking=True)
        off += rows * cols
for rep in range(2):
    t0 = time.perf_counter()
    for b in range(B): reg_copy(b)
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    for b in range(B): rt.cudaHostUnregister(Es[b].ctypes.data)
    t3 = time.perf_counter()
    print(f"register + strided copies: issue {t1 - t0:.3f} s, done {t2 - t0:.3f} s ({8e-9 * B * trap / (t2 - t0):.1f} GB/s), unregister {t3 - t2:.3f} s", flush=True)
# register cost alone
t0 = time.perf_counter()
for b in range(B): rt.cudaHostRegister(Es[b].ctypes.data, Es[b].nbytes, 0)
t1 = time.perf_counter()
for b in range(B): rt.cudaHostUnregister(Es[b].ctypes.data)
t2 = time.perf_counter()
print(f"cudaHostRegister {1e3 * (t1 - t0) / B:.3f} ms per {Es[0].nbytes / 1e6:.1f} MB array, unregister {1e3 * (t2 - t1) / B:.3f} ms", flush=True)
# (c) pageable whole-matrix copies
full = torch.empty((B, n, n), dtype=torch.float64, device="cuda")
t0 = time.perf_counter()
for b in range(B): full[b].copy_(torch.from_numpy(Es[b]), non_blocking=True)
torch.cuda.synchronize()
t1 = time.perf_counter()
print(f"pageable whole matrices: {8e-9 * B * n * n / (t1 - t0):.1f} GB/s of full-matrix bytes ({t1 - t0:.3f} s for {B} matrices)", flush=True)
