"""Host packing bandwidth of the kriging upload path on this box: np.copyto of error-covariance trapezoids into a
pinned staging buffer, by thread count; plus the raw pinned H2D rate.  (Diagnostic for `_dense._h2d_concat`.)"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from concurrent.futures import ThreadPoolExecutor

rng = np.random.default_rng(0)
Es = []
for m in range(60):
    n = int(rng.integers(600, 1600))
    E = rng.standard_normal((n, n)); Es.append(E + E.T)
NB = 128
groups, off = [], 0
for E in Es:
    n = E.shape[0]
    bl = [E[r0:min(n, r0 + NB), :min(n, r0 + NB)] for r0 in range(0, n, NB)]
    groups.append((off, bl)); off += sum(b.size for b in bl)
total = off
pinned = torch.empty(total, dtype=torch.float64, pin_memory=True)
buf = pinned.numpy()
plain = np.empty(total)
plain[:] = 0

def make_pack(dst):
    def pack(item):
        o, g = item
        for a in g:
            np.copyto(dst[o:o + a.size].reshape(a.shape), a, casting="unsafe"); o += a.size
    return pack

for name, dst in (("pinned", buf), ("pageable", plain)):
    for nt in (1, 2, 4, 8, 14, 16):
        pool = ThreadPoolExecutor(max_workers=nt)
        best = 1e9
        for rep in range(4):
            t = time.perf_counter(); list(pool.map(make_pack(dst), groups)); best = min(best, time.perf_counter() - t)
        print(f"{name} threads {nt:2d}: {8 * total / best / 1e9:6.1f} GB/s ({1e3 * best:.1f} ms for {8 * total / 1e6:.0f} MB)", flush=True)
        pool.shutdown()
# finer tasks: one per block
blocks = []
for o, g in groups:
    for a in g:
        blocks.append((o, [a])); o += a.size
pool = ThreadPoolExecutor(max_workers=14)
best = 1e9
for rep in range(4):
    t = time.perf_counter(); list(pool.map(make_pack(buf), blocks)); best = min(best, time.perf_counter() - t)
print(f"per-block tasks, 14 threads: {8 * total / best / 1e9:.1f} GB/s")
# contiguous memcpy of whole matrices (no trapezoid): upper bound of numpy copies
full = [(0, [E]) for E in Es]
o = 0; full = []
for E in Es:
    if o + E.size > total: break
    full.append((o, [E])); o += E.size
best = 1e9
for rep in range(4):
    t = time.perf_counter(); list(pool.map(make_pack(buf), full)); best = min(best, time.perf_counter() - t)
print(f"whole matrices, 14 threads: {8 * o / best / 1e9:.1f} GB/s")
d = torch.empty(total, dtype=torch.float64, device="cuda")
for rep in range(3):
    torch.cuda.synchronize(); t = time.perf_counter(); d.copy_(pinned, non_blocking=True); torch.cuda.synchronize()
    print(f"pinned H2D: {8 * total / (time.perf_counter() - t) / 1e9:.1f} GB/s")
print("cpus", os.cpu_count(), "affinity", len(os.sched_getaffinity(0)))
os.system("lscpu | egrep 'Model name|Thread|Core|Socket|NUMA node\\(s\\)' ")
