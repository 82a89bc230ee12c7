"""Is the tile kernel's sustained time set by the power cap?  Times every launch (CUDA events) of (a) isolated launches
with idle gaps, (b) 40 launches back to back, while nvidia-smi is sampled every 20 ms (SM clock, power, throttle reasons)."""
import os, subprocess, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from glomargridding_b200 import synthetic as synth
from glomargridding_b200._lib import check, lib
from glomargridding_b200.ellipse import EllipseCovarianceBuilder
nu = float(sys.argv[1]) if len(sys.argv) > 1 else 2.2
lats, lons = synth.global_grid(1.0)
Lx, Ly, th, sd = synth.ellipse_fields(lats, lons)
N = lats.size * lons.size
out = torch.empty((N, N), dtype=torch.float64, device="cuda")
L = lib()
b = EllipseCovarianceBuilder(Lx, Ly, th, sd, lats, lons, v=nu, precision=np.float64, rows=(0, 64))
s = torch.cuda.current_stream().cuda_stream
def launch():
    check(L.gmg_ellipse_cov_f64(b._ws.data_ptr(), N, 0, N, nu, -1.0, 0, out.data_ptr(), N, 0, s), "cov")
for _ in range(3):
    launch()
torch.cuda.synchronize()
smi = subprocess.Popen(["nvidia-smi", "--query-gpu=clocks.sm,power.draw,clocks_throttle_reasons.active", "--format=csv,noheader,nounits",
                        "-lms", "20", "-i", "0"], stdout=subprocess.PIPE, text=True)
time.sleep(1.0)
iso = []
for _ in range(8):
    time.sleep(0.25)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); launch(); e1.record(); e1.synchronize()
    iso.append(e0.elapsed_time(e1))
time.sleep(0.5)
t_mark = time.time()
ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(40)]
for e0, e1 in ev:
    e0.record(); launch(); e1.record()
torch.cuda.synchronize()
b2b = [e0.elapsed_time(e1) for e0, e1 in ev]
time.sleep(0.3)
smi.terminate()
lines = [l.strip() for l in smi.stdout.read().splitlines() if l.strip()]
print(f"nu={nu}: isolated launches (250 ms idle before each): " + " ".join(f"{t:.2f}" for t in iso))
print("back to back, launch 1..40: " + " ".join(f"{t:.2f}" for t in b2b))
rows = [l.split(", ") for l in lines]
clk = np.array([float(r[0]) for r in rows]); pw = np.array([float(r[1]) for r in rows])
print(f"nvidia-smi samples: {len(rows)}; SM clock min {clk.min():.0f} median {np.median(clk):.0f} max {clk.max():.0f} MHz; power min {pw.min():.0f} median {np.median(pw):.0f} max {pw.max():.0f} W")
busy = pw > 0.6 * pw.max()
if busy.any():
    print(f"samples above 60 % of max power: {int(busy.sum())}; SM clock there min {clk[busy].min():.0f} median {np.median(clk[busy]):.0f}; reasons seen: {sorted(set(r[2] for r, bb in zip(rows, busy) if bb))}")
