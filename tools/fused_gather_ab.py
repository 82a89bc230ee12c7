"""480 (or argv[1]) months, inputs resident: KrigeBatch.run() with the right-hand-side gather folded into the solve steps
(gmg_krige_solve_f64) against the separate gather kernel + gmg_potrf_solve_f64."""
import os, sys, warnings
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from glomargridding_b200 import synthetic as synth, _dense as dn
from glomargridding_b200.ellipse import EllipseCovarianceBuilder
from bench import month_problems
n_months = int(sys.argv[1]) if len(sys.argv) > 1 else 480
lats, lons = synth.global_grid(5.0)
Lx, Ly, th, sd = synth.ellipse_fields(lats, lons)
cov = EllipseCovarianceBuilder(Lx, Ly, th, sd, lats, lons, v=1.5, delta_x_method="Met_Office", precision=np.float64).cov_ns_device
probs = month_problems(range(n_months), cov.shape[0], lons.size)
kb = dn.KrigeBatch(cov, probs, ordinary=True)
outs = {}
for rep in range(2):
    for fused in (True, False):
        dn.FUSED_RHS_GATHER = fused
        for _ in range(2):
            kb.run()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            kb.run()
        e1.record(); e1.synchronize()
        ms = e0.elapsed_time(e1) / 5
        outs[fused] = kb.out.clone()
        print(f"{n_months} months, fused gather {fused}: {ms:.2f} ms per batch = {n_months / ms * 1e3:.0f} months/s", flush=True)
print("bit-identical:", bool(torch.equal(outs[True], outs[False])))
