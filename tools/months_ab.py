"""Month-batch A/B across two builds of the library: `python tools/months_ab.py old` with the regular build, then
`GLOMAR_B200_LIB=variants/<name>/libglomar_b200.so python tools/months_ab.py new`.  480 and 60 synthetic months, inputs
resident (KrigeBatch.run()), timed; the result planes of the 60-month batch (field, variance, mask, multiplier) are
compared bit for bit with the other build's; one n = 8 000 overlapped factor + solve of 4 096 right-hand sides timed."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from glomargridding_b200 import synthetic as synth, _dense as dn
from glomargridding_b200.ellipse import EllipseCovarianceBuilder
from bench import month_problems
tag = sys.argv[1]
lats, lons = synth.global_grid(5.0)
Lx, Ly, th, sd = synth.ellipse_fields(lats, lons)
cov = EllipseCovarianceBuilder(Lx, Ly, th, sd, lats, lons, v=1.5, delta_x_method="Met_Office", precision=np.float64).cov_ns_device
for n_months in (480, 60):
    kb = dn.KrigeBatch(cov, month_problems(range(n_months), cov.shape[0], lons.size), ordinary=True)
    for _ in range(3):
        kb.run()
    best = 1e9
    for rep in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            kb.run()
        e1.record(); e1.synchronize()
        best = min(best, e0.elapsed_time(e1) / 5)
    print(f"{tag}: {n_months} months: {best:.2f} ms per batch = {n_months / best * 1e3:.0f} months/s", flush=True)
out = kb.out.cpu().numpy()
np.save(f"/tmp/months_{tag}.npy", out)
other = "/tmp/months_old.npy"
if tag != "old" and os.path.exists(other):
    print("60-month result planes bit-identical to the old build:", np.array_equal(np.load(other), out, equal_nan=True))
