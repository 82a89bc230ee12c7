"""Config 5: 5-degree stochastic kriging, 200 ensemble members (HadCRUT-style error covariance) - wall time of
`batched.stochastic_ensemble` next to the reference's per-member cost (oracle: one SVD draw of the 2592 x 2592
covariance per member)."""
import json, os, sys, time, warnings
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from glomargridding_b200 import synthetic as synth
from glomargridding_b200.batched import stochastic_ensemble
from glomargridding_b200.ellipse import EllipseCovarianceBuilder

lats, lons = synth.global_grid(5.0)
Lx, Ly, th, sd = synth.ellipse_fields(lats, lons)
b = EllipseCovarianceBuilder(Lx, Ly, th, sd, lats, lons, v=1.5, delta_x_method="Met_Office", precision=np.float64)
cov = b.cov_ns_device
M = cov.shape[0]
idx, obs = synth.fixed_observations(201403, M, 1341)
E = synth.hadcrut_style_error_cov(idx, lons.size)
ts = []
for rep in range(4):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    r = stochastic_ensemble(cov, idx, obs, E.copy(), 200, seed=20140371)
    torch.cuda.synchronize(); ts.append(time.perf_counter() - t0)
out = {"members": 200, "M": M, "n_obs": 1341, "stochastic_ensemble_s": ts, "members_per_s": 200 / min(ts)}
if "--cpu" in sys.argv:
    from oracle import stochastic_kriging
    covh = b.cov_ns
    np.random.seed(1)
    t0 = time.perf_counter()
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        stochastic_kriging(covh, idx, obs, E.copy(), big_temporaries=False)
    out["reference_port_one_member_s"] = time.perf_counter() - t0
print(json.dumps(out))
