"""BASELINE configs[3] at full size - 1-degree grid (64 800 cells), 20 000 observations, ordinary kriging - against a
vector held by the UNMODIFIED reference.

`tests/golden/krige_1deg_n20000.npz` was produced by `tests/golden/make_golden_1deg_krige.py`: the reference's
`OrdinaryKriging` (kriging.py:540-776) on the reduced grid {20 000 observed cells} U {64 unobserved cells}, whose
covariance comes from the reference's `c_ij_anisotropic_array` on the reference builder's compressed state.  Kriging
at a cell depends only on the observations and that cell, so the full-grid device results at those 20 064 cells must
agree with it: field, kriging variance and constraint mask within 1e-8 max-norm relative (north_star).

The device side runs what a user runs: the full 1-degree covariance built on the GPU (33.6 GB), then (a) the
`OrdinaryKriging` class and (b) `krige_sharded` - the routine the 8-GPU configuration uses - on one rank.
"""

import numpy as np
import pytest
from conftest import load_golden, max_norm_rel

from glomargridding_b200 import synthetic as synth

pytestmark = pytest.mark.gpu
TOL = 1e-8


@pytest.fixture(scope="module")
def case():
    import torch

    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    if torch.cuda.mem_get_info()[0] < 70e9:
        pytest.skip("needs ~60 GB of device memory")
    from glomargridding_b200.ellipse import EllipseCovarianceBuilder

    g = load_golden("krige_1deg_n20000.npz")
    lats, lons = synth.global_grid(1.0)
    N = lats.size * lons.size
    idx, obs = synth.fixed_observations(4, N, int(g["n_obs"]))
    assert np.isin(idx, g["cells"]).all() and g["cells"].size == idx.size + 64
    Lx, Ly, th, sd = synth.ellipse_fields(lats, lons)
    b = EllipseCovarianceBuilder(Lx, Ly, th, sd, lats, lons, v=float(g["nu"]), delta_x_method="Met_Office",
                                 precision=np.float64)
    E = synth.hadcrut_style_error_cov(idx, lons.size)
    yield g, b.cov_ns_device, idx, obs, E
    del b
    torch.cuda.empty_cache()


def _check(g, field, unc, mask):
    cells = g["cells"]
    assert max_norm_rel(field[cells], g["field"]) <= TOL
    assert max_norm_rel(unc[cells] ** 2, g["uncertainty"] ** 2) <= TOL
    assert max_norm_rel(mask[cells], g["constraint_mask"]) <= TOL


def test_ordinary_kriging_class_vs_reference_golden(case):
    from glomargridding_b200 import _dense as dn
    from glomargridding_b200.kriging import OrdinaryKriging

    g, C, idx, obs, E = case
    k = OrdinaryKriging(C, idx, obs, E)
    field, unc, mask = k.solve(), k.get_uncertainty(), k.constraint_mask()
    assert k._batch.lu_fallback == []                      # the DMMA Cholesky path, no LU fallback
    assert field.shape == (C.shape[0],) and np.isfinite(field).all() and np.isfinite(unc).all()
    _check(g, field, unc, mask)
    alpha = k._batch.out[0, dn.OUT_ALPHA].cpu().numpy()
    assert max_norm_rel(alpha[g["cells"]], g["alpha"]) <= 1e-7       # Lagrange multipliers (last weights column)


def test_krige_sharded_one_rank_vs_reference_golden(case):
    from glomargridding_b200.distributed import krige_sharded

    g, C, idx, obs, E = case
    r = krige_sharded(C, 0, idx, obs, E, "ordinary")
    _check(g, r["field"], r["uncertainty"], r["constraint_mask"])


def test_emulated_8_rank_factorisation_at_size(case):
    """The block-cyclic factorisation steps (`gmg_potrf_step_f64`, tile-row ownership t % 8) on the 20 000 x 20 000
    system of this configuration, ranks applied one after the other: bit-equal to the one-rank factorisation."""
    import torch

    from glomargridding_b200 import _dense as dn
    from glomargridding_b200 import _device as dev
    from glomargridding_b200._lib import check, lib

    g, C, idx, obs, E = case
    n = idx.size
    idx_t = torch.from_numpy(idx).cuda()
    S0 = dn.padded(C[idx_t][:, idx_t] + torch.from_numpy(E).cuda())
    ref = S0.clone()
    ws_ref = dn.factor_inplace(ref, n)
    S = S0
    lin, pdb, ldp = dn.solve_ws(n)
    Linv = torch.zeros(lin, dtype=torch.float64, device="cuda")
    P = torch.empty(pdb, dtype=torch.float64, device="cuda")
    info = torch.zeros(1, dtype=torch.int32, device="cuda")
    d = np.zeros(1, dtype=dn.SOLVE_DESC)
    d["S"], d["lds"], d["Linv"], d["P"], d["ldp"], d["info"], d["n"] = (
        S.data_ptr(), S.stride(0), Linv.data_ptr(), P.data_ptr(), ldp, info.data_ptr(), n)
    dt = dn._desc_tensor(d, S.device)
    L, s = lib(), dev.stream_ptr()
    world = 8
    for k in range((n + dn.NB - 1) // dn.NB):
        check(L.gmg_potrf_step_f64(dt.data_ptr(), 1, n, k, 0, 1, 0, s), "diag")
        for phase in (1, 4, 3):                             # panel, rest of the update, next panel column (look-ahead)
            for r in range(world):
                check(L.gmg_potrf_step_f64(dt.data_ptr(), 1, n, k, phase, world, r, s), "step")
    torch.cuda.synchronize()
    assert int(info.item()) == 0
    assert torch.equal(Linv, ws_ref["Linv"])
    for r0 in range(0, n, 4096):                            # lower triangle, in row slabs
        a, b = S[r0:r0 + 4096, :n], ref[r0:r0 + 4096, :n]
        keep = torch.arange(n, device="cuda")[None, :] <= torch.arange(r0, min(n, r0 + 4096), device="cuda")[:, None]
        assert torch.equal(a[keep], b[keep])
