"""The multi-GPU code paths on ONE GPU: every kernel and host routine that the row-block / block-cyclic sharding goes
through, with the ranks emulated sequentially in one process (the collectives of those paths only move data that a
single process already holds), so the driver's one-GPU `-m gpu` run covers them.

* `krige_sharded` at world = 1 against the oracle (5 degrees) - the routine that BASELINE configs[3] runs on 8 GPUs.
* `gmg_potrf_step_f64` (diagonal / panel / update phases with tile-row ownership `tr_mod`, `tr_rem`): the ranks of a
  1-D block-cyclic factorisation applied one after the other to one matrix must reproduce `gmg_potrf_batched_f64`
  bit for bit.
* `gmg_ellipse_cov_p2p_f64`: every rank's launch of the pair-once build, writing into "peer" row blocks that here
  live on the same device, must reproduce the single-rank matrix bit for bit.
* the distributed-factorisation driver (`_dense.factor_inplace_distributed`) on a one-rank group.
"""

import ctypes
import warnings

import numpy as np
import pytest
from conftest import max_norm_rel

from glomargridding_b200 import synthetic as synth

pytestmark = pytest.mark.gpu
TOL = 1e-8


@pytest.fixture(scope="module")
def cuda():
    import torch

    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch


def _cov5(torch, method="Met_Office", v=1.5):
    from glomargridding_b200.ellipse import EllipseCovarianceBuilder

    lats, lons = synth.global_grid(5.0)
    Lx, Ly, th, sd = synth.ellipse_fields(lats, lons)
    b = EllipseCovarianceBuilder(Lx, Ly, th, sd, lats, lons, v=v, delta_x_method=method, precision=np.float64)
    return b, lons.size


@pytest.mark.parametrize("method", ["ordinary", "simple"])
def test_krige_sharded_world1_vs_oracle(cuda, method):
    from glomargridding_b200.distributed import krige_sharded
    from oracle import ordinary_kriging, simple_kriging

    b, nlon = _cov5(cuda)
    C = b.cov_ns_device
    idx, obs = synth.fixed_observations(5, C.shape[0], 900)
    E = synth.hadcrut_style_error_cov(idx, nlon)
    got = krige_sharded(C, 0, idx, obs, E, method, mean=0.25)
    cov = b.cov_ns
    want = ordinary_kriging(cov, idx, obs, E) if method == "ordinary" else simple_kriging(cov, idx, obs, E, mean=0.25)
    assert max_norm_rel(got["field"], want["field"]) <= TOL
    assert max_norm_rel(got["uncertainty"] ** 2, want["uncertainty"] ** 2) <= TOL
    assert max_norm_rel(got["constraint_mask"], want["constraint_mask"]) <= TOL
    # no error covariance at all (the reference's error_cov=None)
    got0 = krige_sharded(C, 0, idx, obs, None, method)
    want0 = ordinary_kriging(cov, idx, obs, None) if method == "ordinary" else simple_kriging(cov, idx, obs, None)
    assert max_norm_rel(got0["field"], want0["field"]) <= 1e-6       # S = C[idx, idx] alone is poorly conditioned


def test_krige_sharded_rejects_float32_rows_and_unsorted_idx(cuda):
    from glomargridding_b200.distributed import krige_sharded

    b, nlon = _cov5(cuda)
    idx, obs = synth.fixed_observations(5, 2592, 100)
    with pytest.raises(ValueError, match="float64"):
        krige_sharded(b.cov_ns_device.float(), 0, idx, obs, None)
    with pytest.raises(ValueError, match="sorted"):
        krige_sharded(b.cov_ns_device, 0, idx[::-1].copy(), obs, None)


def _spd(torch, n, seed):
    g = torch.Generator(device="cuda").manual_seed(seed)
    A = torch.randn((n, n), dtype=torch.float64, device="cuda", generator=g)
    return A @ A.T / n + torch.eye(n, dtype=torch.float64, device="cuda")


@pytest.mark.parametrize("n", [1000, 1152])
@pytest.mark.parametrize("world", [2, 4, 8])
def test_potrf_step_emulated_ranks_bit_equal(cuda, n, world):
    """Rank r of `world` owns the tile rows t with t % world == r.  Per block step: the owner's diagonal phase, then
    every rank's panel phase, then every rank's update phase - the order the distributed driver enforces with its
    exchanges - here on one matrix, one rank after the other."""
    torch = cuda
    from glomargridding_b200 import _dense as dn
    from glomargridding_b200 import _device as dev
    from glomargridding_b200._lib import check, lib

    S0 = dn.padded(_spd(torch, n, n + world))
    L, s = lib(), dev.stream_ptr()
    nblk = (n + dn.NB - 1) // dn.NB

    def run(ranks, phases):
        S = S0.clone()
        lin, pdb, ldp = dn.solve_ws(n)
        Linv = torch.zeros(lin, dtype=torch.float64, device="cuda")
        P = torch.empty(pdb, dtype=torch.float64, device="cuda")
        info = torch.zeros(1, dtype=torch.int32, device="cuda")
        d = np.zeros(1, dtype=dn.SOLVE_DESC)
        d["S"], d["lds"], d["Linv"], d["P"], d["ldp"], d["info"], d["n"] = (
            S.data_ptr(), S.stride(0), Linv.data_ptr(), P.data_ptr(), ldp, info.data_ptr(), n)
        dt = dn._desc_tensor(d, S.device)
        for k in range(nblk):
            check(L.gmg_potrf_step_f64(dt.data_ptr(), 1, n, k, 0, 1, 0, s), "diag")
            for phase in phases:
                for r in range(ranks):
                    check(L.gmg_potrf_step_f64(dt.data_ptr(), 1, n, k, phase, ranks, r, s), "step")
        torch.cuda.synchronize()
        assert int(info.item()) == 0
        return S, Linv

    # the one-rank right-looking factorisation (systems this small are factored LEFT-looking by gmg_potrf_solve_f64 /
    # gmg_potrf_batched_f64 on one GPU - same factor to rounding; the multi-GPU driver is right-looking at any size)
    ref, Linv_ref = run(1, (1, 2))
    # the look-ahead split of the overlapped driver: phase 3 (next panel column) + phase 4 (the rest) == phase 2
    S, Linv = run(world, (1, 4, 3) if world == 4 else (1, 2))
    tri = torch.tril(torch.ones((n, n), dtype=torch.bool, device="cuda"))
    assert torch.equal(S[:n, :n][tri], ref[:n, :n][tri])
    assert torch.equal(Linv, Linv_ref)
    left = S0.clone()
    dn.factor_inplace(left, n)                                  # left-looking on one GPU: equal to rounding
    assert float((left[:n, :n][tri] - ref[:n, :n][tri]).abs().max() / ref[:n, :n].abs().max()) <= 1e-13
    # and it is a Cholesky factor
    Lf = torch.tril(S[:n, :n])
    assert float((Lf @ Lf.T - S0[:n, :n]).abs().max() / S0[:n, :n].abs().max()) <= 1e-13


@pytest.mark.parametrize("sizes,nrhs", [([900], 300), ([128], 2), ([1000, 257, 640], 70), ([1341] * 5, 2594)])
def test_overlapped_factor_solve_bit_equal_to_sequential(cuda, sizes, nrhs):
    """`gmg_potrf_solve_f64` (three streams: critical path / bulk update / forward-solve steps) against
    `gmg_potrf_batched_f64` + `gmg_potrs_batched_f64(FORWARD)`: same factor, same solved right-hand sides, bit for bit -
    for one system, a ragged batch and a batch of month-sized systems."""
    torch = cuda
    from glomargridding_b200 import _dense as dn
    from glomargridding_b200 import _device as dev
    from glomargridding_b200._lib import check, lib

    L, s = lib(), dev.stream_ptr()
    B, n_max = len(sizes), max(sizes)

    def system(seed):
        mats, rhs = [], []
        for b, n in enumerate(sizes):
            mats.append(dn.padded(_spd(torch, n, seed + b)))
            g = torch.Generator(device="cuda").manual_seed(100 + b)
            rhs.append(dn.padded(torch.randn((n, nrhs), dtype=torch.float64, device="cuda", generator=g)))
        ws = [dn.solve_ws(n) for n in sizes]
        Linv = [torch.zeros(w[0], dtype=torch.float64, device="cuda") for w in ws]
        P = [torch.empty(w[1], dtype=torch.float64, device="cuda") for w in ws]
        info = torch.zeros(B, dtype=torch.int32, device="cuda")
        d = np.zeros(B, dtype=dn.SOLVE_DESC)
        for b, n in enumerate(sizes):
            d[b] = (mats[b].data_ptr(), mats[b].stride(0), Linv[b].data_ptr(), rhs[b].data_ptr(), rhs[b].stride(0),
                    P[b].data_ptr(), ws[b][2], info.data_ptr() + 4 * b, n, nrhs)
        return mats, rhs, Linv, P, info, dn._desc_tensor(d, "cuda")

    m0, r0_, li0, _p0, info0, d0 = system(7)
    check(L.gmg_potrf_batched_f64(d0.data_ptr(), B, n_max, s), "potrf")
    check(L.gmg_potrs_batched_f64(d0.data_ptr(), B, n_max, nrhs, 1, s), "potrs")
    m1, r1_, li1, _p1, info1, d1 = system(7)
    for _ in range(2):                                   # twice: the event pool / streams of the context are reused
        for a, b_ in zip(m1 + r1_, system(7)[0] + system(7)[1]):
            a.copy_(b_)
        check(L.gmg_potrf_solve_f64(None, d1.data_ptr(), B, n_max, nrhs, None, 0, s), "potrf_solve")
    torch.cuda.synchronize()
    assert int(info0.abs().sum()) == 0 and int(info1.abs().sum()) == 0
    for b, n in enumerate(sizes):
        tri = torch.tril(torch.ones((n, n), dtype=torch.bool, device="cuda"))
        assert torch.equal(m1[b][:n, :n][tri], m0[b][:n, :n][tri])
        assert torch.equal(li1[b], li0[b])
        assert torch.equal(r1_[b][:, :nrhs], r0_[b][:, :nrhs])
    # factor only
    m2, _r2, li2, _p2, info2, d2 = system(7)
    check(L.gmg_potrf_solve_f64(None, d2.data_ptr(), B, n_max, 0, None, 0, s), "potrf_solve")
    torch.cuda.synchronize()
    for b, n in enumerate(sizes):
        tri = torch.tril(torch.ones((n, n), dtype=torch.bool, device="cuda"))
        assert torch.equal(m2[b][:n, :n][tri], m0[b][:n, :n][tri])
    # a non-positive pivot is reported through info like gmg_potrf_batched_f64 reports it
    m3, _r3, _l3, _p3, info3, d3 = system(7)
    m3[0][5, 5] = -1.0
    check(L.gmg_potrf_solve_f64(None, d3.data_ptr(), B, n_max, nrhs, None, 0, s), "potrf_solve")
    torch.cuda.synchronize()
    assert int(info3[0]) == 6 and int(info3[1:].abs().sum()) == 0


def test_solve_strip_widths_bit_identical(cuda):
    """The forward solve picks its column-strip width from {128, 64, 56} by how the strips fill the SMs (8 104 columns -
    one 1-degree rank of 8 - take 56 on the 16 x 1 warp layout, 64 802 take 64, a month's 2 594 take the two-CTA 64):
    every column must come out the same whichever strip it rode in."""
    torch = cuda
    from glomargridding_b200 import _dense as dn
    from glomargridding_b200 import _device as dev
    from glomargridding_b200._lib import check, lib

    L, s = lib(), dev.stream_ptr()
    n = 2500                                                  # K beyond the two-CTA limit in the later block steps
    S = dn.padded(_spd(torch, n, 3))
    ws = dn.factor_inplace(S, n)
    g = torch.Generator(device="cuda").manual_seed(9)
    cols = {w: c for w, c in (("rank of 8", 8104), ("whole grid", 64802), ("month", 2594))}
    base = torch.randn((n, 8104), dtype=torch.float64, device="cuda", generator=g)
    out = {}
    for name, nrhs in cols.items():
        R = torch.zeros((n, nrhs + (nrhs & 1)), dtype=torch.float64, device="cuda")
        k = min(nrhs, 8104)
        R[:, :k] = base[:, :k]
        d = np.zeros(1, dtype=dn.SOLVE_DESC)
        d[0] = (S.data_ptr(), S.stride(0), ws["Linv"].data_ptr(), R.data_ptr(), R.stride(0), ws["P"].data_ptr(),
                ws["ldp"], ws["info"].data_ptr(), n, nrhs)
        dt = dn._desc_tensor(d, "cuda")
        check(L.gmg_potrs_batched_f64(dt.data_ptr(), 1, n, nrhs, 1, s), "potrs")
        torch.cuda.synchronize()
        out[name] = R[:, :k].clone()
    assert torch.equal(out["rank of 8"], out["whole grid"])
    assert torch.equal(out["rank of 8"][:, :2594], out["month"])
    Lf = torch.tril(S[:n, :n])
    want = torch.linalg.solve_triangular(Lf, base, upper=False)
    assert float((out["rank of 8"] - want).abs().max() / want.abs().max()) <= 1e-11


def test_dist_api_argument_checks(cuda):
    from glomargridding_b200._lib import lib

    L = lib()
    need = ctypes.c_int64()
    assert L.gmg_dist_ws_doubles(20000, 8, ctypes.byref(need)) == 0
    assert need.value == 128 * 128 * (2 + 20 * 9)        # [L_kk | Linv_k] + (1 + world) x ceil(156 / 8) packed tiles
    assert L.gmg_dist_ws_doubles(20000, 1, ctypes.byref(need)) == 0 and need.value == 0
    assert L.gmg_dist_ws_doubles(0, 2, ctypes.byref(need)) == -1
    h = ctypes.c_void_p()
    assert L.gmg_dist_init(None, 2, 0, ctypes.byref(h)) == -1          # a multi-rank context needs the unique id
    assert L.gmg_dist_init(None, 1, 1, ctypes.byref(h)) == -3
    assert L.gmg_dist_init(None, 1, 0, ctypes.byref(h)) == 0 and h.value  # one rank: streams only, no communicator
    assert L.gmg_potrf_solve_f64(h, None, 1, 128, 0, None, 0, None) == -2
    assert L.gmg_dist_destroy(h) == 0
    uid = ctypes.create_string_buffer(128)
    assert L.gmg_dist_unique_id(uid) == 0 and any(uid.raw)              # NCCL is reachable in this process


def test_potrf_step_argument_checks(cuda):
    from glomargridding_b200._lib import lib

    L = lib()
    assert L.gmg_potrf_step_f64(None, 1, 256, 0, 0, 1, 0, None) == -1
    assert L.gmg_potrf_step_f64(ctypes.c_void_p(8), 1, 256, 2, 0, 1, 0, None) == -4      # block step past the end
    assert L.gmg_potrf_step_f64(ctypes.c_void_p(8), 1, 256, 0, 5, 1, 0, None) == -5
    assert L.gmg_potrf_step_f64(ctypes.c_void_p(8), 1, 256, 0, 1, 2, 2, None) == -6


@pytest.mark.parametrize("world", [2, 3, 8])
@pytest.mark.parametrize("v", [1.5, 2.2])
def test_ellipse_cov_p2p_emulated_ranks_bit_equal(cuda, world, v):
    """Every unordered pair of row blocks is computed by exactly one rank, which also stores the transposed tile
    into the other block; after all `world` launches the blocks hold the same matrix as the one-rank build."""
    torch = cuda
    from glomargridding_b200 import _device as dev
    from glomargridding_b200._lib import check, lib
    from glomargridding_b200.ellipse.p2p import tile_aligned_blocks

    b, _ = _cov5(torch, "Modified_Met_Office", v)
    N = b.cov_ns_device.shape[0]
    blocks = tile_aligned_blocks(N, world)
    parts = [torch.full((int(blocks[g + 1] - blocks[g]), N), float("nan"), dtype=torch.float64, device="cuda")
             for g in range(world)]
    ptrs = (ctypes.c_void_p * world)(*[p.data_ptr() for p in parts])
    row0 = (ctypes.c_int64 * (world + 1))(*[int(x) for x in blocks])
    for r in range(world):
        check(lib().gmg_ellipse_cov_p2p_f64(b._ws.data_ptr(), N, float(v), -1.0, 0, world, r, row0, ptrs, N, 0,
                                            dev.stream_ptr()), "p2p")
    torch.cuda.synchronize()
    got = torch.cat(parts, dim=0)
    assert not torch.isnan(got).any()                       # every entry was written by somebody
    assert torch.equal(got, b.cov_ns_device)


def test_factor_inplace_distributed_single_rank_group(cuda, tmp_path):
    """The distributed driver itself (its exchanges degenerate on a one-rank group): same factor as the local one."""
    torch = cuda
    import torch.distributed as dist

    from glomargridding_b200 import _dense as dn

    if dist.is_initialized():
        pytest.skip("a process group already exists in this process")
    dist.init_process_group("nccl", init_method=f"file://{tmp_path}/rdzv", rank=0, world_size=1,
                            device_id=torch.device("cuda", torch.cuda.current_device()))
    try:
        n = 900
        S0 = dn.padded(_spd(torch, n, 17))
        a, bb = S0.clone(), S0.clone()
        ws_a = dn.factor_inplace(a, n)
        ws_b = dn.factor_inplace_distributed(bb, n)
        tri = torch.tril(torch.ones((n, n), dtype=torch.bool, device="cuda"))
        assert torch.equal(a[:n, :n][tri], bb[:n, :n][tri]) and torch.equal(ws_a["Linv"], ws_b["Linv"])
        with pytest.raises(np.linalg.LinAlgError):
            bad = S0.clone()
            bad[5, 5] = -1.0
            dn.factor_inplace_distributed(bad, n)
    finally:
        dist.destroy_process_group()


def test_obs_replaced_after_factorisation(cuda):
    """`k.obs = other; k.solve()` re-evaluates the field with the kept weights (reference kriging.py:417, :677-679:
    `kriging_weights @ self.obs` at call time) - for the Cholesky path, the LU fallback and the stochastic class."""
    from glomargridding_b200.kriging import OrdinaryKriging, SimpleKriging
    from glomargridding_b200.stochastic import StochasticKriging

    b, nlon = _cov5(cuda)
    C = b.cov_ns_device
    idx, obs = synth.fixed_observations(9, C.shape[0], 700)
    obs2 = np.random.default_rng(3).standard_normal(obs.size)
    E = synth.hadcrut_style_error_cov(idx, nlon)
    for cls in (SimpleKriging, OrdinaryKriging):
        k = cls(C, idx, obs.copy(), E.copy())
        z1 = k.solve()
        fresh = cls(C, idx, obs2.copy(), E.copy()).solve()
        k.obs = obs2.copy()
        assert max_norm_rel(k.solve(), fresh) <= 1e-11
        k.obs = obs.copy()
        assert np.array_equal(k.solve(), z1)
    # LU fallback: an indefinite covariance (default displacement method)
    bi, _ = _cov5(cuda, "Modified_Met_Office")
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        k = OrdinaryKriging(bi.cov_ns_device, idx, obs.copy(), E.copy())
        k.solve()
        if k._batch.lu_fallback:
            fresh = OrdinaryKriging(bi.cov_ns_device, idx, obs2.copy(), E.copy()).solve()
            k.obs = obs2.copy()
            assert max_norm_rel(k.solve(), fresh) <= 1e-9
    sk = StochasticKriging(C, idx, obs.copy(), E.copy())
    state = np.zeros(C.shape[0])
    np.random.seed(1)
    sk.solve(simulated_state=state)
    g1 = sk.gridded_field.copy()
    sk.obs = obs2.copy()
    np.random.seed(1)
    sk.solve(simulated_state=state)
    assert max_norm_rel(sk.gridded_field, OrdinaryKriging(C, idx, obs2.copy(), E.copy()).solve()) <= 1e-11
    assert not np.allclose(g1, sk.gridded_field)


def test_observation_cells_validated_like_numpy(cuda):
    from glomargridding_b200.kriging import SimpleKriging

    b, nlon = _cov5(cuda)
    C = b.cov_ns_device
    M = C.shape[0]
    idx, obs = synth.fixed_observations(2, M, 50)
    for bad in (M, M + 7, -M - 1):
        i2 = idx.copy()
        i2[3] = bad
        with pytest.raises(IndexError, match="out of bounds"):
            SimpleKriging(C, i2, obs, None).solve()
    neg = idx.copy()
    neg[-1] = idx[-1] - M                               # numpy wraps negative cells
    assert np.array_equal(SimpleKriging(C, neg, obs, None).solve(), SimpleKriging(C, idx, obs, None).solve())


def test_asymmetric_error_cov_in_few_entries_is_detected(cuda):
    """An error covariance that is asymmetric in a handful of entries passes the sampled pre-filter; the exact test
    inside the packing tasks must still route it to the LU solve of the matrix AS GIVEN (reference kriging.py:342-343,
    :576-585), and `symmetry="assume"` must read the lower triangle only."""
    from glomargridding_b200 import _dense as dn
    from oracle import ordinary_kriging

    b, nlon = _cov5(cuda)
    C = b.cov_ns_device
    idx, obs = synth.fixed_observations(11, C.shape[0], 640)
    E = synth.hadcrut_style_error_cov(idx, nlon)
    Ea = E.copy()
    Ea[17, 400] += 0.05                                 # one entry of the upper triangle
    Ea[3, 630] -= 0.02
    assert dn._looks_symmetric(Ea) and not dn.is_symmetric(Ea) and dn.is_symmetric(E)
    prob = [(idx, obs, Ea)]
    with pytest.warns(UserWarning, match="not symmetric"):
        kb = dn.KrigeBatch(C, prob, ordinary=True).run()
    assert kb.lu_fallback == [0]
    want = ordinary_kriging(b.cov_ns, idx, obs, Ea)
    assert max_norm_rel(kb.out[0, dn.OUT_FIELD].cpu().numpy(), want["field"]) <= TOL
    # several problems, one of them asymmetric: only that one takes the LU route
    probs = [(idx, obs, E), (idx, obs, Ea), (idx, obs * 2, E)]
    with pytest.warns(UserWarning, match="not symmetric"):
        kb3 = dn.KrigeBatch(C, probs, ordinary=True).run()
    with pytest.warns(UserWarning, match="not symmetric"):              # large batches: "exact" on request
        kb9 = dn.KrigeBatch(C, probs * 3, ordinary=True, symmetry="exact").run()
    assert kb9.lu_fallback == [1, 4, 7]
    assert dn.KrigeBatch(C, probs * 3, ordinary=True).run().lu_fallback == []      # "sample" does not see two entries
    assert kb3.lu_fallback == [1]
    assert max_norm_rel(kb3.out[1, dn.OUT_FIELD].cpu().numpy(), want["field"]) <= TOL
    want_sym = ordinary_kriging(b.cov_ns, idx, obs, E)
    assert max_norm_rel(kb3.out[0, dn.OUT_FIELD].cpu().numpy(), want_sym["field"]) <= TOL
    # the caller vouches: upper triangle ignored -> the answer of the symmetric matrix
    kbs = dn.KrigeBatch(C, prob, ordinary=True, symmetry="assume").run()
    assert kbs.lu_fallback == []
    assert max_norm_rel(kbs.out[0, dn.OUT_FIELD].cpu().numpy(), want_sym["field"]) <= TOL


def test_krige_sharded_indefinite_covariance_takes_the_lu_like_the_reference(cuda):
    """The DEFAULT displacement method's covariance is indefinite: the reference solves it by LU all the same
    (np.linalg.solve); `krige_sharded` warns, re-assembles the system and goes through the blocked LU - one rank here,
    every rank alike in a group - and agrees with the oracle's (= the reference's) formulation."""
    import warnings

    from glomargridding_b200.distributed import krige_sharded
    from glomargridding_b200.ellipse import EllipseCovarianceBuilder
    from oracle import ordinary_kriging, simple_kriging

    lats, lons = synth.global_grid(5.0)
    Lx, Ly, th, sd = synth.ellipse_fields(lats, lons)
    b = EllipseCovarianceBuilder(Lx, Ly, th, sd, lats, lons, v=1.5, precision=np.float64)     # Modified_Met_Office
    C, cov = b.cov_ns_device, b.cov_ns
    idx, obs = synth.fixed_observations(7, C.shape[0], 700)
    E = synth.hadcrut_style_error_cov(idx, lons.size)
    for method, oracle in (("ordinary", ordinary_kriging), ("simple", simple_kriging)):
        with warnings.catch_warnings(record=True) as w:
            warnings.simplefilter("always")
            got = krige_sharded(C, 0, idx, obs, E, method)
            want = oracle(cov, idx, obs, E)
        assert any("LU with partial pivoting" in str(x.message) for x in w)
        assert max_norm_rel(got["field"], want["field"]) <= 1e-6
        assert max_norm_rel(got["constraint_mask"], want["constraint_mask"]) <= 1e-6
