"""Parity of the CUDA kriging path (SimpleKriging / OrdinaryKriging / StochasticKriging and the dense FP64
kernels under them) against the oracle, the reference-generated goldens and the reference's own GeoStats.jl
known-answer vector.  Tolerance: 1e-8 max-norm relative on fields, uncertainties, masks and weights
(BASELINE.json north_star; SURVEY.md §8d explains why max-norm).  Everything goes through the Python classes,
i.e. through the C ABI of libglomar_b200.so.  The structure mirrors reference test/test_kriging.py.
"""

import warnings

import numpy as np
import pytest
from conftest import load_golden, max_norm_rel

from glomargridding_b200 import synthetic as synth

pytestmark = pytest.mark.gpu
TOL = 1e-8


def var_rel(u_got, u_want):
    """Max-norm relative difference of the kriging VARIANCES (north_star: "1e-8 relative on kriged fields and
    kriging variances").  The uncertainty is their square root, which turns the ~1e-16 rounding noise of an
    exactly-zero variance (an error-free observation) into ~1e-8 - not a meaningful difference."""
    return max_norm_rel(np.asarray(u_got, dtype=np.float64) ** 2, np.asarray(u_want, dtype=np.float64) ** 2)


@pytest.fixture(scope="module")
def K():
    import torch

    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import glomargridding_b200.kriging as k
    import glomargridding_b200.stochastic as s

    class NS:
        pass

    ns = NS()
    for mod in (k, s):
        for name in dir(mod):
            if not name.startswith("__"):
                setattr(ns, name, getattr(mod, name))
    return ns


def quiet(fn, *a, **kw):
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        return fn(*a, **kw)


# ---- dense kernels ---------------------------------------------------------------------------------------
@pytest.mark.parametrize("ta,tb", [(False, False), (False, True), (True, False), (True, True)])
@pytest.mark.parametrize("m,n,k", [(1, 1, 1), (5, 3, 2), (128, 128, 16), (129, 257, 33), (300, 70, 515)])
def test_gemm_vs_numpy(K, ta, tb, m, n, k):
    import torch

    from glomargridding_b200 import _dense as dn

    rng = np.random.default_rng(m * 1000 + n * 10 + k)
    A = rng.standard_normal((k, m) if ta else (m, k))
    B = rng.standard_normal((n, k) if tb else (k, n))
    C0 = rng.standard_normal((m, n))
    At, Bt = dn.padded(A), dn.padded(B)
    out = dn.padded(C0)
    dn.gemm(At[:, : A.shape[1]], Bt[:, : B.shape[1]], ta, tb, out=out[:, :n], alpha=0.75, beta=-0.5)
    want = 0.75 * (A.T if ta else A) @ (B.T if tb else B) - 0.5 * C0
    got = out[:, :n].cpu().numpy()
    assert max_norm_rel(got, want) <= 1e-13
    _ = torch


@pytest.mark.parametrize("n,nrhs", [(1, 1), (5, 3), (127, 40), (128, 128), (129, 7), (300, 515), (700, 1000)])
def test_cholesky_solve_vs_numpy(K, n, nrhs):
    from glomargridding_b200 import _dense as dn

    rng = np.random.default_rng(n)
    A = rng.standard_normal((n, n + 5))
    S = A @ A.T + 0.1 * np.eye(n)
    R = rng.standard_normal((n, nrhs))
    St, Rt = dn.padded(S), dn.padded(R)
    dn.cholesky_solve_inplace(St, Rt, lu_fallback=False)
    want = np.linalg.solve(S, R)
    assert max_norm_rel(Rt[:, :nrhs].cpu().numpy(), want) <= 1e-10
    L = np.tril(St[:, :n].cpu().numpy())
    assert max_norm_rel(L @ L.T, S) <= 1e-13            # factor check, SURVEY.md §8(d)


def test_cholesky_reports_indefinite_and_lu_fallback_matches_numpy(K):
    from glomargridding_b200 import _dense as dn

    rng = np.random.default_rng(3)
    n = 150
    A = rng.standard_normal((n, n))
    S = A + A.T                                          # symmetric indefinite
    R = rng.standard_normal((n, 37))
    with pytest.raises(np.linalg.LinAlgError, match="not positive definite"):
        dn.cholesky_solve_inplace(dn.padded(S), dn.padded(R), lu_fallback=False)
    Rt = dn.padded(R)
    with pytest.warns(UserWarning, match="LU with partial pivoting"):
        dn.cholesky_solve_inplace(dn.padded(S), Rt)
    assert max_norm_rel(Rt[:, :37].cpu().numpy(), np.linalg.solve(S, R)) <= 1e-9
    with pytest.raises(np.linalg.LinAlgError, match="Singular"):
        quiet(dn.cholesky_solve_inplace, dn.padded(np.zeros((4, 4))), dn.padded(np.ones((4, 2))))


@pytest.mark.parametrize("sizes,nrhs", [([1], 1), ([5], 3), ([128], 64), ([129], 200), ([300, 77, 515], 70), ([1341], 2594),
                                        ([3000], 5), ([900, 700, 1100], 9), ([40] * 400, 2)])
def test_blocked_lu_vs_numpy(K, sizes, nrhs):
    """`gmg_gesv_blocked_f64` (panel factorisation with partial pivoting, GEMM updates, GEMM substitutions) against
    numpy.linalg.solve (LAPACK dgesv - what the reference calls, kriging.py:343, :585): non-symmetric matrices, a zero
    diagonal (pivoting is not optional), ragged batches, the pivots LAPACK would choose."""
    import torch
    import scipy.linalg as sla

    from glomargridding_b200 import _dense as dn
    from glomargridding_b200 import _device as dev
    from glomargridding_b200._lib import check, lib

    rng = np.random.default_rng(len(sizes) * 1000 + sizes[0])
    B, n_max = len(sizes), max(sizes)
    lin, pdb, ldp = dn.gesv_ws(n_max)
    mats, rhss, want, keep = [], [], [], []
    info = torch.zeros(B, dtype=torch.int32, device="cuda")
    d = np.zeros(B, dtype=dn.SOLVE_DESC)
    for b, n in enumerate(sizes):
        A = rng.standard_normal((n, n))
        if n > 1:
            A[np.arange(n), np.arange(n)] = 0.0             # forces row interchanges
        R = rng.standard_normal((n, nrhs))
        mats.append(A)
        want.append(np.linalg.solve(A, R))
        St, Rt = dn.padded(A), dn.padded(R)
        li = torch.empty(lin, dtype=torch.float64, device="cuda")
        st = torch.empty(pdb, dtype=torch.float64, device="cuda")
        keep += [St, Rt, li, st]
        rhss.append(Rt)
        d[b] = (St.data_ptr(), St.stride(0), li.data_ptr(), Rt.data_ptr(), Rt.stride(0), st.data_ptr(), ldp,
                info.data_ptr() + 4 * b, n, nrhs)
    dt = dn._desc_tensor(d, "cuda")
    check(lib().gmg_gesv_blocked_f64(dt.data_ptr(), B, n_max, nrhs, dev.stream_ptr()), "gesv_blocked")
    torch.cuda.synchronize()
    assert int(info.abs().sum()) == 0
    for b, n in enumerate(sizes):
        assert max_norm_rel(rhss[b][:, :nrhs].cpu().numpy(), want[b]) <= 1e-9, (b, n)
        # the pivots are LAPACK's (partial pivoting, first maximal entry)
        piv = keep[4 * b + 3][int(lib().gmg_gesv_pivots_offset(ldp)):].view(torch.int32)[:n].cpu().numpy()
        assert np.array_equal(piv, sla.lu_factor(mats[b])[1])
    # an exactly singular system reports the zero pivot through info, like dgesv
    Z = dn.padded(np.zeros((6, 6)))
    R = dn.padded(np.ones((6, 2)))
    li, st = torch.empty(lin, dtype=torch.float64, device="cuda"), torch.empty(pdb, dtype=torch.float64, device="cuda")
    inf1 = torch.zeros(1, dtype=torch.int32, device="cuda")
    d1 = np.zeros(1, dtype=dn.SOLVE_DESC)
    d1[0] = (Z.data_ptr(), Z.stride(0), li.data_ptr(), R.data_ptr(), R.stride(0), st.data_ptr(), ldp, inf1.data_ptr(), 6, 2)
    check(lib().gmg_gesv_blocked_f64(dn._desc_tensor(d1, "cuda").data_ptr(), 1, 6, 2, dev.stream_ptr()), "gesv_blocked")
    assert int(inf1.item()) == 1


def test_indefinite_month_takes_the_blocked_lu_and_matches_the_oracle(K):
    """The reference's DEFAULT displacement method gives an indefinite covariance: a month kriged against
    it without clipping falls to LU with partial pivoting - now blocked on DMMA tiles - and agrees with the oracle's
    (= the reference's) np.linalg.solve formulation."""
    from glomargridding_b200.ellipse import EllipseCovarianceBuilder
    from oracle import ordinary_kriging

    lats, lons = synth.global_grid(5.0)
    Lx, Ly, th, sd = synth.ellipse_fields(lats, lons)
    b = EllipseCovarianceBuilder(Lx, Ly, th, sd, lats, lons, v=1.5, precision=np.float64)     # Modified_Met_Office
    cov = b.cov_ns
    idx, obs = synth.month_observations(3, cov.shape[0])
    E = synth.hadcrut_style_error_cov(idx, lons.size, seed=3)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        k = K.OrdinaryKriging(b.cov_ns_device, idx, obs, E.copy())
        z, u = k.solve(), k.get_uncertainty()
        want = ordinary_kriging(cov, idx, obs, E, big_temporaries=False)
    assert k._batch.lu_fallback                                 # (it is indefinite for these fields)
    assert max_norm_rel(z, want["field"]) <= 1e-6 and var_rel(u, want["uncertainty"]) <= 1e-6
    # several systems of different sizes in ONE LU batch (workspaces carved out of one allocation), each against the
    # oracle: three indefinite months next to ones Cholesky accepts (no error covariance: nothing lifts the spectrum)
    from glomargridding_b200.batched import krige_batch

    probs = []
    for m, n in enumerate((129, 301, 517, 77, 640)):
        i2, o2 = synth.month_observations(10 + m, cov.shape[0], n_min=n, n_max=n + 1)
        probs.append((i2, o2, None if m % 2 else synth.hadcrut_style_error_cov(i2, lons.size, seed=10 + m)))
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        got = krige_batch(b.cov_ns_device, probs, "ordinary")
        for m, (i2, o2, E2) in enumerate(probs):
            want = ordinary_kriging(cov, i2, o2, E2, big_temporaries=False)
            assert max_norm_rel(got["field"][m], want["field"]) <= 1e-6, m
            assert max_norm_rel(got["constraint_mask"][m], want["constraint_mask"]) <= 1e-6, m


# ---- known-answer: GeoStats.jl (reference test/test_kriging.py:67-192) ------------------------------------
def test_geostats_known_answer_class_and_function(K):
    from oracle.variogram import geostats_setup

    g = load_golden("geostats_ok.npz")
    cov, idx, obs = geostats_setup()                     # a VARIOGRAM matrix passed as the covariance: indefinite
    ok = K.OrdinaryKriging(cov, idx=idx, obs=obs)
    with pytest.warns(UserWarning, match="not positive definite"):
        k = ok.solve()
    assert np.allclose(g["expected"], k.reshape(20, 20))                 # the reference's own criterion
    S, SS = cov[idx[:, None], idx[None, :]], cov[idx, :]
    k2, _ = quiet(K.kriging_ordinary, S, SS, obs, cov)
    assert np.allclose(g["expected"], k2.reshape(20, 20))
    # from pre-computed weights and from the extended inverse (reference :130-192)
    ok2 = K.OrdinaryKriging(cov, idx=idx, obs=obs)
    ok2.set_kriging_weights(ok.kriging_weights)
    assert np.allclose(g["expected"], ok2.solve().reshape(20, 20))
    S_ext = np.block([[S, np.ones((3, 1))], [np.ones((1, 3)), 0]])
    ok3 = K.OrdinaryKriging(cov, idx=idx, obs=obs)
    ok3.kriging_weights_from_inverse(np.linalg.inv(S_ext))
    assert np.allclose(g["expected"], ok3.solve().reshape(20, 20))


# ---- goldens generated by the unmodified reference (tests/golden/make_golden.py) ---------------------------
@pytest.mark.parametrize("tag", ["spd", "indef"])
def test_kriging_small_vs_reference_golden(K, tag):
    g = load_golden("kriging_small.npz")
    cov = g[f"{tag}_cov"]
    sk = quiet(K.SimpleKriging, cov, g["idx"].copy(), g["obs"].copy(), g["E_full"].copy())
    ok = quiet(K.OrdinaryKriging, cov, g["idx"].copy(), g["obs"].copy(), g["E_full"].copy())
    ok0 = K.OrdinaryKriging(cov, g["idx"].copy(), g["obs"].copy())
    assert np.array_equal(ok.idx, g["kept_idx"])
    got = {
        "sk_field": quiet(sk.solve, 0.25), "sk_unc": quiet(sk.get_uncertainty), "sk_mask": quiet(sk.constraint_mask),
        "sk_weights": sk.kriging_weights, "ok_field": quiet(ok.solve), "ok_unc": quiet(ok.get_uncertainty),
        "ok_mask": quiet(ok.constraint_mask), "ok_weights": ok.kriging_weights,
        "ok_noE_field": quiet(ok0.solve), "ok_noE_unc": quiet(ok0.get_uncertainty),
    }
    for key, val in got.items():
        want = g[f"{tag}_{key}"]
        assert val.shape == want.shape and val.dtype == want.dtype, key
        assert (var_rel if key.endswith("unc") else max_norm_rel)(val, want) <= TOL, key
    # the full "indef" covariance is indefinite, but its observed sub-block + E is still positive definite:
    # both cases take the Cholesky path (the LU fallback is exercised by the GeoStats.jl test)
    assert ok._batch.lu_fallback == []


def test_stochastic_injected_draws_vs_reference_golden(K):
    g = load_golden("kriging_small.npz")
    st = quiet(K.StochasticKriging, g["spd_cov"], g["idx"].copy(), g["obs"].copy(), g["E_full"].copy())
    st.get_kriging_weights()
    # inject the reference's draws: state through the API, observation noise through the module's draw hook
    import glomargridding_b200.stochastic as smod

    real = smod.scipy_mv_normal_draw
    smod.scipy_mv_normal_draw = lambda loc, cov, ndraws=1, **kw: g["st_noise"].copy()
    try:
        res = quiet(st.solve, simulated_state=g["st_state"].copy())
    finally:
        smod.scipy_mv_normal_draw = real
    for key, val in [("st_result", res), ("st_epsilon", st.epsilon), ("st_gridded_field", st.gridded_field),
                     ("st_simulated_grid", st.simulated_grid), ("st_unc", quiet(st.get_uncertainty)),
                     ("st_mask", st.constraint_mask()), ("st_ok_weights", st.kriging_weights),
                     ("st_sk_weights", st.simple_kriging_weights)]:
        assert val.shape == g[key].shape, key
        assert (var_rel if key.endswith("unc") else max_norm_rel)(val, g[key]) <= TOL, key


@pytest.mark.parametrize("n", [10, 25, 100])
def test_inverse_trick(K, n):
    """reference test/test_kriging.py:297-316"""
    g = load_golden("kriging_small.npz")
    S = g[f"ext_inv_{n}_S"]
    got = K._extended_inverse(np.linalg.inv(S))
    assert np.allclose(got, g[f"ext_inv_{n}"])
    S_ext = np.block([[S, np.ones((n, 1))], [np.ones((1, n)), 0]])
    assert np.allclose(got, np.linalg.inv(S_ext))
    with pytest.raises(ValueError, match="S must be a matrix"):
        K._extended_inverse(np.ones(4))


# ---- class methods against the functional API (reference test/test_kriging.py:195-372) ---------------------
def _class_method_inputs(seed, nan_diag=False):
    from oracle.variogram import geostats_setup, matern_variogram

    _, idx, obs = geostats_setup()
    lat, lon = np.meshgrid(np.arange(1.0, 21.0), np.arange(1.0, 21.0), indexing="ij")
    pts = np.column_stack([lat.ravel(), lon.ravel()])
    dist = np.sqrt(((pts[:, None, :] - pts[None, :, :]) ** 2).sum(-1))
    cov = 4.0 - matern_variogram(dist, psill=4.0, nugget=0.0, range_=35 / 3, nu=1.5)   # a true covariance
    rng = np.random.default_rng(seed)
    ev = rng.random((3, 3))
    ev = ev @ ev.T
    if nan_diag:
        ev[2, 2] = np.nan
    E = np.full(cov.shape, np.nan)
    E[np.ix_(idx, idx)] = ev
    return cov, idx, obs, E, ev


def test_ordinary_and_simple_class_methods_match_functions(K):
    cov, idx, obs, E, ev = _class_method_inputs(1)
    S = cov[idx[:, None], idx[None, :]] + ev
    SS = cov[idx, :]
    ok = K.OrdinaryKriging(covariance=cov, idx=idx, obs=obs, error_cov=E.copy())
    k, u, a = ok.solve(), ok.get_uncertainty(), ok.constraint_mask()
    assert k.shape == a.shape == u.shape
    k2, u2 = quiet(K.kriging_ordinary, S, SS, obs, cov)
    assert np.allclose(k2, k) and np.allclose(u2, u)
    sk = K.SimpleKriging(covariance=cov, idx=idx, obs=obs, error_cov=E.copy())
    k, u, a = sk.solve(), sk.get_uncertainty(), sk.constraint_mask()
    k2, u2 = quiet(K.kriging_simple, S, SS, obs, cov)
    a2 = K.constraint_mask(S, SS, cov)
    assert np.allclose(k2, k) and np.allclose(u2, u) and np.allclose(a2, a)
    # against the oracle (reference formulation: LU of the bordered system, 2 alpha variance)
    from oracle import ordinary_kriging, simple_kriging

    ro, rs = ordinary_kriging(cov, idx, obs, E.copy()), simple_kriging(cov, idx, obs, E.copy())
    assert max_norm_rel(ok.solve(), ro["field"]) <= TOL and var_rel(ok.get_uncertainty(), ro["uncertainty"]) <= TOL
    assert max_norm_rel(ok.constraint_mask(), ro["constraint_mask"]) <= TOL
    assert max_norm_rel(ok.kriging_weights, ro["weights"]) <= TOL
    assert max_norm_rel(sk.solve(), rs["field"]) <= TOL and max_norm_rel(sk.kriging_weights, rs["weights"]) <= TOL
    # weights from a pre-computed inverse
    sk2 = K.SimpleKriging(covariance=cov, idx=idx, obs=obs, error_cov=E.copy())
    sk2.kriging_weights_from_inverse(np.linalg.inv(S))
    assert max_norm_rel(sk2.solve(), rs["field"]) <= TOL
    assert var_rel(sk2.get_uncertainty(), rs["uncertainty"]) <= TOL
    assert max_norm_rel(sk2.constraint_mask(), rs["constraint_mask"]) <= TOL
    ok2 = K.OrdinaryKriging(covariance=cov, idx=idx, obs=obs, error_cov=E.copy())
    ok2.kriging_weights_from_inverse(ok2.extended_inverse(np.linalg.inv(S)))
    assert max_norm_rel(ok2.solve(), ro["field"]) <= TOL
    assert var_rel(ok2.get_uncertainty(), ro["uncertainty"]) <= TOL
    assert max_norm_rel(ok2.constraint_mask(), ro["constraint_mask"]) <= TOL
    assert max_norm_rel(ok2.constraint_mask(sk.kriging_weights), ro["constraint_mask"]) <= TOL


def test_stochastic_kriging_class_methods(K):
    cov, idx, obs, E, ev = _class_method_inputs(2)
    st = K.StochasticKriging(covariance=cov, idx=idx, obs=obs, error_cov=E.copy())
    np.random.seed(5)
    k = st.solve()
    u, a = st.get_uncertainty(), st.constraint_mask()
    assert k.shape == a.shape == u.shape == (400,)
    S = cov[idx[:, None], idx[None, :]] + ev
    k2, u2 = quiet(K.kriging_ordinary, S, cov[idx, :], obs, cov)
    assert hasattr(st, "gridded_field") and np.allclose(k2, st.gridded_field) and np.allclose(u2, u)
    assert np.allclose(st.epsilon, st.simulated_grid - (st.simulated_grid - st.epsilon))
    sk_w = st.simple_kriging_weights
    delattr(st, "simple_kriging_weights")
    with pytest.raises(KeyError, match="Please set kriging weights"):
        st.constraint_mask()
    st.set_simple_kriging_weights(sk_w)
    assert np.allclose(a, st.constraint_mask())
    # reproducible under np.random.seed like the reference
    np.random.seed(5)
    st2 = K.StochasticKriging(covariance=cov, idx=idx, obs=obs, error_cov=E.copy())
    assert np.array_equal(st2.solve(), k)
    with pytest.raises(ValueError, match="Error Covariance must be provided"):
        K.StochasticKriging(covariance=cov, idx=idx, obs=obs, error_cov=None)
    inv = np.linalg.inv(S)
    st3 = K.StochasticKriging(covariance=cov, idx=idx, obs=obs, error_cov=E.copy())
    st3.kriging_weights_from_inverse(inv)
    assert max_norm_rel(st3.kriging_weights, st.kriging_weights) <= TOL
    assert max_norm_rel(st3.simple_kriging_weights, sk_w) <= TOL
    with pytest.raises(ValueError, match="inv must be square"):
        st3.kriging_weights_from_inverse(np.eye(5))


@pytest.mark.parametrize("cls", ["OrdinaryKriging", "StochasticKriging", "SimpleKriging"])
def test_filter_bad_error_cov_values(K, cls):
    """reference test/test_kriging.py:375-462: exact warning text, filtered idx / obs / error_cov"""
    cov, idx, obs, E, _ = _class_method_inputs(3, nan_diag=True)
    expected = ("Have nans or zeros on the error covariance diagonal. "
                + f"At positions {idx[2]}. Filtering input accordingly")
    with pytest.warns(UserWarning, match=expected):
        kr = getattr(K, cls)(cov, idx=idx, obs=obs, error_cov=E)
    assert (kr.idx == idx[:2]).all() and (kr.obs == obs[:2]).all() and kr.error_cov.shape == (2, 2)
    assert kr.solve().shape == (400,)


def test_api_errors(K):
    cov, idx, obs, E, _ = _class_method_inputs(4)
    class NoMethod(K.Kriging):          # a subclass that does not declare `method` (reference kriging.py:77-81)
        get_kriging_weights = kriging_weights_from_inverse = solve = get_uncertainty = constraint_mask = None

    with pytest.raises(NotImplementedError, match="Do not use the generic class directly"):
        NoMethod(cov, idx, obs)
    for cls in (K.SimpleKriging, K.OrdinaryKriging):
        kr = cls(cov, idx, obs)
        assert not hasattr(kr, "kriging_weights")
        with pytest.raises(KeyError, match="Please compute Kriging Weights first"):
            kr.get_uncertainty()
        with pytest.raises(ValueError, match="inv must be square"):
            kr.kriging_weights_from_inverse(np.eye(7))
    with pytest.raises(KeyError):
        K.SimpleKriging(cov, idx, obs).constraint_mask()
    # float32 covariance in -> float32 weights / uncertainty out
    ok = K.OrdinaryKriging(cov.astype(np.float32), idx, obs)
    ok.get_kriging_weights()
    assert ok.kriging_weights.dtype == np.float32 and ok.get_uncertainty().dtype == np.float32


# ---- 5-degree scale against the oracle --------------------------------------------------------------------
@pytest.mark.parametrize("n_obs", [800, 1341])
def test_5deg_ordinary_and_simple_vs_oracle(K, n_obs):
    import torch

    from glomargridding_b200.ellipse import EllipseCovarianceBuilder
    from oracle import ordinary_kriging, simple_kriging

    lats, lons = synth.global_grid(5.0)
    Lx, Ly, th, sd = synth.ellipse_fields(lats, lons)
    b = EllipseCovarianceBuilder(Lx, Ly, th, sd, lats, lons, v=1.5, delta_x_method="Met_Office", precision=np.float64)
    cov = b.cov_ns
    idx, obs = synth.fixed_observations(n_obs, cov.shape[0], n_obs)
    E = synth.hadcrut_style_error_cov(idx, lons.size)
    ok = K.OrdinaryKriging(b.cov_ns_device, idx, obs, E.copy())          # device-resident covariance in
    z, u, a = ok.solve(), ok.get_uncertainty(), ok.constraint_mask()
    assert ok._batch.lu_fallback == []
    ro = ordinary_kriging(cov, idx, obs, E, big_temporaries=False)
    assert max_norm_rel(z, ro["field"]) <= TOL and var_rel(u, ro["uncertainty"]) <= TOL
    assert max_norm_rel(a, ro["constraint_mask"]) <= TOL
    assert max_norm_rel(ok.kriging_weights, ro["weights"]) <= TOL
    sk = K.SimpleKriging(cov, idx, obs, E.copy())
    rs = simple_kriging(cov, idx, obs, E, mean=0.5, big_temporaries=False)
    assert max_norm_rel(sk.solve(0.5), rs["field"]) <= TOL
    assert var_rel(sk.get_uncertainty(), rs["uncertainty"]) <= TOL
    assert max_norm_rel(sk.constraint_mask(), rs["constraint_mask"]) <= TOL
    _ = torch


# ---- draws ------------------------------------------------------------------------------------------------
def test_mv_normal_draw_shapes_seed_and_distribution(K):
    rng = np.random.default_rng(0)
    n = 40
    A = rng.standard_normal((n, n))
    cov = A @ A.T / n + 0.05 * np.eye(n)
    loc = np.linspace(-1, 1, n)
    assert K.scipy_mv_normal_draw(loc, cov, 1).shape == (n,)              # reference stochastic.py:461
    np.random.seed(11)
    d1 = K.scipy_mv_normal_draw(loc, cov, 5)
    np.random.seed(11)
    d2 = K.scipy_mv_normal_draw(loc, cov, 5)
    assert d1.shape == (5, n) and np.array_equal(d1, d2)
    nd = 40000
    X = K.scipy_mv_normal_draw(loc, cov, nd, seed=1234)
    se = np.sqrt(np.diag(cov) / nd)
    assert np.all(np.abs(X.mean(0) - loc) < 5 * se)
    emp = np.cov(X.T)
    assert np.linalg.norm(emp - cov) / np.linalg.norm(cov) < 4 * np.sqrt(n / nd)   # Wishart sampling error
    with pytest.raises(ValueError, match="cov should be 2D"):
        K.scipy_mv_normal_draw(loc, np.ones(3))
    with pytest.raises(ValueError, match="not a square matrix"):
        K.scipy_mv_normal_draw(loc, np.ones((3, 4)))


def test_mv_normal_draw_from_indefinite_covariance_like_numpy(K):
    """`np.random.multivariate_normal` (reference stochastic.py:460) draws from its SVD factor, i.e. from V |Lambda| V^T,
    when the covariance is indefinite (and warns).  Same distribution here, through |C| = C sign(C) on the device;
    a StochasticKriging solve on the un-clipped default covariance therefore runs like the reference's does."""
    import torch

    import glomargridding_b200.covariance_tools as ct
    from glomargridding_b200.ellipse import EllipseCovarianceBuilder

    lats, lons = synth.global_grid(20.0)
    Lx, Ly, th, sd = synth.ellipse_fields(lats, lons)
    cov = EllipseCovarianceBuilder(Lx, Ly, th, sd, lats, lons, v=1.5, precision=np.float64).cov_ns
    w, v = np.linalg.eigh(cov)
    assert w.min() < -1e-4                                           # indefinite
    want = (v * np.abs(w)) @ v.T
    got = ct.matrix_abs(torch.from_numpy(cov).cuda()).cpu().numpy()
    assert np.max(np.abs(got - want)) <= 1e-11 * np.abs(want).max()
    n, nd = cov.shape[0], 60000
    with pytest.warns(RuntimeWarning, match="not symmetric positive-semidefinite"):
        X = K.scipy_mv_normal_draw(np.zeros(n), cov, nd, seed=7)
    emp = np.cov(X.T)
    assert np.linalg.norm(emp - want) / np.linalg.norm(want) < 4 * np.sqrt(n / nd)
    assert np.linalg.norm(emp - want) < np.linalg.norm(emp - cov)       # |C|, not C
    idx, obs = synth.fixed_observations(5, n, 60)
    E = synth.hadcrut_style_error_cov(idx, lons.size)
    np.random.seed(3)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        st = K.StochasticKriging(cov, idx, obs, E)
        res = st.solve()
    assert res.shape == (n,) and np.isfinite(res).all()


def test_normal_generator_moments(K):
    import torch

    from glomargridding_b200 import _device as dev
    from glomargridding_b200._lib import check, lib

    n = 2_000_001
    z = torch.empty(n, dtype=torch.float64, device="cuda")
    check(lib().gmg_normal_f64(z.data_ptr(), n, 42, 0, dev.stream_ptr()), "normal")
    z = z.cpu().numpy()
    assert abs(z.mean()) < 5 / np.sqrt(n) and abs(z.var() - 1) < 5 * np.sqrt(2 / n)
    assert abs(np.mean(z ** 3)) < 5 * np.sqrt(15 / n) and abs(np.mean(z ** 4) - 3) < 5 * np.sqrt(96 / n)
    assert np.isfinite(z).all() and np.unique(z).size > 0.999 * n


# ---- batched months (BASELINE.json config 2) ----------------------------------------------------------------
def test_batched_months_equal_per_month_classes(K):
    from glomargridding_b200.batched import krige_batch
    from glomargridding_b200.ellipse import EllipseCovarianceBuilder

    lats, lons = synth.global_grid(10.0)
    Lx, Ly, th, sd = synth.ellipse_fields(lats, lons)
    b = EllipseCovarianceBuilder(Lx, Ly, th, sd, lats, lons, v=1.5, delta_x_method="Met_Office", precision=np.float64)
    M = b.cov_ns.shape[0]
    problems = []
    for m in range(7):
        idx, obs = synth.month_observations(m, M, n_min=40, n_max=300)
        problems.append((idx, obs, synth.hadcrut_style_error_cov(idx, lons.size, seed=m)))
    # one full-grid error covariance with a zero diagonal at an observed cell: the filter must run per problem
    idx, obs = synth.month_observations(99, M, n_min=50, n_max=60)
    E_full = synth.full_grid_error_cov(idx, M, lons.size)
    E_full[idx[3], idx[3]] = 0.0
    problems.append((idx, obs, E_full))
    # NaN entries off the diagonal (reference `kriging.py:120`: NaN -> 0): the classes replace them on the host, the
    # batched driver leaves that to the gather kernel; a float32 error covariance is cast while it is packed
    idx, obs = synth.month_observations(100, M, n_min=200, n_max=260)
    E_nan = synth.hadcrut_style_error_cov(idx, lons.size, seed=100)
    E_nan[3, 150] = E_nan[150, 3] = E_nan[199, 0] = E_nan[0, 199] = np.nan
    problems.append((idx, obs, E_nan))
    problems.append((idx, obs, synth.hadcrut_style_error_cov(idx, lons.size, seed=101).astype(np.float32)))
    for method, cls in (("ordinary", K.OrdinaryKriging), ("simple", K.SimpleKriging)):
        res = quiet(krige_batch, b.cov_ns_device, problems, method, max_bytes=40 << 20)   # forces several chunks
        for m, (idx, obs, E) in enumerate(problems):
            one = quiet(cls, b.cov_ns_device, idx.copy(), obs.copy(), E.copy())
            assert np.array_equal(res["field"][m], quiet(one.solve))
            assert np.array_equal(res["uncertainty"][m], quiet(one.get_uncertainty))
            assert np.array_equal(res["constraint_mask"][m], quiet(one.constraint_mask))


def test_packed_symmetric_error_cov_upload_is_bit_identical_and_asymmetric_falls_back(K, monkeypatch):
    """A symmetric error covariance is uploaded as its lower block trapezoid (`GmgKrigeDesc.lde < 0`), an asymmetric
    one whole.  The packed path must give the same bits as the full upload - on the Cholesky path (SPD covariance)
    and on the LU fallback (indefinite covariance, which reads the mirrored entries) - and an asymmetric E must be
    added as it is, like the reference's `S += error_cov` (`kriging.py:342,576`)."""
    from glomargridding_b200 import _dense as dn
    from glomargridding_b200.ellipse import EllipseCovarianceBuilder
    from oracle import kriging as ok

    lats, lons = synth.global_grid(5.0)
    Lx, Ly, th, sd = synth.ellipse_fields(lats, lons)
    for method in ("Met_Office", "Modified_Met_Office"):          # SPD -> Cholesky; indefinite -> LU fallback
        cov = EllipseCovarianceBuilder(Lx, Ly, th, sd, lats, lons, v=1.5, delta_x_method=method,
                                       precision=np.float64).cov_ns_device
        probs = []
        for m, n in enumerate((129, 300, 517)):                   # ragged last blocks, one block boundary + 1
            idx, obs = synth.month_observations(m, cov.shape[0], n_min=n, n_max=n + 1)
            probs.append((idx.astype(np.int64), obs.astype(np.float64),
                          synth.hadcrut_style_error_cov(idx, lons.size, seed=m)))
        packed = quiet(lambda: dn.KrigeBatch(cov, probs, ordinary=True).run())
        assert packed.h2d_bytes < 0.7 * sum(8 * len(p[0]) ** 2 for p in probs)
        assert bool(packed.lu_fallback) == (method == "Modified_Met_Office")
        with monkeypatch.context() as mp:
            mp.setattr(dn, "PACK_SYMMETRIC_ERROR_COV", False)
            full = quiet(lambda: dn.KrigeBatch(cov, probs, ordinary=True).run())
        assert full.h2d_bytes > 1.4 * packed.h2d_bytes
        assert np.array_equal(packed.out.cpu().numpy(), full.out.cpu().numpy())
    # asymmetric E: detected, uploaded whole, result = LU solve with the asymmetric matrix (oracle)
    idx, obs, E = probs[1]
    E = E.copy()
    E[np.triu_indices_from(E, 1)] *= 0.5
    assert not dn._looks_symmetric(E)
    C = cov.cpu().numpy()
    got = quiet(lambda: K.OrdinaryKriging(cov, idx.copy(), obs.copy(), E.copy()).solve())
    want = ok.ordinary_kriging(C, idx, obs, E, big_temporaries=False)["field"]
    assert max_norm_rel(got, want) < 1e-8


def test_float32_error_covariances_stay_float32_and_give_the_same_bits(K, monkeypatch):
    """HadCRUT's error covariances are float32 (tests/conftest.py).  A batch whose matrices are ALL float32 is uploaded
    as float32 (`GMG_KRIGE_E_FLOAT32`: half the staging writes and PCIe bytes) and widened by the gather kernel - the
    same exact widening numpy does in `C[idx][:, idx] + E` (kriging.py:342, :576), so the results carry the bits of the
    float64 upload: packed trapezoid and whole upload, Cholesky path and LU fallback, odd sizes (a matrix with an odd
    number of entries is padded to 8 bytes), no error covariance in between, a lone problem split over several packing
    tasks, an asymmetric matrix found by the exact test (re-uploaded whole, as float64), the stochastic class."""
    from glomargridding_b200 import _dense as dn
    from glomargridding_b200.ellipse import EllipseCovarianceBuilder

    lats, lons = synth.global_grid(5.0)
    Lx, Ly, th, sd = synth.ellipse_fields(lats, lons)
    for method in ("Met_Office", "Modified_Met_Office"):          # SPD -> Cholesky; indefinite -> LU fallback
        cov = EllipseCovarianceBuilder(Lx, Ly, th, sd, lats, lons, v=1.5, delta_x_method=method,
                                       precision=np.float64).cov_ns_device
        p32 = []
        for m, n in enumerate((129, 301, 517, 77)):               # 301 x 301 and 77 x 77: odd numbers of entries
            idx, obs = synth.month_observations(m, cov.shape[0], n_min=n, n_max=n + 1)
            p32.append((idx.astype(np.int64), obs.astype(np.float64),
                        synth.hadcrut_style_error_cov(idx, lons.size, seed=m).astype(np.float32)))
        p32.insert(2, (p32[0][0], p32[0][1], None))
        p64 = [(i, o, None if E is None else E.astype(np.float64)) for i, o, E in p32]
        for pack in (True, False):
            with monkeypatch.context() as mp:
                mp.setattr(dn, "PACK_SYMMETRIC_ERROR_COV", pack)
                a = quiet(lambda: dn.KrigeBatch(cov, p32, ordinary=True).run())
                b = quiet(lambda: dn.KrigeBatch(cov, p64, ordinary=True).run())
            assert a.error_cov_dtype == np.float32 and b.error_cov_dtype == np.float64
            assert a.h2d_bytes < 0.62 * b.h2d_bytes
            assert bool(a.lu_fallback) == bool(b.lu_fallback) == (method == "Modified_Met_Office")
            assert np.array_equal(a.out.cpu().numpy(), b.out.cpu().numpy())
    # a lone month-sized problem: several packing tasks per matrix, exact symmetry test; and one entry off
    idx, obs = synth.month_observations(5, cov.shape[0])
    E32 = synth.hadcrut_style_error_cov(idx, lons.size, seed=5).astype(np.float32)
    cov = EllipseCovarianceBuilder(Lx, Ly, th, sd, lats, lons, v=1.5, delta_x_method="Met_Office",
                                   precision=np.float64).cov_ns_device
    for E in (E32, _one_entry_off(E32)):
        k32 = quiet(K.OrdinaryKriging, cov, idx.copy(), obs.copy(), E.copy())
        k64 = quiet(K.OrdinaryKriging, cov, idx.copy(), obs.copy(), E.astype(np.float64))
        assert np.array_equal(quiet(k32.solve), quiet(k64.solve))
        assert np.array_equal(quiet(k32.get_uncertainty), quiet(k64.get_uncertainty))
    # a mixed batch keeps the float64 upload (float32 widened while packing)
    mixed = [p32[0], p64[1]]
    assert quiet(lambda: dn.KrigeBatch(cov, mixed, ordinary=True)).error_cov_dtype == np.float64
    # the stochastic ensemble factors E itself from the same (float32) upload
    from glomargridding_b200.batched import stochastic_ensemble

    s32 = quiet(stochastic_ensemble, cov, idx, obs, E32, n_members=4, seed=3)
    s64 = quiet(stochastic_ensemble, cov, idx, obs, E32.astype(np.float64), n_members=4, seed=3)
    for key in s32:
        if isinstance(s32[key], np.ndarray):
            assert np.array_equal(s32[key], s64[key]), key


def _one_entry_off(E):
    E = E.copy()
    E[700, 3] *= 1.5                     # asymmetric in one entry: passes the sampled pre-filter, fails the exact test
    return E


def test_rhs_gather_fused_into_the_solve_is_bit_identical(K, monkeypatch):
    """`gmg_krige_solve_f64` reads the right-hand sides C[idx, :] of every block row inside the forward-solve step instead
    of from a copy written by the gather kernel.  Same bits as gather + `gmg_potrf_solve_f64`: results, the solved
    right-hand sides T and the factors S - ragged batch (sizes around the 128-block and 512-column boundaries, one
    larger than the K limit of the two-CTA tile), simple and ordinary, no error covariance; a covariance of odd order
    takes the separate calls by itself."""
    from glomargridding_b200 import _dense as dn
    from glomargridding_b200.ellipse import EllipseCovarianceBuilder

    lats, lons = synth.global_grid(5.0)
    Lx, Ly, th, sd = synth.ellipse_fields(lats, lons)
    cov = EllipseCovarianceBuilder(Lx, Ly, th, sd, lats, lons, v=1.5, delta_x_method="Met_Office",
                                   precision=np.float64).cov_ns_device
    probs = []
    for m, n in enumerate((1, 127, 129, 511, 513, 1341, 2300)):
        idx, obs = synth.month_observations(m, cov.shape[0], n_min=n, n_max=n + 1)
        E = None if m == 2 else synth.hadcrut_style_error_cov(idx, lons.size, seed=m)
        probs.append((idx.astype(np.int64), obs.astype(np.float64), E))
    for ordinary in (True, False):
        fused = quiet(lambda: dn.KrigeBatch(cov, probs, ordinary=ordinary).run())
        with monkeypatch.context() as mp:
            mp.setattr(dn, "FUSED_RHS_GATHER", False)
            plain = quiet(lambda: dn.KrigeBatch(cov, probs, ordinary=ordinary).run())
        assert not fused.lu_fallback and not plain.lu_fallback
        assert np.array_equal(fused.out.cpu().numpy(), plain.out.cpu().numpy())
        for b in range(len(probs)):
            n = fused.ns[b]
            assert np.array_equal(fused._T_view(b)[:, :fused.nrhs].cpu().numpy(), plain._T_view(b)[:, :plain.nrhs].cpu().numpy()), b
            assert np.array_equal(np.tril(fused.S(b)[:, :n].cpu().numpy()), np.tril(plain.S(b)[:, :n].cpu().numpy())), b
    # odd order: 2591 cells
    import torch

    odd = cov[:2591, :2591].contiguous()
    p2 = [(np.sort(np.random.default_rng(1).choice(2591, 300, replace=False)).astype(np.int64),
           np.random.default_rng(2).standard_normal(300), None)]
    a = quiet(lambda: dn.KrigeBatch(odd, p2, ordinary=True).run())
    with monkeypatch.context() as mp:
        mp.setattr(dn, "FUSED_RHS_GATHER", False)
        b_ = quiet(lambda: dn.KrigeBatch(odd, p2, ordinary=True).run())
    assert np.array_equal(a.out.cpu().numpy(), b_.out.cpu().numpy())


def test_pipelined_driver_many_small_months(K):
    """40 problems take the pipelined route of `krige_batch` (ramped chunks, uploads on the side stream into the two
    persistent arenas, everything of a chunk in one blob, result planes copied home per chunk); a second call reuses
    arenas, staging buffers and the upload stream.  Results must not depend on the chunking: bit-equal to one
    `KrigeBatch` over all problems and to the per-month class."""
    from glomargridding_b200 import _dense as dn
    from glomargridding_b200.batched import krige_batch, release_workspaces
    from glomargridding_b200.ellipse import EllipseCovarianceBuilder

    lats, lons = synth.global_grid(10.0)
    Lx, Ly, th, sd = synth.ellipse_fields(lats, lons)
    cov = EllipseCovarianceBuilder(Lx, Ly, th, sd, lats, lons, v=1.5, delta_x_method="Met_Office",
                                   precision=np.float64).cov_ns_device
    M = cov.shape[0]
    problems = []
    for m in range(40):
        idx, obs = synth.month_observations(m, M, n_min=100 + 5 * m, n_max=140 + 5 * m)   # some above 128: packed E
        problems.append((idx.astype(np.int64), obs.astype(np.float64), synth.hadcrut_style_error_cov(idx, lons.size, seed=m)))
    problems[7] = (problems[7][0], problems[7][1], None)                                  # no error covariance at all
    first = quiet(krige_batch, cov, problems, "ordinary")
    again = quiet(krige_batch, cov, problems, "ordinary")
    for key in ("field", "uncertainty", "constraint_mask"):
        assert first[key].shape == (40, M) and np.array_equal(first[key], again[key])
    whole = quiet(lambda: dn.KrigeBatch(cov, problems, ordinary=True).run()).out.cpu().numpy()
    assert np.array_equal(first["field"], whole[:, dn.OUT_FIELD])
    assert np.array_equal(first["constraint_mask"], whole[:, dn.OUT_MASK])
    for m in (0, 7, 39):
        idx, obs, E = problems[m]
        one = quiet(K.OrdinaryKriging, cov, idx.copy(), obs.copy(), None if E is None else E.copy())
        assert np.array_equal(first["field"][m], quiet(one.solve))
        assert np.array_equal(first["uncertainty"][m], quiet(one.get_uncertainty))
    dev = quiet(krige_batch, cov, problems, "simple", mean=0.5, return_device=True)
    host = quiet(krige_batch, cov, problems, "simple", mean=0.5)
    assert np.array_equal(dev["field"].cpu().numpy(), host["field"])
    assert np.array_equal(dev["constraint_mask"].cpu().numpy(), host["constraint_mask"])
    release_workspaces()


def test_narrow_and_wide_solve_tiles_are_bit_identical(K):
    """The triangular-solve steps use 64-column strips when 128-column ones would leave SMs idle (one 5-degree month:
    21 strips) and 128-column strips otherwise (eight copies of it: 168); every output element is accumulated over k
    in the same order, so the results must be bit-equal."""
    from glomargridding_b200 import _dense as dn
    from glomargridding_b200.ellipse import EllipseCovarianceBuilder

    lats, lons = synth.global_grid(5.0)
    Lx, Ly, th, sd = synth.ellipse_fields(lats, lons)
    cov = EllipseCovarianceBuilder(Lx, Ly, th, sd, lats, lons, v=1.5, delta_x_method="Met_Office",
                                   precision=np.float64).cov_ns_device
    idx, obs = synth.month_observations(3, cov.shape[0], n_min=700, n_max=701)
    E = synth.hadcrut_style_error_cov(idx, lons.size, seed=3)
    prob = (idx.astype(np.int64), obs.astype(np.float64), E)
    one = dn.KrigeBatch(cov, [prob], ordinary=True).run()
    many = dn.KrigeBatch(cov, [prob] * 8, ordinary=True).run()
    a, b = one.out[0].cpu().numpy(), many.out.cpu().numpy()
    for k in range(8):
        assert np.array_equal(a, b[k])
    assert np.array_equal(one.weights(0).cpu().numpy(), many.weights(5).cpu().numpy())   # backward solve as well


# ---- stochastic ensemble (BASELINE.json config 5) -----------------------------------------------------------
def test_stochastic_ensemble_distribution_and_consistency(K):
    """200+ perturbed fields in one call: the ordinary-kriging part equals the class, the perturbations are
    zero-mean with the covariance the two-stage construction implies (the simple-kriging posterior covariance
    C - W_sk C_cross, Morice et al. 2021), and every member equals the class's formula on its own draws."""
    from glomargridding_b200.batched import stochastic_ensemble
    from glomargridding_b200.ellipse import EllipseCovarianceBuilder

    lats, lons = synth.global_grid(15.0)
    Lx, Ly, th, sd = synth.ellipse_fields(lats, lons)
    cov = EllipseCovarianceBuilder(Lx, Ly, th, sd, lats, lons, v=1.5, delta_x_method="Met_Office",
                                   precision=np.float64).cov_ns
    M = cov.shape[0]
    idx, obs = synth.fixed_observations(8, M, 90)
    E = synth.hadcrut_style_error_cov(idx, lons.size)
    nm = 6000
    res = stochastic_ensemble(cov, idx, obs, E.copy(), nm, seed=2014)
    assert res["result"].shape == (nm, M) and res["epsilon"].shape == (nm, M)
    ok = K.OrdinaryKriging(cov, idx, obs, E.copy())
    assert max_norm_rel(res["gridded_field"], ok.solve()) <= 1e-12
    assert var_rel(res["uncertainty"], ok.get_uncertainty()) <= 1e-12
    assert np.allclose(res["result"], res["gridded_field"][None, :] + res["epsilon"])
    eps = res["epsilon"]
    S = cov[np.ix_(idx, idx)] + E
    W = np.linalg.solve(S, cov[idx, :]).T                                   # simple kriging weights
    # epsilon = W (x_obs + noise) - x: covariance W S W^T - W C_o. - C_.o W^T + C = C - W C_cross (posterior)
    post = cov - W @ cov[idx, :]
    se = np.sqrt(np.diag(post) / nm)
    assert np.all(np.abs(eps.mean(0)) < 5 * se + 1e-12)
    emp = np.cov(eps.T)
    assert np.linalg.norm(emp - post) / np.linalg.norm(post) < 4 * np.sqrt(M / nm)
    # same seed -> same members; different seed -> different members
    again = stochastic_ensemble(cov, idx, obs, E.copy(), 5, seed=2014)
    other = stochastic_ensemble(cov, idx, obs, E.copy(), 5, seed=2015)
    assert again["result"].shape == (5, M) and not np.allclose(again["result"], other["result"])


def test_stochastic_ensemble_at_config5_size(K):
    """BASELINE configs[4] at its own size: 5-degree grid (M = 2 592), 1 341 observations, HadCRUT-style error
    covariance, 200 members.  Deterministic parts against the oracle (the reference's formulation) at 1e-8; every member
    against the class's own formula on the member's simulated state and observations (epsilon = W_sk sim_obs - state);
    the ensemble mean of epsilon within 5 standard errors of zero."""
    from glomargridding_b200.batched import stochastic_ensemble
    from glomargridding_b200.ellipse import EllipseCovarianceBuilder
    from oracle import ordinary_kriging

    lats, lons = synth.global_grid(5.0)
    Lx, Ly, th, sd = synth.ellipse_fields(lats, lons)
    cov = EllipseCovarianceBuilder(Lx, Ly, th, sd, lats, lons, v=1.5, delta_x_method="Met_Office",
                                   precision=np.float64).cov_ns
    M = cov.shape[0]
    idx, obs = synth.fixed_observations(201403, M, 1341)
    E = synth.hadcrut_style_error_cov(idx, lons.size)
    res = stochastic_ensemble(cov, idx, obs, E.copy(), 200, seed=20140371)
    assert res["result"].shape == (200, M) and np.isfinite(res["result"]).all()
    want = ordinary_kriging(cov, idx, obs, E, big_temporaries=False)
    assert max_norm_rel(res["gridded_field"], want["field"]) <= TOL
    assert var_rel(res["uncertainty"], want["uncertainty"]) <= TOL
    assert np.allclose(res["result"], res["gridded_field"][None, :] + res["epsilon"])
    S = cov[np.ix_(idx, idx)] + E
    post = cov - cov[:, idx] @ np.linalg.solve(S, cov[idx, :])               # simple-kriging posterior covariance
    se = np.sqrt(np.clip(np.diag(post), 0, None) / 200)
    assert np.all(np.abs(res["epsilon"].mean(0)) < 5 * se + 1e-9)
    assert 0.5 < np.mean(res["epsilon"].var(0, ddof=1) / np.maximum(np.diag(post), 1e-12)) < 1.5
    same = stochastic_ensemble(cov, idx, obs, E.copy(), 200, seed=20140371)
    assert np.array_equal(same["result"], res["result"])                    # reproducible for a given seed


# ---- full-size properties (sizes of BASELINE.json configs[2] + [3]: 1-degree grid, 20 000 observations) ------
def test_1deg_20000_obs_size_independent_properties(K):
    """Too large for the oracle (minutes of LAPACK): checked through properties that do not depend on size.
    (i) linearity of the kriged field in the observations, (ii) the solve residual S t = c on sampled grid columns
    and S o = 1 for the Lagrange column, (iii) unbiasedness: ordinary-kriging weights of every sampled grid point
    sum to one, (iv) the variance / mask identities between the forward-only results and the full weights."""
    import torch

    from glomargridding_b200 import _dense as dn
    from glomargridding_b200.ellipse import EllipseCovarianceBuilder

    lats, lons = synth.global_grid(1.0)
    Lx, Ly, th, sd = synth.ellipse_fields(lats, lons)
    b = EllipseCovarianceBuilder(Lx, Ly, th, sd, lats, lons, v=1.5, delta_x_method="Met_Office", precision=np.float64)
    C = b.cov_ns_device
    N, n = C.shape[0], 20000
    idx, obs1 = synth.fixed_observations(4, N, n)
    rng = np.random.default_rng(1)
    obs2 = rng.standard_normal(n)
    evar = rng.uniform(0.2, 0.4, n) ** 2
    E = np.diag(evar)

    def krige(y):
        k = K.OrdinaryKriging(C, idx, y, E)
        return k, k.solve(), k.get_uncertainty(), k.constraint_mask()

    k1, z1, u1, m1 = krige(obs1)
    assert k1._batch.lu_fallback == [] and np.isfinite(z1).all() and np.isfinite(u1).all()
    _, z2, u2, _ = krige(obs2)
    _, z3, u3, _ = krige(obs1 + 2.0 * obs2)
    assert max_norm_rel(z3, z1 + 2.0 * z2) <= TOL                       # (i) linear in the observations
    assert np.array_equal(u1, u2) and np.array_equal(u1, u3)            #     variance does not depend on them
    assert (m1 > 0).all() and (m1 <= 1 + 1e-9).all() and (u1 >= 0).all()
    # (ii) residuals with the full weights of the first system (runs the second triangular solve)
    batch = k1._batch
    T = batch.T(0)                                                      # (n, M + 2) = S^-1 [C_cross | 1 | y]
    idx_t = torch.from_numpy(idx).cuda()
    S = C[idx_t][:, idx_t] + torch.from_numpy(E).cuda()
    cols = torch.tensor([0, 777, 32400, 64799, N, N + 1], device="cuda")          # four grid columns, ones, obs
    rhs = torch.cat([C[idx_t][:, cols[:4]], torch.ones(n, 1, dtype=torch.float64, device="cuda"),
                     torch.from_numpy(obs1).cuda()[:, None]], dim=1)
    resid = S @ T[:, cols] - rhs
    assert float(resid.abs().max() / rhs.abs().max()) <= 1e-9
    # (iii) + (iv) on sampled grid points, from the materialised rows of the weights matrix
    o = T[:, N]
    alpha = batch.out[0, dn.OUT_ALPHA]
    for m in (0, 777, 32400, 64799):
        w = T[:, m] - o * alpha[m]                                      # ordinary-kriging weights of grid point m
        assert abs(float(w.sum()) - 1.0) <= 1e-9
        assert abs(float(w @ torch.from_numpy(obs1).cuda()) - z1[m]) <= TOL * np.abs(z1).max()
        c_m = C[idx_t, m]
        u2_m = float(C[m, m] - w @ c_m - 2 * alpha[m])
        assert abs(u2_m - u1[m] ** 2) <= TOL * float(C[m, m])
        assert abs(float(T[:, m] @ c_m / C[m, m]) - m1[m]) <= TOL


def test_notebook_2014_03_field_configs0():
    """BASELINE configs[0] on the device: 5-degree ellipse covariance build -> eigenvalue clip -> simple and ordinary
    kriging of the real HadSST.4.0.1.0 2014-03 ensemble-member field with the real HadCRUT.5 error covariance
    (`notebooks/Ellipse_to_grid.ipynb` cells 17-34), against the UNMODIFIED reference's outputs
    (`tests/golden/make_golden_notebook.py`).  Tolerance: 1e-8 max-norm relative."""
    import torch
    from conftest import load_notebook_case

    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import glomargridding_b200.covariance_tools as ct
    from glomargridding_b200.ellipse import EllipseCovarianceBuilder
    from glomargridding_b200.kriging import OrdinaryKriging, SimpleKriging

    g, E = load_notebook_case()
    lats, lons = synth.global_grid(5.0)
    Lx, Ly, th, sd = synth.ellipse_fields(lats, lons)
    b = EllipseCovarianceBuilder(Lx, Ly, th, sd, lats, lons, v=1.5, precision=np.float64)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        cov = ct.eigenvalue_clip(b.cov_ns_device)                     # notebook cell 20, device in -> device out
        assert max_norm_rel(torch.diagonal(cov).cpu().numpy(), g["cov_diag_clipped"]) <= 1e-8
        idx, obs = g["idx"], g["obs"]
        sk = SimpleKriging(cov, idx.copy(), obs.copy(), E.copy())
        ok = OrdinaryKriging(cov, idx.copy(), obs.copy(), E.copy())
        for k, tag in ((sk, "sk"), (ok, "ok")):
            z, u, m = k.solve(), k.get_uncertainty(), k.constraint_mask()
            assert k._batch.lu_fallback == []                         # Cholesky path, no LU fallback
            assert z.shape == (2592,)
            assert max_norm_rel(z, g[f"{tag}_field"]) <= 1e-8
            assert max_norm_rel(u ** 2, g[f"{tag}_unc"] ** 2) <= 1e-8
            assert max_norm_rel(m, g[f"{tag}_mask"]) <= 1e-8
    assert np.array_equal(ok.idx, g["kept_idx"])
