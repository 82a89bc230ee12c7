"""Device eigenvalue clipping (glomargridding_b200.covariance_tools) against the reference algorithm restated
with numpy.linalg.eigh, and the properties the reference's own tests assert (test/test_covariance_tools.py:14-83:
trace preserved, top eigenvalues unchanged, all eigenvalues positive afterwards)."""

import warnings

import numpy as np
import pytest

from glomargridding_b200 import synthetic as synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def CT():
    import torch

    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import glomargridding_b200.covariance_tools as ct

    return ct


def reference_clip(cov, frac=0.95):
    """`explained_variance_clip` + `_eigenvalue_clip` of the reference, restated (oracle/covariance_tools.py)."""
    from oracle.covariance_tools import explained_variance_clip

    return explained_variance_clip(cov, frac)


@pytest.mark.parametrize("n", [7, 64, 301])
def test_jacobi_eigh_vs_numpy(CT, n):
    import torch

    rng = np.random.default_rng(n)
    A = rng.standard_normal((n, n))
    A = A + A.T
    w, U = CT._jacobi_eigh(torch.from_numpy(A).cuda())
    w, U = w.cpu().numpy(), U.cpu().numpy()
    assert np.allclose(np.sort(w), np.linalg.eigvalsh(A), rtol=0, atol=1e-12 * np.abs(A).max() * n)
    assert np.allclose(U.T @ U, np.eye(n), atol=1e-12)
    assert np.allclose(U @ np.diag(w) @ U.T, A, atol=1e-11 * np.abs(A).max() * n)
    assert np.all(np.diff(w) <= 0)


@pytest.mark.parametrize("n", [1, 2, 3, 5, 64, 301, 1000])
def test_sym_eigvals_vs_numpy(CT, n):
    """Householder tridiagonalisation + Sturm bisection on the device (csrc/spectrum.cu) against numpy eigvalsh."""
    import torch

    rng = np.random.default_rng(100 + n)
    A = rng.standard_normal((n, n))
    A = A + A.T
    if n == 64:
        A[:, 10:20] = 0.0                     # already-tridiagonal columns: zero Householder tails
        A[10:20, :] = 0.0
    w = CT.sym_eigvals(torch.from_numpy(A).cuda()).cpu().numpy()
    want = np.linalg.eigvalsh(A)
    assert w.shape == (n,) and np.all(np.diff(w) >= 0)
    assert np.max(np.abs(w - want)) <= 1e-13 * max(1.0, np.abs(want).max()) * max(n, 8)


def test_projector_route_vs_eigenvector_routes(CT, monkeypatch):
    """The default route (eigenvalues + sign-function projector) and the Jacobi eigenvector route agree."""
    from oracle import EllipseOracle

    lats, lons = synth.global_grid(10.0)
    Lx, Ly, th, sd = synth.ellipse_fields(lats, lons)
    cov = EllipseOracle(Lx, Ly, th, sd, lats, lons, 1.5).build()
    want, k, _ = reference_clip(cov, 0.9)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        got = CT.explained_variance_clip(cov, 0.9)
        monkeypatch.setattr(CT, "SPECTRAL_MAX", 0)
        jac = CT.explained_variance_clip(cov, 0.9)
    assert np.max(np.abs(got - want)) <= 1e-9 * np.max(np.abs(want))
    assert np.max(np.abs(jac - want)) <= 1e-9 * np.max(np.abs(want))
    assert np.array_equal(got, got.T)
    # nothing to clip: every eigenvalue is kept
    eye = CT.explained_variance_clip(np.diag([3.0, 2.0, 1.0]), 1.0 - 1e-12)
    assert np.allclose(eye, np.diag([3.0, 2.0, 1.0]), atol=1e-14)


@pytest.mark.parametrize("res,meth", [(10.0, "Modified_Met_Office"), (5.0, "Modified_Met_Office")])
def test_eigenvalue_clip_vs_reference_algorithm(CT, res, meth):
    from oracle import EllipseOracle

    lats, lons = synth.global_grid(res)
    Lx, Ly, th, sd = synth.ellipse_fields(lats, lons)
    cov = EllipseOracle(Lx, Ly, th, sd, lats, lons, 1.5, meth).build()          # indefinite
    assert np.linalg.eigvalsh(cov).min() < 0
    with warnings.catch_warnings():
        warnings.simplefilter("error")
        got = CT.eigenvalue_clip(cov)
    want, k, w = reference_clip(cov)
    assert got.shape == cov.shape and got.dtype == np.float64
    assert np.max(np.abs(got - want)) <= 1e-8 * np.max(np.abs(want))
    ev = np.linalg.eigvalsh(got)
    assert (ev > 0).all()                                                        # test_covariance_tools.py:80-83
    assert abs(np.trace(got) - np.trace(cov)) <= 1e-9 * np.trace(cov)            # trace preserved
    assert np.allclose(ev[-k:], w[-k:], rtol=1e-9)                               # kept eigenvalues unchanged
    assert np.array_equal(got, got.T)


def test_clip_then_krige_default_workflow(CT):
    """Notebook cells 19-40 on the device: indefinite default covariance -> eigenvalue_clip -> ordinary kriging on
    the Cholesky path (no LU fallback), against the oracle run on the reference-clipped matrix."""
    from glomargridding_b200.ellipse import EllipseCovarianceBuilder
    from glomargridding_b200.kriging import OrdinaryKriging
    from oracle import ordinary_kriging

    lats, lons = synth.global_grid(10.0)
    Lx, Ly, th, sd = synth.ellipse_fields(lats, lons)
    b = EllipseCovarianceBuilder(Lx, Ly, th, sd, lats, lons, v=1.5, precision=np.float64)
    clipped = CT.eigenvalue_clip(b.cov_ns_device)                 # device in -> device out
    idx, obs = synth.fixed_observations(2, clipped.shape[0], 150)
    E = synth.hadcrut_style_error_cov(idx, lons.size)
    ok = OrdinaryKriging(clipped, idx, obs, E.copy())
    z, u = ok.solve(), ok.get_uncertainty()
    assert ok._batch.lu_fallback == []
    want, _, _ = reference_clip(b.cov_ns)
    ref = ordinary_kriging(want, idx, obs, E)
    assert np.max(np.abs(z - ref["field"])) <= 1e-7 * np.max(np.abs(ref["field"]))
    assert np.max(np.abs(u ** 2 - ref["uncertainty"] ** 2)) <= 1e-7 * np.max(ref["uncertainty"] ** 2)


def test_clip_vs_reference_golden(CT):
    """Both clipping methods against the unmodified reference's outputs (tests/golden/clip_small.npz)."""
    from conftest import from_triu, load_golden

    g = load_golden("clip_small.npz")
    cov = from_triu(g["cov"])
    scale = np.abs(cov).max()
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        cases = [(dict(), "ev95"), (dict(method="explained_variance", target_variance_fraction=0.8), "ev80"),
                 (dict(method="Laloux_2000", num_time_pts=40), "laloux_t40"),
                 (dict(method="Laloux_2000", num_time_pts=400, num_grid_pts=cov.shape[0]), "laloux_t400")]
        for kw, tag in cases:
            got = CT.eigenvalue_clip(cov, **kw)
            assert got.shape == cov.shape and got.dtype == np.float64
            assert np.max(np.abs(got - from_triu(g[tag]))) <= 1e-9 * scale, tag


def test_clip_errors(CT):
    with pytest.raises(ValueError, match="target_variance_fraction"):
        CT.explained_variance_clip(np.eye(4), 1.5)
    with pytest.raises(ValueError, match="Unknown clipping method"):
        CT.eigenvalue_clip(np.eye(4), method="nope")
    with pytest.raises(ValueError, match="Correlation Diagonal"):
        CT.eigenvalue_clip(np.diag([1.0, 0.0, 2.0]), method="Laloux_2000")


def test_subspace_iteration_path_matches_full_jacobi(CT, monkeypatch):
    """Force the large-matrix route (orthogonal iteration + Rayleigh-Ritz) on a matrix the direct route can check."""
    from oracle import EllipseOracle

    lats, lons = synth.global_grid(10.0)
    Lx, Ly, th, sd = synth.ellipse_fields(lats, lons)
    cov = EllipseOracle(Lx, Ly, th, sd, lats, lons, 1.5).build()
    want, k, _ = reference_clip(cov, 0.9)
    monkeypatch.setattr(CT, "FULL_JACOBI_MAX", 0)
    monkeypatch.setattr(CT, "SPECTRAL_MAX", 0)
    got = CT.explained_variance_clip(cov, 0.9)
    assert np.max(np.abs(got - want)) <= 1e-8 * np.max(np.abs(want))
