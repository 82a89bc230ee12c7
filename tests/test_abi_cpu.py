"""CPU-side checks of the drop-in boundary: the C-ABI library loads and exports every symbol declared in
include/glomar_b200.h, descriptor layouts match, and the product package fails loudly without a GPU
(no compute calls are made here)."""

import ctypes
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    from glomargridding_b200._lib import LIB_PATH, exported_symbols, lib

    assert os.path.exists(LIB_PATH), "run `python __graft_entry__.py` first"
    L = lib()
    names = exported_symbols()
    assert len(names) >= 20
    missing = [s for s in names if not hasattr(L, s)]
    assert not missing, missing
    assert L.gmg_version() >= 100
    assert L.gmg_strerror(2).decode().startswith("matrix is not positive definite")
    assert b"invalid argument" in L.gmg_strerror(-3)


def test_every_entry_point_is_mapped_to_the_reference_in_integration_md():
    """INTEGRATION.md's table says, for each exported entry point, which reference lines it replaces (or that it has no
    counterpart); a new export without a row fails here."""
    from glomargridding_b200._lib import exported_symbols

    text = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    missing = [s for s in exported_symbols() if s not in text and not (s.startswith("gmg_ipc_") and "gmg_ipc_" in text)]
    assert not missing, missing


def test_argument_validation_without_gpu():
    """Bad arguments are rejected before any CUDA call: negative return = 1-based argument index."""
    from glomargridding_b200._lib import lib

    L = lib()
    nbytes = ctypes.c_size_t()
    assert L.gmg_ellipse_ws_bytes(0, ctypes.byref(nbytes)) == -1
    assert L.gmg_ellipse_ws_bytes(100, ctypes.byref(nbytes)) == 0 and nbytes.value > 100 * 64 * 2
    assert L.gmg_ellipse_cov_f64(None, 10, 0, 10, 1.5, -1.0, 0, None, 10, 0, None) == -1
    assert L.gmg_gemm_f64(0, 0, -1, 1, 1, 1.0, None, 2, None, 2, 0.0, None, 2, None) == -3
    a, b, c = ctypes.c_int64(), ctypes.c_int64(), ctypes.c_int64()
    assert L.gmg_solve_ws_doubles(300, ctypes.byref(a), ctypes.byref(b), ctypes.byref(c)) == 0
    assert (a.value, b.value, c.value) == (3 * 128 * 128, 128 * 384, 384)
    assert L.gmg_potrf_batched_f64(None, 1, 10, None) == -1


def test_descriptor_layouts_match_the_header():
    from glomargridding_b200 import _dense as dn

    assert dn.SOLVE_DESC.itemsize == 72 and dn.KRIGE_DESC.itemsize == 104
    assert list(dn.SOLVE_DESC.names) == ["S", "lds", "Linv", "R", "ldr", "P", "ldp", "info", "n", "nrhs"]
    header = open(os.path.join(ROOT, "include", "glomar_b200.h")).read()
    for name in dn.SOLVE_DESC.names + dn.KRIGE_DESC.names:
        assert name in header


def test_product_fails_loudly_without_cuda():
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from glomargridding_b200 import synthetic as synth
    from glomargridding_b200._lib import GlomarB200Error
    from glomargridding_b200.ellipse import EllipseCovarianceBuilder
    from glomargridding_b200.kriging import OrdinaryKriging

    lats, lons = synth.global_grid(30.0)
    with pytest.raises(GlomarB200Error, match="no CPU fallback"):
        EllipseCovarianceBuilder(*synth.ellipse_fields(lats, lons), lats, lons, v=1.5)
    with pytest.raises(GlomarB200Error, match="no CPU fallback"):
        OrdinaryKriging(np.eye(4), np.array([0, 2]), np.array([1.0, 2.0])).solve()


def test_product_never_imports_the_oracle():
    code = ("import sys; sys.path.insert(0, %r); import glomargridding_b200, glomargridding_b200.ellipse, "
            "glomargridding_b200.kriging, glomargridding_b200.stochastic, glomargridding_b200.batched; "
            "assert not any(m == 'oracle' or m.startswith('oracle.') for m in sys.modules), 'oracle imported'" % ROOT)
    subprocess.run([sys.executable, "-c", code], check=True)
    for dirpath, _, files in os.walk(os.path.join(ROOT, "glomargridding_b200")):
        for f in files:
            if f.endswith(".py"):
                text = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in text and "from oracle" not in text, f


def test_install_as_glomar_gridding_aliases_import_paths():
    code = ("import sys; sys.path.insert(0, %r); import glomargridding_b200 as g; g.install_as_glomar_gridding(); "
            "from glomar_gridding.ellipse import EllipseCovarianceBuilder; "
            "from glomar_gridding.kriging import SimpleKriging, OrdinaryKriging, kriging_simple; "
            "from glomar_gridding.stochastic import StochasticKriging, scipy_mv_normal_draw; "
            "from glomar_gridding.utils import adjust_small_negative, cov_2_cor" % ROOT)
    subprocess.run([sys.executable, "-c", code], check=True)


@pytest.mark.parametrize("n", [129, 256, 300, 1341])
def test_block_trapezoid_layout_of_symmetric_error_covariances(n):
    """Host packing (`KrigeBatch`: row block b of 128 rows with its first min(n, 128 (b + 1)) columns) against the
    index formula of `gather_kernel` (`csrc/kriging.cu`, `GmgKrigeDesc.lde < 0`), restated here in numpy: every
    entry (i, j) with j inside row i's trapezoid - in particular the whole lower triangle and the whole diagonal
    blocks - is found at offset hb^2 b (b + 1) / 2 + (i - hb b) L_b + j."""
    from glomargridding_b200 import _dense as dn

    hb = dn.NB
    rng = np.random.default_rng(n)
    E = rng.standard_normal((n, n))
    E = E + E.T
    assert dn._looks_symmetric(E) and not dn._looks_symmetric(np.triu(E) + 0.5 * np.tril(E, -1))
    packed = np.concatenate([E[r0:min(n, r0 + hb), :min(n, r0 + hb)].reshape(-1) for r0 in range(0, n, hb)])
    assert packed.size == dn.packed_error_cov_doubles(n)
    i, j = np.tril_indices(n)
    extra = (np.arange(n)[:, None] // hb == np.arange(n)[None, :] // hb).nonzero()      # diagonal blocks, both halves
    i, j = np.concatenate([i, extra[0]]), np.concatenate([j, extra[1]])
    b = i // hb
    Lb = np.minimum(n, hb * (b + 1))
    off = hb * hb * (b * (b + 1) // 2) + (i - hb * b) * Lb + j
    assert np.array_equal(packed[off], E[i, j])


# ---- scalar helpers of the clipping path (host logic; reference covariance_tools.py:393-517) ----------------------
@pytest.mark.parametrize("n, expected_i", [(10, 7), (25, 24), (100, 100), (15, 1)])
def test_cumsum(n, expected_i):
    """reference test/test_covariance_tools.py:140-157"""
    from glomargridding_b200.covariance_tools import csum_up_to_val

    vals = np.arange(n + 1)
    target_sum = expected_i * (expected_i + 1) / 2
    csum, i = csum_up_to_val(vals, target_sum - 1, reverse=False)      # cumulatively sum to just under the target
    assert expected_i == i - 1
    assert target_sum == csum


def test_cumsum_exceeds_target_and_clip_indices(reference):
    """reference test/test_covariance_tools.py:160-169, then the helpers against the live reference on random spectra
    (skipped where the reference sources are absent)."""
    import contextlib
    import importlib
    import io
    import warnings

    import glomargridding_b200.covariance_tools as mine

    rng = np.random.default_rng(50)
    vals = np.sort(rng.random(50))
    target_csum = 0.9 * vals.sum()
    csum, _ = mine.csum_up_to_val(vals, target_csum, reverse=True)
    assert csum >= target_csum
    ct = importlib.import_module("glomar_gridding.covariance_tools")
    with warnings.catch_warnings(), contextlib.redirect_stdout(io.StringIO()):
        warnings.simplefilter("ignore")
        for _ in range(100):
            n = int(rng.integers(0, 30))
            v = rng.random(n)
            target = rng.random() * max(v.sum(), 0.1) * 1.2
            for rev in (True, False):
                assert ct.csum_up_to_val(v, target, reverse=rev, niter=2, csum=0.5) == \
                    mine.csum_up_to_val(v, target, reverse=rev, niter=2, csum=0.5)
        ev = np.sort(rng.random(60))
        assert ct._find_index_explained_variance(ev, 0.9) == mine._find_index_explained_variance(ev, 0.9)
        assert ct._find_index_aspect_ratio(ev * 300, 2592, 40) == mine._find_index_aspect_ratio(ev * 300, 2592, 40)
    m = rng.standard_normal((6, 6))
    assert np.array_equal(ct.clean_small(m * 1e-5), mine.clean_small(m * 1e-5))
    assert mine.check_symmetric(m + m.T) and not mine.check_symmetric(m)


def test_cumsum_warnings():
    from glomargridding_b200.covariance_tools import _find_index_explained_variance, csum_up_to_val

    with pytest.warns(UserWarning, match="empty"):
        assert csum_up_to_val(np.array([]), 1.0) == (0.0, 0)
    with pytest.warns(UserWarning, match="target not exceeded"):
        csum_up_to_val(np.ones(4), 10.0)
    with pytest.raises(ValueError, match="must be a vector"):
        csum_up_to_val(np.ones((2, 2)), 1.0)
    with pytest.raises(ValueError, match="not exceeded"), pytest.warns(UserWarning):
        _find_index_explained_variance(np.ones(4), 1.0)
