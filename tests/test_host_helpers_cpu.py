"""Host-side helpers of the hot path (`utils.py`, the observation-preparation helpers of `kriging.py`): pure numpy
logic that runs without a GPU.  Fixed expectations first (they hold on the GPU box too), then a comparison with the
live reference on random inputs (skipped where `/root/reference` is absent)."""

import contextlib
import importlib
import io
import warnings

import numpy as np
import pytest


def test_adjust_small_negative_texts_and_values(capsys):
    """reference utils.py:132-175: (-1e-8, 0) -> 0 with a warning and a print of the values; other negatives only warn."""
    from glomargridding_b200.utils import adjust_small_negative

    a = np.array([1.0, -5e-9, 0.0, 2.0])
    with pytest.warns(UserWarning, match="Small negative vals are detected. Setting to 0."):
        out = adjust_small_negative(a)
    assert np.array_equal(out, [1.0, 0.0, 0.0, 2.0]) and a[1] == -5e-9            # input untouched
    assert "-5.e-09" in capsys.readouterr().out
    with pytest.warns(UserWarning, match="Negative values are detected"):
        out = adjust_small_negative(np.array([1.0, -0.5]))
    assert np.array_equal(out, [1.0, -0.5])
    with warnings.catch_warnings():
        warnings.simplefilter("error")
        assert np.array_equal(adjust_small_negative(np.array([0.0, 3.0], dtype=np.float32)), [0.0, 3.0])


def test_mask_array_and_uncompress_masked():
    """reference utils.py:674-694 (mask_array) and :550-608 (uncompress_masked)."""
    from glomargridding_b200.utils import mask_array, uncompress_masked

    m = mask_array(np.array([[1.0, np.nan], [3.0, 4.0]]))
    assert isinstance(m, np.ma.MaskedArray)
    already = np.ma.masked_greater(np.arange(4.0), 2)
    assert np.array_equal(np.ma.getmaskarray(mask_array(already)), [False, False, False, True])
    mask = np.array([False, True, False, True, False])
    full = uncompress_masked(np.array([1.0, 2.0, 3.0]), mask, fill_value=-9.0)
    assert np.array_equal(np.asarray(full), [1.0, -9.0, 2.0, -9.0, 3.0])
    masked = uncompress_masked(np.array([1.0, 2.0, 3.0]), mask, apply_mask=True)
    assert isinstance(masked, np.ma.MaskedArray) and np.array_equal(np.ma.getmaskarray(masked), mask)
    with pytest.raises(ValueError):
        uncompress_masked(np.array([1.0, 2.0]), mask)


def test_batched_and_intersect():
    """reference utils.py:697-712 (batched), :246-279 (intersect_mtlb)."""
    from glomargridding_b200.utils import batched, intersect_mtlb

    assert [tuple(b) for b in batched(range(7), 3)] == [(0, 1, 2), (3, 4, 5), (6,)]
    with pytest.raises(ValueError):
        list(batched(range(3), 0))
    c, ia, ib = intersect_mtlb(np.array([5, 1, 3, 3, 9]), np.array([3, 9, 7, 5]))
    assert np.array_equal(c, [3, 5, 9]) and np.array_equal(np.array([5, 1, 3, 3, 9])[ia], c)
    assert np.array_equal(np.array([3, 9, 7, 5])[ib], c)


@pytest.mark.gpu
def test_get_spatial_mean_on_device():
    """reference utils.py:766-793: the BLUE of a constant mean, 1^T S^-1 y / 1^T S^-1 1 (the solve runs on the GPU)."""
    import torch

    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from glomargridding_b200.utils import get_spatial_mean

    rng = np.random.default_rng(3)
    for n in (2, 7, 150):
        A = rng.standard_normal((n, n))
        covx = A @ A.T + n * np.eye(n)
        obs = rng.standard_normal(n)
        ones = np.ones(n)
        want = (ones @ np.linalg.solve(covx, obs)) / (ones @ np.linalg.solve(covx, ones))
        assert np.isclose(get_spatial_mean(obs, covx), want, rtol=1e-10)


def test_host_helpers_against_the_live_reference(reference):
    import glomargridding_b200.kriging as mk
    import glomargridding_b200.utils as mu

    ru = importlib.import_module("glomar_gridding.utils")
    rk = importlib.import_module("glomar_gridding.kriging")
    rng = np.random.default_rng(11)
    with warnings.catch_warnings(), contextlib.redirect_stdout(io.StringIO()):
        warnings.simplefilter("ignore")
        for _ in range(20):
            v = rng.standard_normal(30) * 10.0 ** rng.integers(-10, 1, 30)
            assert np.array_equal(ru.adjust_small_negative(v), mu.adjust_small_negative(v))
            a, b = rng.integers(0, 40, 25), rng.integers(0, 40, 18)
            for x, y in zip(ru.intersect_mtlb(a, b), mu.intersect_mtlb(a, b)):
                assert np.array_equal(x, y)
            mask = rng.random(50) < 0.4
            comp = rng.standard_normal(int((~mask).sum()))
            for apply_mask in (False, True):
                x = ru.uncompress_masked(comp, mask, fill_value=1.5, apply_mask=apply_mask)
                y = mu.uncompress_masked(comp, mask, fill_value=1.5, apply_mask=apply_mask)
                assert np.array_equal(np.ma.getmaskarray(x), np.ma.getmaskarray(y))
                assert np.array_equal(np.ma.filled(x, 0.0), np.ma.filled(y, 0.0))   # data under the mask is np.empty
            unmask = np.sort(rng.choice(60, 30, replace=False))
            uniq = np.sort(rng.choice(unmask, 12, replace=False))
            assert np.array_equal(rk.get_unmasked_obs_indices(unmask, uniq), mk.get_unmasked_obs_indices(unmask, uniq))


def test_exact_symmetry_test_of_error_covariances():
    """`gmg_host_rows_symmetric` (native, host only) behind `_dense.is_symmetric` / the packing tasks of `KrigeBatch`:
    every mirrored pair of entries, NaN equal to NaN, float64 / float32 / other layouts."""
    from glomargridding_b200 import _dense as dn
    from glomargridding_b200._lib import lib

    rng = np.random.default_rng(0)
    for n in (1, 5, 127, 128, 129, 700):
        A = rng.standard_normal((n, n))
        S = A + A.T
        assert dn.is_symmetric(S) and dn.is_symmetric(S.astype(np.float32)) and dn.is_symmetric(np.asfortranarray(S))
        if n > 1:
            for (i, j) in ((0, n - 1), (n - 1, 0), (n // 2, n // 2 - 1)):
                S2 = S.copy()
                S2[i, j] = np.nextafter(S2[i, j], np.inf)          # one ulp in one entry
                assert not dn.is_symmetric(S2)
            S3 = S.copy()
            S3[n // 2, 0] = np.nan
            assert not dn.is_symmetric(S3)
            S3[0, n // 2] = np.nan
            assert dn.is_symmetric(S3)                             # NaN mirrors NaN: both are read as 0
        # row blocks: rows r0:r1 against columns r0:r1 only
        if n >= 129:
            S4 = S.copy()
            S4[3, 100] += 1.0                                      # lives in row block 0 / column block 0
            assert not dn._block_row_symmetric(S4, 0, 128) and dn._block_row_symmetric(S4, 128, n)
    E = np.zeros((4, 4))
    assert lib().gmg_host_rows_symmetric(None, 0, 4, 4, 0, 4) == -1
    assert lib().gmg_host_rows_symmetric(E.ctypes.data, 0, 4, 3, 0, 4) == -3
    assert lib().gmg_host_rows_symmetric(E.ctypes.data, 0, 4, 4, 2, 5) == -5


def test_native_packers_of_the_upload_pipeline():
    """`gmg_host_pack_f64` (float64 or float32 source, widened) and `gmg_host_pack_f32` (float32 kept float32): strided
    blocks -> contiguous staging memory, any destination alignment, NaN preserved; and the blob layout rule of
    `_group_doubles` (float32 groups padded to whole doubles)."""
    from glomargridding_b200 import _dense as dn
    from glomargridding_b200._lib import lib

    L = lib()
    rng = np.random.default_rng(3)
    E = rng.standard_normal((301, 301))
    E[5, 7] = np.nan
    for dtype in (np.float64, np.float32):
        Es = E.astype(dtype)
        for (r0, r1, c) in ((0, 128, 128), (128, 256, 256), (256, 301, 301), (0, 77, 77), (10, 11, 3)):
            blk = Es[r0:r1, :c]
            for shift in (0, 1, 2, 3):
                buf = np.zeros(blk.size + 8, dtype=np.float64)
                assert L.gmg_host_pack_f64(blk.ctypes.data, int(dtype == np.float32), blk.shape[0], blk.shape[1],
                                           blk.strides[0] // blk.itemsize, buf[shift:].ctypes.data) == 0
                assert np.array_equal(buf[shift:shift + blk.size].reshape(blk.shape), blk.astype(np.float64), equal_nan=True)
                assert not buf[shift + blk.size:].any() and not buf[:shift].any()
                if dtype == np.float32:
                    b32 = np.zeros(blk.size + 8, dtype=np.float32)
                    assert L.gmg_host_pack_f32(blk.ctypes.data, blk.shape[0], blk.shape[1], blk.strides[0] // 4,
                                               b32[shift:].ctypes.data) == 0
                    assert np.array_equal(b32[shift:shift + blk.size].reshape(blk.shape), blk, equal_nan=True)
                    assert not b32[shift + blk.size:].any() and not b32[:shift].any()
    assert L.gmg_host_pack_f32(None, 1, 1, 1, E.ctypes.data) == -1
    assert L.gmg_host_pack_f32(E.ctypes.data, 2, 4, 3, E.ctypes.data) == -3
    a32, a64 = np.zeros((3, 5), dtype=np.float32), np.zeros((3, 5))
    assert dn._group_doubles([a32, a32], True) == 15 and dn._group_doubles([a32], True) == 8      # 30 -> 15, 15 -> 8
    assert dn._group_doubles([a32, a64], True) == 23 and dn._group_doubles([a32], False) == 15
