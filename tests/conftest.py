"""pytest configuration: `gpu` marker, repo root on sys.path, shared fixture loaders."""

import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on a B200)")


def load_golden(name: str):
    return np.load(os.path.join(GOLDEN, name), allow_pickle=False)


def from_triu(vec: np.ndarray) -> np.ndarray:
    """Rebuild a symmetric matrix from its row-major upper triangle (diagonal included)."""
    n = int((np.sqrt(8 * vec.size + 1) - 1) / 2)
    out = np.zeros((n, n), dtype=vec.dtype)
    iu = np.triu_indices(n)
    out[iu] = vec
    out.T[iu] = vec
    return out


def max_rel(a, b, floor=1e-290):
    """Entrywise max |a-b| / max(|b|, floor) - the covariance parity measure."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), floor))) if a.size else 0.0


def max_norm_rel(a, b):
    """||a-b||_inf / ||b||_inf - the kriging parity measure."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b)) / np.max(np.abs(b)))


@pytest.fixture(scope="session")
def reference():
    from oracle.reference_shim import load_reference, reference_available

    if not reference_available():
        pytest.skip("reference sources not present (GPU box)")
    return load_reference()


def load_notebook_case():
    """configs[0] fixture (tests/golden/make_golden_notebook.py): the HadSST.4.0.1.0 2014-03 field of the reference's
    notebook with the HadCRUT.5 error covariance of that month, plus the unmodified reference's kriging outputs.
    Returns (golden npz, full-grid float32 error covariance rebuilt from its non-zero entries)."""
    g = load_golden("notebook_2014_03.npz")
    E = np.zeros((2592, 2592), dtype=np.float32)
    E[g["E_rows"], g["E_cols"]] = g["E_vals"]
    return g, E
