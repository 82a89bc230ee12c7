"""Golden vector for BASELINE configs[3]: 1-degree grid, 20 000 observations, ordinary kriging - produced by the
UNMODIFIED reference classes on a REDUCED grid.

Kriging at a target cell depends only on the observations and on that target, so the reference does not need the
64 800 x 64 800 matrix (33.6 GB, and a (64 800 x 64 800) temporary inside `get_uncertainty`): the reduced grid is
{20 000 observed cells} U {64 unobserved target cells}.  Its 20 064 x 20 064 covariance comes from the reference's own
`c_ij_anisotropic_array` (ellipse/covariance.py:405-468) on the reference builder's own compressed (masked) state -
every cell outside the reduced set is masked, the builder's `_get_mask` (:149-180) does the compression - with the
pre-terms of `_calculate_covariance` (:218-224) and the mirrored scatter + diagonal of `calculate_covariance_array`
(:270-279, :246); the full pair list of `itertools.combinations` (2e8 tuples) is replaced by row blocks of
`np.triu_indices`-style index arrays, which is the only deviation from running the constructor as is.  The kriging is
`OrdinaryKriging(cov, idx, obs, error_cov)` -> `solve()`, `get_uncertainty()`, `constraint_mask()` (kriging.py:540-776)
untouched.

    python tests/golden/make_golden_1deg_krige.py        # ~15 min, ~25 GB of host memory, needs /root/reference

Output: tests/golden/krige_1deg_n20000.npz (cells, field, uncertainty, constraint mask at the 20 064 reduced cells).
"""

from __future__ import annotations

import contextlib
import io
import os
import sys
import time
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from glomargridding_b200 import synthetic as synth  # noqa: E402
from oracle.reference_shim import load_reference  # noqa: E402

ref = load_reference()
NU, METHOD, N_OBS, N_TARGETS, SEED = 1.5, "Met_Office", 20_000, 64, 4


class _PreTermsOnly(ref.EllipseCovarianceBuilder):
    def _calculate_covariance(self):
        cov_mod = ref.ellipse_covariance
        self._get_disp_fn()
        self.sigmas = cov_mod._sigma_rot_func_multi(
            self.Lx_compressed, self.Ly_compressed, self.theta_compressed).astype(self.precision)
        self.sqrt_dets = np.sqrt(cov_mod._det_22_multi(self.sigmas))
        self.gamma_v_term = cov_mod.gamma(self.v) * (2 ** (self.v - 1))
        self.sqrt_v_term = np.sqrt(self.v) * 2


def inputs():
    """Shared with tests/test_kriging_1deg_gpu.py: the cells of the reduced grid and the kriging inputs."""
    lats, lons = synth.global_grid(1.0)
    N = lats.size * lons.size
    idx, obs = synth.fixed_observations(SEED, N, N_OBS)
    rest = np.setdiff1d(np.arange(N), idx)
    targets = rest[np.linspace(0, rest.size - 1, N_TARGETS).astype(np.int64)]      # both poles, every latitude band
    cells = np.union1d(idx, targets)
    return lats, lons, idx, obs, cells


def main():
    t0 = time.time()
    lats, lons, idx, obs, cells = inputs()
    Lx, Ly, th, sd = synth.ellipse_fields(lats, lons)
    keep = np.zeros(lats.size * lons.size, dtype=bool)
    keep[cells] = True
    mask = ~keep.reshape(Lx.shape)
    with contextlib.redirect_stdout(io.StringIO()):
        b = _PreTermsOnly(*(np.ma.masked_where(mask, a) for a in (Lx, Ly, th, sd)), lats, lons, v=NU,
                          delta_x_method=METHOD, precision=np.float64)
    n = int(b.covar_size)
    assert n == cells.size
    cov = np.zeros((n, n))
    for r0 in range(0, n - 1, 256):
        r1 = min(n - 1, r0 + 256)
        i_s = np.repeat(np.arange(r0, r1), n - 1 - np.arange(r0, r1))
        j_s = np.concatenate([np.arange(i + 1, n) for i in range(r0, r1)])
        cov[i_s, j_s] = b.c_ij_anisotropic_array(i_s, j_s)
    cov = cov + cov.T
    cov[np.diag_indices(n)] += b.stdev_compressed ** 2
    print(f"covariance {n} x {n}: {time.time() - t0:.0f} s", flush=True)

    ridx = np.searchsorted(cells, idx)
    E = synth.hadcrut_style_error_cov(idx, lons.size)
    with contextlib.redirect_stdout(io.StringIO()), warnings.catch_warnings():
        warnings.simplefilter("ignore")
        ok = ref.OrdinaryKriging(cov, ridx, obs, E)
        field = ok.solve()
        print(f"solve: {time.time() - t0:.0f} s", file=sys.stderr, flush=True)
        unc = ok.get_uncertainty()
        print(f"uncertainty: {time.time() - t0:.0f} s", file=sys.stderr, flush=True)
        alpha = np.asarray(ok.kriging_weights[:, -1]).copy()
        del ok.kriging_weights
        msk = ok.constraint_mask()
    np.savez_compressed(os.path.join(HERE, "krige_1deg_n20000.npz"), cells=cells, field=field, uncertainty=unc,
                        constraint_mask=msk, alpha=alpha, nu=NU, n_obs=N_OBS)
    print(f"done: {time.time() - t0:.0f} s", flush=True)


if __name__ == "__main__":
    main()
