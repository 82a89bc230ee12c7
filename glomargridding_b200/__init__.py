"""glomargridding_b200 - B200-native ellipse covariance + kriging hot path of GloMarGridding.

Keeps the `glomar_gridding` Python API for that path (same class names, constructor signatures,
attributes, warnings and errors) over hand-written sm_100a CUDA kernels reached through the C ABI in
`include/glomar_b200.h`.  `install_as_glomar_gridding()` aliases this package under the reference's
import paths so that existing scripts run unchanged.
"""

from __future__ import annotations

import sys

__version__ = "0.1.0"

_SUBMODULES = ["ellipse", "ellipse.covariance", "ellipse.model", "ellipse.estimate", "grid", "kriging", "stochastic", "utils", "constants", "types",
               "distances", "variogram", "covariance_tools", "batched", "distributed"]


def install_as_glomar_gridding(name: str = "glomar_gridding") -> None:
    """Make `import glomar_gridding.ellipse` / `.kriging` / `.stochastic` resolve to this package."""
    import importlib

    me = sys.modules[__name__]
    sys.modules[name] = me
    for sub in _SUBMODULES:
        try:
            sys.modules[f"{name}.{sub}"] = importlib.import_module(f"{__name__}.{sub}")
        except ModuleNotFoundError:
            pass
