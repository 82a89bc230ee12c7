"""Literal types of the hot-path keyword arguments (reference `glomar_gridding/types.py:46-48`)."""

from typing import Literal

DeltaXMethod = Literal["Met_Office", "Modified_Met_Office"]
CovarianceMethod = Literal["array", "low_memory", "batched"]
