"""Ellipse-based covariance (drop-in for `glomar_gridding.ellipse`, reference `ellipse/__init__.py:17-25`).

Only the covariance builder is on the accelerated path; the ellipse *fitting* classes
(`EllipseModel`, `EllipseBuilder`) are out of scope.
"""

from .covariance import EllipseCovarianceBuilder

__all__ = ["EllipseCovarianceBuilder"]
