"""Ellipse-based covariance (drop-in for `glomar_gridding.ellipse`, reference `ellipse/__init__.py:17-25`): the covariance
builder, the ellipse model (likelihood + Nelder-Mead fit) and the builder that fits every grid cell - all on the device."""

from .model import EllipseModel  # noqa: I001
from .estimate import EllipseBuilder
from .covariance import EllipseCovarianceBuilder

__all__ = ["EllipseBuilder", "EllipseCovarianceBuilder", "EllipseModel"]
