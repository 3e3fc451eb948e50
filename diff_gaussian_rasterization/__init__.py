"""Import-compatible alias of the third-party ``diff_gaussian_rasterization`` package for
``/root/reference/core/gaussians/gs.py:8-11``: the two names SIGMAN imports, served by the sm_100a library."""
from sigman_release_b200.rasterizer import (  # noqa: F401
    GaussianRasterizationSettings,
    GaussianRasterizer,
    rasterize_gaussians,
)

__all__ = ["GaussianRasterizationSettings", "GaussianRasterizer", "rasterize_gaussians"]
