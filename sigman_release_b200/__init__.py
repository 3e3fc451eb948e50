"""sigman_release_b200 — B200-native (sm_100a) Gaussian-splat rasteriser behind the
``GaussianRasterizer`` / ``GaussianRasterizationSettings`` API used by
``/root/reference/core/gaussians/gs.py``.  See DESIGN.md."""
__version__ = "0.1.0"
