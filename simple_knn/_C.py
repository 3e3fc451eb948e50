"""``simple_knn._C.distCUDA2`` served by libsgr_b200.so (sgr_knn_mean_dist2)."""
from sigman_release_b200.renderer import distCUDA2  # noqa: F401

__all__ = ["distCUDA2"]
