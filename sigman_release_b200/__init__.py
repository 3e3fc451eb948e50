"""sigman_release_b200 — B200-native (sm_100a) Gaussian-splat rasteriser behind the
``GaussianRasterizer`` / ``GaussianRasterizationSettings`` API used by
``/root/reference/core/gaussians/gs.py``.  See DESIGN.md."""
__version__ = "0.1.0"

from .rasterizer import (  # noqa: E402,F401
    GaussianRasterizationSettings,
    GaussianRasterizer,
    rasterize_batch,
    rasterize_gaussians,
    cov3d_from_scale_rot,
    last_status,
    check_status,
    graph_status,
    render_l1_loss,
    sh_colors,
)
from .graphs import GraphedStep  # noqa: E402,F401
from .renderer import GaussianRenderer, distCUDA2, get_covariance, prep_cov3d  # noqa: E402,F401
