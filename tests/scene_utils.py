"""Small seeded scenes shared by the CPU and GPU tests."""
import numpy as np

from sigman_release_b200 import cameras, scenes


def small_scene(n=40, seed=0, spread=0.25, smin=0.01, smax=0.06, omax=0.9):
    """A handful of Gaussians in front of orbit view 30, sized to overlap on a small image."""
    rng = np.random.default_rng(seed)
    xyz = rng.uniform(-spread, spread, (n, 3))
    scale = rng.uniform(smin, smax, (n, 3))
    quat = rng.normal(size=(n, 4))
    rot = scenes.quat_to_rotmat(quat)
    return dict(means3D=xyz, cov3D=scenes.covariance6(scale, rot), colors=rng.uniform(0, 1, (n, 3)),
                opacities=rng.uniform(0.1, omax, (n,)))


def camera(view_id=30):
    vm, pm, cp = cameras.rasterizer_matrices(cameras.orbit_w2c(view_id))
    return vm.reshape(-1), pm.reshape(-1), cp


TAN = cameras.tan_half_fov()
