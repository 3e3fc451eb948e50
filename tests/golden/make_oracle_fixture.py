"""Generates tests/golden/oracle_scene_v1.npz: one small seeded scene rendered by the CPU oracle (fp32 spec) — the
regression fixture of the oracle's own outputs (images, radii, depth-ordered lists, gradients).  It does NOT pin the
oracle to the reference (the reference's rasteriser source is absent, see oracle/sgr_oracle.cpp); it pins the spec
the CUDA path is held to, so that neither side can drift silently.

    python tests/golden/make_oracle_fixture.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))
import oracle
from scene_utils import small_scene
from sigman_release_b200 import cameras

H, W, VIEW, N, SEED = 48, 64, 37, 400, 21


def scene():
    sc = small_scene(n=N, seed=SEED, spread=0.3, smin=0.005, smax=0.07, omax=0.95)
    return {k: np.asarray(v, np.float32) for k, v in sc.items()}


def render(sc):
    vm, pm, _ = cameras.orbit_cameras([VIEW])
    tan = cameras.tan_half_fov()
    r = oracle.Rasterizer(np.float32)
    o = r.forward(sc["means3D"], sc["cov3D"], sc["colors"], sc["opacities"], vm[0].reshape(-1), pm[0].reshape(-1), tan, tan,
                  (1.0, 0.5, 0.25), H, W)
    g = np.random.default_rng(SEED).normal(size=(3, H, W)).astype(np.float32)
    grads = r.backward(g)
    b = r.binning()
    return o, b, g, grads


def render64(sc, g):
    """fp64 instantiation of the same formulas: the reference the gradient-accuracy bound is measured against."""
    vm, pm, _ = cameras.orbit_cameras([VIEW])
    tan = cameras.tan_half_fov()
    r = oracle.Rasterizer(np.float64)
    r.forward(sc["means3D"], sc["cov3D"], sc["colors"], sc["opacities"], vm[0].reshape(-1), pm[0].reshape(-1), tan, tan,
              (1.0, 0.5, 0.25), H, W)
    return r.backward(g.astype(np.float64))


if __name__ == "__main__":
    sc = scene()
    o, b, g, grads = render(sc)
    grads64 = render64(sc, g)
    np.savez_compressed(os.path.join(HERE, "oracle_scene_v1.npz"), color=o.color, depth=o.depth, alpha=o.alpha, radii=o.radii,
                        point_list=b["point_list"], ranges=b["ranges"], n_contrib=b["n_contrib"],
                        **{"grad_" + k: v for k, v in grads.items()}, **{"grad64_" + k: v for k, v in grads64.items()})
    print("wrote oracle_scene_v1.npz:", o.num_instances, "instances,", o.blends, "blends")
