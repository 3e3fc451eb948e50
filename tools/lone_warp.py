"""Experiment: cycles per trip of a warp that has an SM (sub-partition) to itself — a few dense block lists only
(large, nearly transparent Gaussians stacked on one spot, 64x64 image = 128 block items on 148 SMs)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import numpy as np
import torch
from sigman_release_b200 import scenes
from common import gpu_forward

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2800
rng = np.random.default_rng(5)
xyz = rng.normal(scale=0.01, size=(n, 3))
rot = scenes.quat_to_rotmat(rng.normal(size=(n, 4)))
sc = dict(means3D=xyz, cov3D=scenes.covariance6(np.full((n, 3), 0.3), rot), colors=rng.uniform(0, 1, (n, 3)),
          opacities=np.full((n,), 0.02))
for hw in (64, 128):
    for it in range(3):
        gpu_forward(sc, [30], hw, hw)
    torch.cuda.synchronize()
    from sigman_release_b200 import _native
    import ctypes
    L = _native.lib()
    L.sgr_profile_enable(1)
    for it in range(5):
        gpu_forward(sc, [30], hw, hw)
    torch.cuda.synchronize()
    ms = (ctypes.c_double * len(_native.STAGES))(); cnt = (ctypes.c_uint32 * len(_native.STAGES))()
    L.sgr_profile_collect(ms, cnt)
    L.sgr_profile_enable(0)
    i = _native.STAGES.index("blend_forward")
    t = ms[i] / cnt[i] * 1e-3
    items = (hw // 16) ** 2 * 8
    print(f"{hw}x{hw}: {items} block items of {n} records (all survive): blend_forward {t * 1e6:.1f} us -> "
          f"{t * 1.9e9 / n:.0f} cycles per trip at 1.9 GHz (items per SM: {items / 148:.2f})")
