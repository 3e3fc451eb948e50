"""Helpers shared by the profiling / diagnostic drivers in tools/ (product code only: the CPU oracle is test
infrastructure and is not imported here)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from sigman_release_b200 import cameras, rasterizer

TAN = cameras.tan_half_fov()


def to_dev(a, dtype=torch.float32):
    return torch.as_tensor(np.ascontiguousarray(a), dtype=dtype).cuda()


def gpu_forward(sc, view_ids, H, W, bg=(1.0, 1.0, 1.0), requires_grad=False, **kw):
    """One batched forward of a scene dict (numpy arrays [N,...]) from the given orbit views; returns
    (outputs, input tensors, (viewmatrices, projmatrices))."""
    t = dict(means3D=to_dev(sc["means3D"])[None], cov3D=to_dev(sc["cov3D"])[None], colors=to_dev(sc["colors"])[None],
             opacities=to_dev(sc["opacities"]).reshape(1, -1))
    if requires_grad:
        for v in t.values():
            v.requires_grad_(True)
    vm, pm, _ = cameras.orbit_cameras(view_ids)
    out = rasterizer.rasterize_batch(t["means3D"], t["cov3D"], t["colors"], t["opacities"], to_dev(vm)[None],
                                     to_dev(pm)[None], to_dev(np.asarray(bg, np.float32)), H, W, TAN, TAN, **kw)
    return out, t, (vm, pm)


def saved_state(color):
    """The autograd node of a rasterize_batch output -> (state tensor, dims)."""
    fn = color.grad_fn
    assert fn is not None, "forward must be run with requires_grad inputs to keep the state"
    return fn.saved_tensors[-1], fn.dims
