"""Helpers shared by the ``-m gpu`` parity tests: run the CUDA path through the C ABI (via the Python shim) and the
CPU oracle on the same seeded inputs."""
import ctypes

import numpy as np
import torch

import oracle
from sigman_release_b200 import _native, cameras, rasterizer

TAN = cameras.tan_half_fov()


def to_dev(a, dtype=torch.float32):
    return torch.as_tensor(np.ascontiguousarray(a), dtype=dtype).cuda()


def scene_tensors(sc, requires_grad=False):
    """dict of numpy arrays [N,...] -> batched CUDA tensors with B = 1."""
    t = dict(means3D=to_dev(sc["means3D"])[None], cov3D=to_dev(sc["cov3D"])[None], colors=to_dev(sc["colors"])[None],
             opacities=to_dev(sc["opacities"]).reshape(1, -1))
    if requires_grad:
        for v in t.values():
            v.requires_grad_(True)
    return t


def view_tensors(view_ids):
    vm, pm, cp = cameras.orbit_cameras(view_ids)
    return to_dev(vm)[None], to_dev(pm)[None], vm, pm


def gpu_forward(sc, view_ids, H, W, bg=(1.0, 1.0, 1.0), simple=False, requires_grad=False, tensors=None, **kw):
    """One batched forward through the C ABI.  exact_exp defaults to True here: the parity tests compare bit for bit
    with the oracle's exp_spec; the default (SFU) exponential is covered by tests/test_gpu_fast_exp.py."""
    kw.setdefault("exact_exp", True)
    t = tensors if tensors is not None else scene_tensors(sc, requires_grad)
    vmt, pmt, vm, pm = view_tensors(view_ids)
    out = rasterizer.rasterize_batch(t["means3D"], t["cov3D"], t["colors"], t["opacities"], vmt, pmt,
                                     to_dev(np.asarray(bg, np.float32)), H, W, TAN, TAN, simple_blend=simple, **kw)
    return out, t, (vm, pm)


def oracle_forward(sc, vm, pm, H, W, bg=(1.0, 1.0, 1.0), dtype=np.float32):
    r = oracle.Rasterizer(dtype)
    out = r.forward(sc["means3D"], sc["cov3D"], sc["colors"], sc["opacities"], vm.reshape(-1), pm.reshape(-1), TAN, TAN,
                    bg, H, W)
    return r, out


def debug_state(fn_ctx_state, B, V, N, H, W, caps, render):
    """(tile_ranges [T,2], n_contrib [H,W], point_list) of one render from the saved state tensor of a forward;
    caps = dims[7] of the autograd node = (max_instances, max_block_records, flags)."""
    L = _native.lib()
    cap, cap_b, flags = caps
    T = ((W + 15) // 16) * ((H + 15) // 16)
    ranges = torch.zeros((T, 2), dtype=torch.int32, device="cuda")
    ncon = torch.zeros((H, W), dtype=torch.int32, device="cuda")
    cap_pl = max(int(cap), 1)
    pl = torch.full((cap_pl,), -1, dtype=torch.int32, device="cuda")
    st = torch.cuda.current_stream()
    _native.check(L.sgr_debug_copy_state(ctypes.c_void_p(fn_ctx_state.data_ptr()), B, V, N, H, W, int(cap), int(cap_b),
                                         int(flags), render,
                                         ctypes.c_void_p(ranges.data_ptr()), ctypes.c_void_p(ncon.data_ptr()),
                                         ctypes.c_void_p(pl.data_ptr()), cap_pl, None,
                                         ctypes.c_void_p(st.cuda_stream)))
    torch.cuda.synchronize()
    ranges = ranges.cpu().numpy().astype(np.uint32)
    n = int(ranges[:, 1].max()) if T else 0
    return ranges, ncon.cpu().numpy().astype(np.uint32), pl.cpu().numpy().astype(np.uint32)[:n]


def saved_state(color):
    """The autograd node of a rasterize_batch output -> (state tensor, dims)."""
    fn = color.grad_fn
    assert fn is not None, "forward must be run with requires_grad inputs to keep the state"
    state = fn.saved_tensors[-1]
    return state, fn.dims
