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


# SGR_FLAG_EXACT_EXP guarantee of the quarter-item blend kernels against the oracle: the transmittance chain, both
# thresholds and therefore n_contrib are bit-identical; colour / depth / alpha are four partial sums per pixel combined
# at the end, i.e. they differ from the oracle's single running sum by rounding only.  The upstream-shaped kernels
# (simple_blend) accumulate in the oracle's order and stay bitwise equal.
EXACT_MODE_ATOL = 2e-6
EXACT_MODE_RTOL = 2e-6


def assert_images(c, d, a, ora, bitwise):
    """colour [3,H,W], depth [1,H,W], alpha [1,H,W] (numpy) against an oracle ForwardResult."""
    if bitwise:
        np.testing.assert_array_equal(c, ora.color)
        np.testing.assert_array_equal(d, ora.depth)
        np.testing.assert_array_equal(a, ora.alpha)
    else:
        np.testing.assert_allclose(c, ora.color, atol=EXACT_MODE_ATOL, rtol=EXACT_MODE_RTOL)
        np.testing.assert_allclose(d, ora.depth, atol=EXACT_MODE_ATOL, rtol=EXACT_MODE_RTOL)
        np.testing.assert_allclose(a, ora.alpha, atol=EXACT_MODE_ATOL, rtol=EXACT_MODE_RTOL)


def grad_error_stats(got, want64):
    """(max abs error / max |g|, median and 99th percentile of the relative error over the elements with
    |g| > 1e-3 max|g|) of a gradient array against the fp64 oracle."""
    got = np.asarray(got, np.float64); want64 = np.asarray(want64, np.float64)
    s = np.abs(want64).max() + 1e-300
    e = np.abs(got - want64)
    big = np.abs(want64) > 1e-3 * s
    rel = e[big] / np.abs(want64[big]) if big.any() else np.zeros(1)
    return e.max() / s, float(np.percentile(rel, 50)), float(np.percentile(rel, 99))


def assert_grads_like_fp32(got, ref32, ref64, names=("means3D", "cov3D", "colors", "opacities"), slack=2.0, what=""):
    """Gradient parity (no global-only bound).  Per-Gaussian gradients are sums of strongly
    cancelling per-pixel terms, so the plain fp32 evaluation of the formulas (the fp32 oracle) is itself ~1e-3 of
    max|g| away from the fp64 result at config 2.  The CUDA path must be no less accurate than that evaluation, in the
    maximum AND in the distribution of per-element relative errors (elements down to 1e-3 of the largest gradient):
        max|err|       <= slack * max|err of the fp32 oracle| + 2e-5 * max|g|
        p50, p99(rel)  <= slack * the fp32 oracle's            + 1e-4
    all measured against the fp64 oracle.  got / ref32 / ref64: dicts of arrays."""
    for k in names:
        g = got[k].detach().cpu().numpy() if hasattr(got[k], "detach") else np.asarray(got[k])
        g = g.reshape(np.asarray(ref64[k]).shape)
        m_gpu, p50_gpu, p99_gpu = grad_error_stats(g, ref64[k])
        m_o32, p50_o32, p99_o32 = grad_error_stats(ref32[k], ref64[k])
        msg = (f"{what}{k}: gpu max {m_gpu:.2e} p50 {p50_gpu:.1e} p99 {p99_gpu:.1e} | "
               f"fp32 oracle max {m_o32:.2e} p50 {p50_o32:.1e} p99 {p99_o32:.1e}")
        assert m_gpu <= slack * m_o32 + 2e-5, msg
        assert p50_gpu <= slack * p50_o32 + 1e-4, msg
        assert p99_gpu <= slack * p99_o32 + 1e-4, msg


def oracle_grads(sc, vms, pms, H, W, gcs, gds=None, gas=None, bg=(1.0, 1.0, 1.0)):
    """Sum over the views of the fp32 and of the fp64 oracle's gradients for per-view image gradients gcs[v] (and
    optional depth / alpha gradients): returns (ref32, ref64, [fp32 forward results])."""
    ref32 = ref64 = None
    outs = []
    for v in range(len(vms)):
        r32, o32 = oracle_forward(sc, vms[v], pms[v], H, W, bg=bg)
        r64, _ = oracle_forward(sc, vms[v], pms[v], H, W, bg=bg, dtype=np.float64)
        gc = gcs[v](o32) if callable(gcs[v]) else gcs[v]
        gd = None if gds is None else gds[v]
        ga = None if gas is None else gas[v]
        g32 = r32.backward(gc, gd, ga)
        f64 = lambda a: None if a is None else np.asarray(a, np.float64)
        g64 = r64.backward(f64(gc), f64(gd), f64(ga))
        ref32 = g32 if ref32 is None else {k: ref32[k] + g32[k] for k in g32}
        ref64 = g64 if ref64 is None else {k: ref64[k] + g64[k] for k in g64}
        outs.append((r32, o32))
    return ref32, ref64, outs


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
