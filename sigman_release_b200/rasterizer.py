"""Drop-in for the third-party ``diff_gaussian_rasterization`` Python API used by
``/root/reference/core/gaussians/gs.py:8-11,82-106``, backed by the
sm_100a C-ABI library ``libsgr_b200.so``.

* ``GaussianRasterizationSettings`` — NamedTuple with upstream's 12 fields in upstream's order.
* ``GaussianRasterizer(raster_settings)(means3D, means2D, opacities, shs=None, colors_precomp=None, scales=None,
  rotations=None, cov3D_precomp=None)`` -> ``(color [3,H,W], radii [N] int32, depth [1,H,W], alpha [1,H,W])``.
* ``rasterize_batch`` — the B200-native entry point: B subjects x V views in ONE launch set (replaces the Python
  double loop of gs.py:62-109).

PyTorch is used for device memory, streams and autograd plumbing only; all arithmetic runs in the CUDA library.
"""
from __future__ import annotations

import ctypes
import os
from typing import NamedTuple, Optional

import torch

from . import _native

__all__ = ["GaussianRasterizationSettings", "GaussianRasterizer", "rasterize_batch", "render_l1_loss", "rasterize_gaussians",
           "cov3d_from_scale_rot", "sh_colors", "last_status", "check_status", "graph_status"]


class GaussianRasterizationSettings(NamedTuple):
    image_height: int
    image_width: int
    tanfovx: float
    tanfovy: float
    bg: torch.Tensor
    scale_modifier: float
    viewmatrix: torch.Tensor
    projmatrix: torch.Tensor
    sh_degree: int
    campos: torch.Tensor
    prefiltered: bool
    debug: bool


# ------------------------------------------------------------------------------------------------ workspace policy
# The number of (Gaussian, tile) instances is data dependent and counted on the device.  The shim sizes the instance
# arrays from a per-shape estimate (1.5x the largest requirement seen).  How an overflow of that estimate is caught:
#   * the first call for a shape, every call of the per-view ``GaussianRasterizer`` module (the zero-change drop-in:
#     upstream reads ``num_rendered`` back per view as well) and every call in ``SGR_OVERFLOW_CHECK=sync`` mode read
#     the device status back and grow + retry INSIDE the call: their outputs are never wrong;
#   * batched calls in the default ``deferred`` mode fetch the status with an asynchronous 64-byte copy.  A forward
#     that a backward follows is verified at the END of that backward's launch sequence (the host waits for the
#     forward's status copy while the GPU already works on the queued backward kernels: no idle time), so an
#     overflowed step raises ``SgrError`` out of ``loss.backward()`` — before its gradients can be consumed;
#   * a forward-only (no-grad) deferred call cannot be repaired after the fact: the overflow is reported with a
#     ``RuntimeWarning`` by the next rasteriser call and raised by ``check_status()``; it never raises out of an
#     unrelated later forward.
_OVERFLOW_MODE = os.environ.get("SGR_OVERFLOW_CHECK", "deferred")      # "sync" | "deferred"
_HEADROOM = 1.5
_est_per_render: dict = {}
_max_tile_hint: dict = {}      # (N, H, W) -> longest per-tile list observed (sizes the long-list sort's shared memory)
_last_status: Optional[dict] = None
_pending: list = []            # deferred status checks, in launch order
_unreported: list = []         # overflow messages of forward-only calls, raised by check_status()


class _Pending:
    """One deferred status copy: pinned 64-byte buffer + the event recorded behind the copy."""
    __slots__ = ("ev", "pinned", "key", "R", "has_backward", "done", "status")

    def __init__(self, ev, pinned, key, R, has_backward):
        self.ev, self.pinned, self.key, self.R, self.has_backward = ev, pinned, key, R, has_backward
        self.done, self.status = False, None


def last_status() -> Optional[dict]:
    """Status block of the most recent checked forward (instances, longest tile list, ...)."""
    return _last_status


_STATUS_FIELDS = ("instances_required", "instances_capacity", "overflow", "max_tile_instances", "nonempty_tiles",
                  "block_records_required", "block_records_capacity")


def _status_from_bytes(t: torch.Tensor) -> dict:
    st = _native.SgrStatus.from_buffer_copy(bytes(t.numpy().tobytes()[: ctypes.sizeof(_native.SgrStatus)]))
    return {f: int(getattr(st, f)) for f in _STATUS_FIELDS}


def _note_status(key, R, st: dict) -> dict:
    """Folds a status block into the per-shape capacity estimates (instances and block records per render)."""
    _max_tile_hint[key[1:]] = max(_max_tile_hint.get(key[1:], 0), st["max_tile_instances"])
    per = int(st["instances_required"] / max(R, 1) * _HEADROOM) + 1024
    per_blk = int(st["block_records_required"] / max(R, 1) * _HEADROOM) + 2048
    old = _est_per_render.get(key, (0, 0))
    _est_per_render[key] = (max(old[0], per), max(old[1], per_blk))
    return st


_status_pool: list = []        # recycled (pinned 64-byte buffer, event) pairs of the deferred status copies


def _status_slot():
    if _status_pool:
        return _status_pool.pop()
    return torch.empty((64,), dtype=torch.uint8, pin_memory=True), torch.cuda.Event()


def _overflow_message(st: dict) -> str:
    return (f"the render overflowed its instance capacity ({st['instances_required']} (Gaussian, tile) instances "
            f"needed, {st['instances_capacity']} available; {st['block_records_required']} block records needed, "
            f"{st['block_records_capacity']} available): the dropped renders / tiles are background-only.  The "
            "estimate has been raised; re-run the step (or set SGR_OVERFLOW_CHECK=sync).")


def _complete(entry: "_Pending") -> dict:
    """Waits for the entry's status copy, folds it into the per-shape estimates and recycles its buffers."""
    global _last_status
    if entry.done:
        return entry.status
    entry.ev.synchronize()
    st = _status_from_bytes(entry.pinned)
    _status_pool.append((entry.pinned, entry.ev))
    entry.done, entry.status, entry.pinned, entry.ev = True, st, None, None
    _last_status = _note_status(entry.key, entry.R, st)
    return st


def _drain_pending(block: bool) -> None:
    """Polls (or waits for) the deferred status copies in launch order.  Never raises: an overflowed forward with a
    backward is reported by that backward (``_verify``), a forward-only one by a warning here and by check_status()."""
    while _pending:
        entry = _pending[0]
        if not entry.done and not block and not entry.ev.query():
            break                  # copies complete in launch order: everything behind stays unpolled
        _pending.pop(0)
        if entry.done:
            continue
        st = _complete(entry)
        if st["overflow"] and not entry.has_backward:
            msg = "a forward-only render: " + _overflow_message(st)
            _unreported.append(msg)
            import warnings
            warnings.warn(msg, RuntimeWarning, stacklevel=3)


def _verify(entry: Optional["_Pending"]) -> None:
    """Same-step verification of a deferred forward (called at the end of its backward): raises on overflow."""
    if entry is None:
        return
    st = _complete(entry)
    if st["overflow"]:
        raise _native.SgrError(_native.SGR_E_INSTANCE_OVERFLOW, "this step's forward: " + _overflow_message(st) +
                               "  The gradients of this step must be discarded.")


_graph_status: dict = {}       # device index -> persistent pinned status buffer of graph-captured forwards


def _graph_status_buffer(dev) -> torch.Tensor:
    buf = _graph_status.get(dev.index)
    if buf is None:
        buf = torch.zeros((64,), dtype=torch.uint8).pin_memory()
        _graph_status[dev.index] = buf
    return buf


def graph_status(device=None) -> Optional[dict]:
    """Status block written by the most recent replay of a CUDA graph that captured a forward on ``device`` (call after
    synchronising the replay stream); raises ``SgrError`` if that replay ran out of instance capacity — re-run the
    step eagerly once (the estimate grows) and capture again."""
    idx = torch.cuda.current_device() if device is None else torch.device(device).index
    buf = _graph_status.get(idx)
    if buf is None:
        return None
    st = _status_from_bytes(buf)
    if st["overflow"]:
        raise _native.SgrError(_native.SGR_E_INSTANCE_OVERFLOW,
                               f"a graph replay overflowed its instance capacity ({st['instances_required']} instances "
                               f"needed, {st['instances_capacity']} available; {st['block_records_required']} block "
                               f"records needed, {st['block_records_capacity']} available); re-run eagerly and re-capture")
    return st


def check_status() -> Optional[dict]:
    """Waits for every deferred overflow check (``SGR_OVERFLOW_CHECK=deferred``) and raises ``SgrError`` if a
    forward-only render since the last check ran out of instance capacity (forwards with a backward are verified by
    their own backward); returns the latest status block.  Call it after a batch of no-grad renders (eval orbits)."""
    _drain_pending(block=True)
    if _unreported:
        msg = _unreported[0] + (f" (+{len(_unreported) - 1} more)" if len(_unreported) > 1 else "")
        _unreported.clear()
        raise _native.SgrError(_native.SGR_E_INSTANCE_OVERFLOW, msg)
    return _last_status


_bytes_cache: dict = {}


def _buffer_bytes(L, B, V, N, H, W, cap, cap_b, flags, rpc):
    """(state bytes, scratch bytes) of a problem shape (cached: the launch path is host-latency sensitive)."""
    simple = flags & _native.FLAG_SIMPLE_BLEND
    key = (B, V, N, H, W, cap, cap_b, simple, rpc)
    got = _bytes_cache.get(key)
    if got is None:
        if len(_bytes_cache) > 256:
            _bytes_cache.clear()
        got = (int(L.sgr_state_bytes(B, V, N, H, W, cap, cap_b, simple)),
               int(L.sgr_scratch_bytes(B, V, N, H, W, cap, rpc)))
        _bytes_cache[key] = got
    return got


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def _check_input(name: str, t: torch.Tensor, shape) -> torch.Tensor:
    if not isinstance(t, torch.Tensor):
        raise TypeError(f"{name} must be a torch.Tensor")
    if not t.is_cuda:
        raise ValueError(f"{name} must be a CUDA tensor (sigman_release_b200 has no CPU path)")
    if t.dtype != torch.float32:
        raise ValueError(f"{name} must be float32, got {t.dtype}")
    if tuple(t.shape) != tuple(shape):
        raise ValueError(f"{name} must have shape {tuple(shape)}, got {tuple(t.shape)}")
    return t.contiguous()


def _fill_problem(p: _native.SgrProblem, B, V, N, H, W, tanfovx, tanfovy, means3D, cov3D, colors, opacities, viewm,
                  projm, bg, cap, cap_b, flags, rpc):
    p.num_subjects, p.views_per_subject, p.num_gaussians = B, V, N
    p.image_height, p.image_width = H, W
    p.tanfovx, p.tanfovy = float(tanfovx), float(tanfovy)
    p.means3D, p.cov3D, p.colors, p.opacities = _ptr(means3D), _ptr(cov3D), _ptr(colors), _ptr(opacities)
    p.viewmatrix, p.projmatrix, p.bg = _ptr(viewm), _ptr(projm), _ptr(bg)
    p.max_instances = cap
    p.max_block_records = cap_b
    p.renders_per_chunk = rpc
    p.flags = flags
    p.max_tile_instances_hint = _max_tile_hint.get((N, H, W), 0)


class _RasterizeBatch(torch.autograd.Function):
    """means3D [B,N,3], cov3D [B,N,6], colors [B,N,3], opacities [B,N], means2D [B,V,N,3] (gradient slot only).
    Returns (color, radii, depth, alpha, loss | None, lpips_feed | None)."""

    @staticmethod
    def forward(ctx, means3D, cov3D, colors, opacities, means2D, viewmatrix, projmatrix, bg, H, W, tanfovx, tanfovy,
                flags, renders_per_chunk, loss_target=None, loss_mask=None, loss_scale=None, force_sync=False,
                want_feed=False, out=None):
        global _last_status
        L = _native.lib()
        B, N = int(means3D.shape[0]), int(means3D.shape[1])
        V = int(viewmatrix.shape[1])
        R = B * V
        dev = means3D.device
        with torch.cuda.device(dev):
            stream = torch.cuda.current_stream(dev)
            if out is None:
                color = torch.empty((B, V, 3, H, W), dtype=torch.float32, device=dev)
                depth = torch.empty((B, V, 1, H, W), dtype=torch.float32, device=dev)
                alpha = torch.empty((B, V, 1, H, W), dtype=torch.float32, device=dev)
            else:                                # caller-owned output planes (e.g. slices of an all-gather buffer)
                color, depth, alpha = out
            radii = torch.empty((B, V, N), dtype=torch.int32, device=dev)
            fused = loss_target is not None
            loss = torch.empty((), dtype=torch.float32, device=dev) if fused else None
            g_fused = torch.empty((B, V, 3, H, W), dtype=torch.float32, device=dev) if fused else None
            feed = torch.empty((B, V, 3, H // 2, W // 2), dtype=torch.float32, device=dev) if want_feed else None
            key = (dev.index, N, H, W)
            pending = None
            capturing = torch.cuda.is_current_stream_capturing()
            if not capturing:
                _drain_pending(block=False)
                _graph_status_buffer(dev)                # pinned allocation must not happen inside a later capture
            first = key not in _est_per_render
            if capturing and first:
                raise RuntimeError("run the step at least once before capturing it in a CUDA graph: the first call for "
                                   "a problem shape sizes the instance buffers with a synchronous status read")
            per, per_blk = _est_per_render.get(key, (max(3 * N, 4096), max(6 * N, 8192)))
            sync_check = (first or force_sync or _OVERFLOW_MODE == "sync") and not capturing
            while True:
                cap = min(int(per) * R, (1 << 32) - 2)
                cap_b = min(int(per_blk) * R, (1 << 32) - 2)
                state_bytes, scratch_bytes = _buffer_bytes(L, B, V, N, H, W, cap, cap_b, flags, renders_per_chunk)
                state = torch.empty((state_bytes,), dtype=torch.uint8, device=dev)
                scratch = torch.empty((scratch_bytes,), dtype=torch.uint8, device=dev)
                a = _native.SgrForwardArgs()
                _fill_problem(a.p, B, V, N, H, W, tanfovx, tanfovy, means3D, cov3D, colors, opacities, viewmatrix,
                              projmatrix, bg, cap, cap_b, flags, renders_per_chunk)
                a.out_color, a.out_depth, a.out_alpha, a.radii = _ptr(color), _ptr(depth), _ptr(alpha), _ptr(radii)
                a.state, a.state_bytes = _ptr(state), state_bytes
                a.scratch, a.scratch_bytes = _ptr(scratch), scratch_bytes
                a.stream = ctypes.c_void_p(stream.cuda_stream)
                a.out_lpips_feed = _ptr(feed)
                if fused:
                    a.loss_target, a.loss_mask = _ptr(loss_target), _ptr(loss_mask)
                    a.loss_dL_dcolor, a.loss_out = _ptr(g_fused), _ptr(loss)
                    a.loss_scale = float(loss_scale)
                _native.check(L.sgr_forward(ctypes.byref(a)))
                if capturing:
                    # inside a CUDA graph capture: no events, no host reads.  The status block of every replay lands
                    # in a persistent pinned buffer that graph_status() decodes after the caller has synchronised.
                    _graph_status_buffer(dev).copy_(state[:64], non_blocking=True)
                    break
                if not sync_check:
                    pinned, ev = _status_slot()
                    pinned.copy_(state[:64], non_blocking=True)
                    ev.record(stream)
                    pending = _Pending(ev, pinned, key, R, has_backward=not (flags & _native.FLAG_FORWARD_ONLY))
                    _pending.append(pending)
                    break
                st = _native.SgrStatus()
                rc = L.sgr_read_status(_ptr(state), ctypes.c_void_p(stream.cuda_stream), ctypes.byref(st))
                if rc not in (_native.SGR_OK, _native.SGR_E_INSTANCE_OVERFLOW):
                    _native.check(rc)
                _last_status = _note_status(key, R, {f: int(getattr(st, f)) for f in _STATUS_FIELDS})
                if rc == _native.SGR_OK:
                    break
                if max(int(st.instances_required), int(st.block_records_required)) >= (1 << 32) - 2:
                    raise _native.SgrError(rc, "more than 2^32 (Gaussian, tile) instances / block records in one call; "
                                               "split the batch")
                per, per_blk = _est_per_render[key]
                del state, scratch
        ctx.save_for_backward(means3D, cov3D, colors, opacities, viewmatrix, projmatrix, bg, alpha, radii, state)
        ctx.dims = (B, V, N, H, W, float(tanfovx), float(tanfovy), (cap, cap_b, flags), flags, renders_per_chunk,
                    state_bytes)
        ctx.want_means2D = means2D is not None and means2D.requires_grad
        ctx.set_materialize_grads(False)     # unused outputs arrive as None, not as zero tensors
        ctx.g_fused = g_fused
        ctx.pending = pending                # deferred overflow check of this forward, verified by its backward
        ctx.mark_non_differentiable(radii)
        return color, radii, depth, alpha, loss, feed

    @staticmethod
    def backward(ctx, g_color, _g_radii, g_depth, g_alpha, g_loss, g_feed):
        if g_color is None and g_depth is None and g_alpha is None and g_feed is None and \
                (g_loss is None or ctx.g_fused is None):
            return (None,) * 20
        L = _native.lib()
        means3D, cov3D, colors, opacities, viewmatrix, projmatrix, bg, alpha, radii, state = ctx.saved_tensors
        B, V, N, H, W, tanfovx, tanfovy, caps, flags, rpc, state_bytes = ctx.dims
        cap, cap_b = caps[0], caps[1]
        dev = means3D.device
        # A deferred overflow check never blocks BEFORE the backward kernels are queued (a wait here would idle the GPU
        # for the whole launch latency of the backward); this forward's status is verified right after the launch.
        if not torch.cuda.is_current_stream_capturing():
            _drain_pending(block=False)
        with torch.cuda.device(dev):
            stream = torch.cuda.current_stream(dev)
            f32c = lambda t: None if t is None else t.detach().to(device=dev, dtype=torch.float32).contiguous()
            g_color, g_depth, g_alpha, g_feed = f32c(g_color), f32c(g_depth), f32c(g_alpha), f32c(g_feed)
            g_loss = f32c(g_loss) if ctx.g_fused is not None else None
            if flags & _native.FLAG_SIMPLE_BLEND and g_color is None:
                g_color = torch.zeros((B, V, 3, H, W), dtype=torch.float32, device=dev)
            d_means3D = torch.empty((B, N, 3), dtype=torch.float32, device=dev)
            d_cov3D = torch.empty((B, N, 6), dtype=torch.float32, device=dev)
            d_colors = torch.empty((B, N, 3), dtype=torch.float32, device=dev)
            d_opac = torch.empty((B, N), dtype=torch.float32, device=dev)
            d_means2D = torch.empty((B, V, N, 3), dtype=torch.float32, device=dev) if ctx.want_means2D else None
            scratch_bytes = _buffer_bytes(L, B, V, N, H, W, cap, cap_b, flags, rpc)[1]
            scratch = torch.empty((scratch_bytes,), dtype=torch.uint8, device=dev)
            a = _native.SgrBackwardArgs()
            _fill_problem(a.p, B, V, N, H, W, tanfovx, tanfovy, means3D, cov3D, colors, opacities, viewmatrix,
                          projmatrix, bg, cap, cap_b, flags, rpc)
            a.out_alpha, a.radii = _ptr(alpha), _ptr(radii)
            a.dL_dcolor, a.dL_ddepth, a.dL_dalpha = _ptr(g_color), _ptr(g_depth), _ptr(g_alpha)
            a.dL_dmeans3D, a.dL_dcov3D, a.dL_dcolors, a.dL_dopacities = (_ptr(d_means3D), _ptr(d_cov3D),
                                                                        _ptr(d_colors), _ptr(d_opac))
            a.dL_dmeans2D = _ptr(d_means2D)
            a.state, a.state_bytes = _ptr(state), state_bytes
            a.scratch, a.scratch_bytes = _ptr(scratch), scratch_bytes
            a.stream = ctypes.c_void_p(stream.cuda_stream)
            a.dL_dlpips_feed = _ptr(g_feed)
            if ctx.g_fused is not None:
                a.fused_clamp = 1
                if g_loss is not None:           # the fused loss's own gradient, scaled by the upstream scalar
                    a.loss_dL_dcolor, a.dL_dcolor_scale = _ptr(ctx.g_fused), _ptr(g_loss)
            _native.check(L.sgr_backward(ctypes.byref(a)))
        # the GPU now works on the queued backward while the host waits for the forward's 64-byte status copy: an
        # overflowed step raises here, out of loss.backward(), before anything can consume its gradients
        _verify(ctx.pending)
        return (d_means3D, d_cov3D, d_colors, d_opac, d_means2D) + (None,) * 15


def _checked(means3D, cov3D, colors, opacities, viewmatrix, projmatrix, bg):
    if means3D.dim() != 3 or means3D.shape[-1] != 3:
        raise ValueError("means3D must be [B,N,3]")
    B, N = int(means3D.shape[0]), int(means3D.shape[1])
    if viewmatrix.dim() != 4 or tuple(viewmatrix.shape[2:]) != (4, 4) or int(viewmatrix.shape[0]) != B:
        raise ValueError("viewmatrix must be [B,V,4,4]")
    V = int(viewmatrix.shape[1])
    if opacities.dim() == 3:
        opacities = opacities.reshape(B, N)
    return (B, N, V, _check_input("means3D", means3D, (B, N, 3)), _check_input("cov3D", cov3D, (B, N, 6)),
            _check_input("colors", colors, (B, N, 3)), _check_input("opacities", opacities, (B, N)),
            _check_input("viewmatrix", viewmatrix, (B, V, 4, 4)), _check_input("projmatrix", projmatrix, (B, V, 4, 4)),
            _check_input("bg", bg, (3,)))


def _common_flags(grad_inputs, exact_exp) -> int:
    flags = _native.FLAG_EXACT_EXP if (exact_exp or _EXACT_EXP_ENV) else 0
    if not (torch.is_grad_enabled() and any(t is not None and t.requires_grad for t in grad_inputs)):
        flags |= _native.FLAG_FORWARD_ONLY           # no backward can follow: skip the forward's bookkeeping for it
    if os.environ.get("SGR_TILE_TIMING"):            # diagnostics (tools/tile_timing.py)
        flags |= _native.FLAG_TILE_TIMING
    return flags


_EXACT_EXP_ENV = os.environ.get("SGR_EXACT_EXP", "0") not in ("", "0")


def rasterize_batch(means3D, cov3D, colors, opacities, viewmatrix, projmatrix, bg, image_height, image_width,
                    tanfovx, tanfovy, means2D=None, clamp_color=False, simple_blend=False, renders_per_chunk=0,
                    sync_if_no_grad=False, exact_exp=False, lpips_feed=False, out=None):
    """Render B subjects x V views in one launch set.

    means3D [B,N,3], cov3D [B,N,6] (xx,xy,xz,yy,yz,zz — gs.py:32-37), colors [B,N,3], opacities [B,N] or [B,N,1],
    viewmatrix / projmatrix [B,V,4,4] in the layout of gs.py:89-90 (``cam_view``, ``cam_view_proj``), bg [3].
    Returns ``(color [B,V,3,H,W], radii [B,V,N] int32, depth [B,V,1,H,W], alpha [B,V,1,H,W])``; differentiable
    w.r.t. means3D, cov3D, colors, opacities (gradients summed over a subject's views) and, if given, the per-render
    gradient slot ``means2D [B,V,N,3]``.

    * ``clamp_color`` fuses gs.py:107's ``clamp(0, 1)`` into the blend epilogue; the clamp's gradient mask is kept in
      the rasteriser state and applied inside the backward kernel (no clamp kernel, no saved image).
    * ``lpips_feed=True`` (needs ``clamp_color``; even H, W) appends a fifth output
      ``F.interpolate(color * 2 - 1, (H/2, W/2), mode='bilinear', align_corners=False)`` computed in the same epilogue
      (whole_loss.py:132-136), differentiable as well.
    * ``exact_exp=True`` evaluates ``exp(power)`` with the oracle's IEEE fp32 sequence (bit-exact against
      ``oracle/``); the default uses the SFU like upstream's own ``exp`` (see include/sgr.h SGR_FLAG_EXACT_EXP).
    * ``sync_if_no_grad``: a call that no backward can follow checks its instance capacity synchronously (grow and
      retry inside the call) instead of deferring the check — see the workspace policy at the top of this module.
    * ``out=(color, depth, alpha)``: caller-owned contiguous float32 CUDA tensors of the result shapes that the
      kernels write directly (no copy) — e.g. slices of an all-gather send buffer (``orbit.py``).
    """
    B, N, V, means3D, cov3D, colors, opacities, viewmatrix, projmatrix, bg = _checked(
        means3D, cov3D, colors, opacities, viewmatrix, projmatrix, bg)
    if means2D is not None and tuple(means2D.shape) != (B, V, N, 3):
        raise ValueError("means2D must be [B,V,N,3]")
    if lpips_feed and not clamp_color:
        raise ValueError("lpips_feed needs clamp_color=True (the reference resizes the clamped image)")
    flags = (_native.FLAG_CLAMP_COLOR if clamp_color else 0) | (_native.FLAG_SIMPLE_BLEND if simple_blend else 0)
    flags |= _common_flags((means3D, cov3D, colors, opacities, means2D), exact_exp)
    force_sync = bool(sync_if_no_grad) and bool(flags & _native.FLAG_FORWARD_ONLY)
    if out is not None:
        H_, W_ = int(image_height), int(image_width)
        for t, ch, name in zip(out, (3, 1, 1), ("color", "depth", "alpha")):
            if (not isinstance(t, torch.Tensor) or not t.is_cuda or t.dtype != torch.float32 or not t.is_contiguous()
                    or tuple(t.shape) != (B, V, ch, H_, W_)):
                raise ValueError(f"out[{name}] must be a contiguous float32 CUDA tensor of shape {(B, V, ch, H_, W_)}")
        out = tuple(out)
    color, radii, depth, alpha, _loss, feed = _RasterizeBatch.apply(
        means3D, cov3D, colors, opacities, means2D, viewmatrix, projmatrix, bg, int(image_height), int(image_width),
        float(tanfovx), float(tanfovy), flags, int(renders_per_chunk), None, None, None, force_sync, bool(lpips_feed),
        out)
    return (color, radii, depth, alpha, feed) if lpips_feed else (color, radii, depth, alpha)


def _loss_scale(reduction, B, V, H, W) -> float:
    if isinstance(reduction, (int, float)):
        return float(reduction)
    if reduction == "reference":         # torch.sum(l1(pred * m, gt * m)) / (B * V), whole_loss.py:130,139
        return 1.0 / float(B * V)
    if reduction == "mean":
        return 1.0 / float(B * V * 3 * H * W)
    if reduction == "sum":
        return 1.0
    raise ValueError("reduction must be 'reference', 'mean', 'sum' or a number (the factor applied to the sum)")


def render_l1_loss(means3D, cov3D, colors, opacities, viewmatrix, projmatrix, bg, image_height, image_width,
                   tanfovx, tanfovy, target, mask=None, renders_per_chunk=0, reduction="reference", exact_exp=False,
                   lpips_feed=False):
    """Render B x V views and evaluate SIGMAN's reconstruction loss in the blend epilogue.

    Equivalent to ``image = rasterize_batch(...)[0].clamp(0, 1)`` (gs.py:107) followed by the reference's L1 term:
    ``loss_l1 = l1(image * mask, target * mask)`` is an UNREDUCED ``|x - y|``
    (/root/reference/core/loss/whole_loss.py:49-50,130) that the reference reduces as
    ``torch.sum(loss_l1) / loss_l1.shape[0]`` with ``shape[0] = B * V`` (whole_loss.py:139) — a per-image sum
    averaged over the B*V images, NOT a mean over elements.  ``reduction``:

    * ``"reference"`` (default): ``sum / (B * V)``, the reference's weighting against its LPIPS / KL / GAN terms;
    * ``"mean"``: ``sum / (B * V * 3 * H * W)``;  ``"sum"``: the plain sum;  a number: that factor times the sum.

    The reference then divides by ``exp(logvar)`` (whole_loss.py:142): multiply the returned scalar by your device
    scalar — the product's gradient reaches the backward kernel as a device scalar (``dL_dcolor_scale``), no sync.

    The loss, the clamp mask and dL/dcolour never leave the blend kernel: no elementwise torch kernels and no saved
    image-sized tensors for the L1 term.  ``target`` [B,V,3,H,W], ``mask`` [B,V,1,H,W] or None.  Returns
    ``(loss, image [B,V,3,H,W] clamped, radii, depth, alpha)``.  ``loss`` is differentiable w.r.t. means3D, cov3D,
    colors, opacities; ``image`` (the clamped render), ``depth`` and ``alpha`` are differentiable as well, so further
    image-space terms (LPIPS on the resized image, the generator's GAN term — whole_loss.py:132-170) can be added on
    top: their gradient w.r.t. the clamped image is masked by the forward's clamp mask inside the backward kernel and
    added to the fused L1 gradient.  ``lpips_feed=True`` (even H, W) appends the LPIPS input of whole_loss.py:132-136,
    ``F.interpolate(image * 2 - 1, (H/2, W/2), mode='bilinear', align_corners=False)``, computed in the same epilogue
    and differentiable.  ``exact_exp``: see ``rasterize_batch``.
    """
    B, N, V, means3D, cov3D, colors, opacities, viewmatrix, projmatrix, bg = _checked(
        means3D, cov3D, colors, opacities, viewmatrix, projmatrix, bg)
    H, W = int(image_height), int(image_width)
    target = _check_input("target", target.detach(), (B, V, 3, H, W))
    if mask is not None:
        mask = _check_input("mask", mask.detach(), (B, V, 1, H, W))
    flags = _common_flags((means3D, cov3D, colors, opacities), exact_exp)
    color, radii, depth, alpha, loss, feed = _RasterizeBatch.apply(
        means3D, cov3D, colors, opacities, None, viewmatrix, projmatrix, bg, H, W, float(tanfovx), float(tanfovy),
        flags, int(renders_per_chunk), target, mask, _loss_scale(reduction, B, V, H, W), False, bool(lpips_feed), None)
    return (loss, color, radii, depth, alpha, feed) if lpips_feed else (loss, color, radii, depth, alpha)


# ------------------------------------------------------------------------------------------------ optional inputs
class _Cov3DFromScaleRot(torch.autograd.Function):
    """upstream computeCov3D (+ backward): Sigma = R diag((mod*s)^2) R^T from scales [N,3], quaternions [N,4]."""

    @staticmethod
    def forward(ctx, scales, rotations, scale_modifier):
        L = _native.lib()
        N = int(scales.shape[0])
        cov = torch.empty((N, 6), dtype=torch.float32, device=scales.device)
        with torch.cuda.device(scales.device):
            s = torch.cuda.current_stream(scales.device)
            _native.check(L.sgr_cov3d_from_scale_rot(_ptr(scales), _ptr(rotations), float(scale_modifier), N,
                                                     _ptr(cov), ctypes.c_void_p(s.cuda_stream)))
        ctx.save_for_backward(scales, rotations)
        ctx.mod = float(scale_modifier)
        return cov

    @staticmethod
    def backward(ctx, g_cov):
        L = _native.lib()
        scales, rotations = ctx.saved_tensors
        N = int(scales.shape[0])
        g_cov = g_cov.contiguous().float()
        ds = torch.empty_like(scales)
        dr = torch.empty_like(rotations)
        with torch.cuda.device(scales.device):
            s = torch.cuda.current_stream(scales.device)
            _native.check(L.sgr_cov3d_from_scale_rot_backward(_ptr(scales), _ptr(rotations), ctx.mod, N, _ptr(g_cov),
                                                              _ptr(ds), _ptr(dr), ctypes.c_void_p(s.cuda_stream)))
        return ds, dr, None


def cov3d_from_scale_rot(scales, rotations, scale_modifier=1.0):
    N = int(scales.shape[0])
    scales = _check_input("scales", scales, (N, 3))
    rotations = _check_input("rotations", rotations, (N, 4))
    return _Cov3DFromScaleRot.apply(scales, rotations, float(scale_modifier))


class _SHColors(torch.autograd.Function):
    """upstream computeColorFromSH (+ backward): colours [N,3] from shs [N,K,3] seen from ``campos``."""

    @staticmethod
    def forward(ctx, means3D, shs, campos, degree):
        L = _native.lib()
        N, K = int(shs.shape[0]), int(shs.shape[1])
        colors = torch.empty((N, 3), dtype=torch.float32, device=means3D.device)
        clamped = torch.empty((N, 3), dtype=torch.uint8, device=means3D.device)
        with torch.cuda.device(means3D.device):
            st = torch.cuda.current_stream(means3D.device)
            _native.check(L.sgr_sh_colors(_ptr(means3D), _ptr(shs), _ptr(campos), N, int(degree), K, _ptr(colors),
                                          _ptr(clamped), ctypes.c_void_p(st.cuda_stream)))
        ctx.save_for_backward(means3D, shs, campos, clamped)
        ctx.degree = int(degree)
        return colors

    @staticmethod
    def backward(ctx, g_colors):
        L = _native.lib()
        means3D, shs, campos, clamped = ctx.saved_tensors
        N, K = int(shs.shape[0]), int(shs.shape[1])
        g_colors = g_colors.contiguous().float()
        d_shs = torch.empty_like(shs)
        d_means = torch.empty_like(means3D)
        with torch.cuda.device(means3D.device):
            st = torch.cuda.current_stream(means3D.device)
            _native.check(L.sgr_sh_colors_backward(_ptr(means3D), _ptr(shs), _ptr(campos), N, ctx.degree, K,
                                                   _ptr(clamped), _ptr(g_colors), _ptr(d_shs), _ptr(d_means),
                                                   ctypes.c_void_p(st.cuda_stream)))
        return d_means, d_shs, None, None


def sh_colors(means3D, shs, campos, degree):
    """``max(0, SH_degree(normalize(means3D - campos)) . shs + 0.5)`` — the ``shs`` input path of the upstream API.
    means3D [N,3], shs [N,K,3] with K >= (degree+1)^2, campos [3]; differentiable w.r.t. means3D and shs."""
    N = int(means3D.shape[0])
    if shs.dim() != 3 or int(shs.shape[0]) != N or int(shs.shape[2]) != 3:
        raise ValueError("shs must be [N,K,3]")
    degree = int(degree)
    if not 0 <= degree <= 3 or int(shs.shape[1]) < (degree + 1) ** 2:
        raise ValueError("sh_degree must be 0..3 with at least (degree+1)^2 coefficients per Gaussian")
    means3D = _check_input("means3D", means3D, (N, 3))
    shs = _check_input("shs", shs, tuple(shs.shape))
    campos = _check_input("campos", campos.to(means3D.device, torch.float32).reshape(3), (3,))
    return _SHColors.apply(means3D, shs, campos, degree)


def rasterize_gaussians(means3D, means2D, sh, colors_precomp, opacities, scales, rotations, cov3Ds_precomp,
                        raster_settings: GaussianRasterizationSettings):
    """One view of one Gaussian set — upstream ``rasterize_gaussians`` (same argument order)."""
    N = int(means3D.shape[0])
    if sh is not None and sh.numel() > 0:       # upstream's SH colour path (SIGMAN passes colors_precomp, gs.py:102)
        colors_precomp = sh_colors(means3D, sh, raster_settings.campos, raster_settings.sh_degree)
    if cov3Ds_precomp is None or cov3Ds_precomp.numel() == 0:
        cov3D = cov3d_from_scale_rot(scales, rotations, raster_settings.scale_modifier)
    else:
        cov3D = cov3Ds_precomp
    s = raster_settings
    H, W = int(s.image_height), int(s.image_width)
    dev = means3D.device
    m2 = None
    if means2D is not None and means2D.requires_grad:
        m2 = means2D.reshape(1, 1, N, 3)
    color, radii, depth, alpha = rasterize_batch(
        means3D.reshape(1, N, 3), cov3D.reshape(1, N, 6), colors_precomp.reshape(1, N, 3), opacities.reshape(1, N),
        s.viewmatrix.to(dev, torch.float32).reshape(1, 1, 4, 4), s.projmatrix.to(dev, torch.float32).reshape(1, 1, 4, 4),
        s.bg.to(dev, torch.float32).reshape(3), H, W, float(s.tanfovx), float(s.tanfovy), means2D=m2,
        sync_if_no_grad=True)
    return color[0, 0], radii[0, 0], depth[0, 0], alpha[0, 0]


class GaussianRasterizer(torch.nn.Module):
    """Same constructor, ``forward`` keywords, return tuple and exceptions as upstream's module
    (call site gs.py:96-106)."""

    def __init__(self, raster_settings: GaussianRasterizationSettings):
        super().__init__()
        self.raster_settings = raster_settings

    def markVisible(self, positions: torch.Tensor) -> torch.Tensor:
        """upstream ``_C.mark_visible``: boolean [N], True where the view-space depth exceeds 0.2."""
        L = _native.lib()
        s = self.raster_settings
        with torch.no_grad():
            pos = positions.detach().contiguous().float()
            N = int(pos.shape[0])
            vis = torch.empty((N,), dtype=torch.uint8, device=pos.device)
            vm = s.viewmatrix.to(pos.device, torch.float32).contiguous()
            pm = s.projmatrix.to(pos.device, torch.float32).contiguous()
            with torch.cuda.device(pos.device):
                st = torch.cuda.current_stream(pos.device)
                _native.check(L.sgr_mark_visible(_ptr(pos), N, _ptr(vm), _ptr(pm), _ptr(vis),
                                                 ctypes.c_void_p(st.cuda_stream)))
        return vis.bool()

    def forward(self, means3D, means2D, opacities, shs=None, colors_precomp=None, scales=None, rotations=None,
                cov3D_precomp=None):
        if (shs is None and colors_precomp is None) or (shs is not None and colors_precomp is not None):
            raise Exception("Please provide excatly one of either SHs or precomputed colors!")
        if ((scales is None or rotations is None) and cov3D_precomp is None) or (
                (scales is not None or rotations is not None) and cov3D_precomp is not None):
            raise Exception("Please provide exactly one of either scale/rotation pair or precomputed 3D covariance!")
        return rasterize_gaussians(means3D, means2D, shs, colors_precomp, opacities, scales, rotations, cov3D_precomp,
                                   self.raster_settings)
