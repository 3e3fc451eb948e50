"""Multi-GPU orbit rendering: independent views shard across ranks, results are all-gathered
(BASELINE.json config 4; mirrors ``self.accelerator.gather(out['images_pred'])`` of
``/root/reference/core/loss/eval.py:81-82``).

One process per GPU; ``torch.distributed`` (NCCL over NVLink on the GPU box, gloo in the CPU tests) is only plumbing —
the renders themselves need no collective.
"""
from __future__ import annotations

from typing import Callable, Sequence

import torch
import torch.distributed as dist

__all__ = ["shard_views", "shard_range", "render_orbit_sharded", "render_orbit_overlapped", "OrbitRenderer", "to_wire",
           "from_wire"]


def shard_views(num_views: int, rank: int, world: int) -> list[int]:
    """Round-robin: rank r renders views r, r + world, r + 2*world, ..."""
    if not 0 <= rank < world:
        raise ValueError("rank outside [0, world)")
    return list(range(rank, num_views, world))


def to_wire(x: torch.Tensor, wire_dtype) -> torch.Tensor:
    """Wire format of the gather: fp32 (exact), fp16, or uint8 for images in [0, 1] (round to nearest
    of x * 255) — 2x / 4x fewer NVLink bytes for evaluation renders."""
    if wire_dtype in (None, torch.float32):
        return x
    if wire_dtype == torch.uint8:
        return (x.clamp(0, 1) * 255.0 + 0.5).to(torch.uint8)
    return x.to(wire_dtype)


def from_wire(x: torch.Tensor, wire_dtype) -> torch.Tensor:
    if wire_dtype in (None, torch.float32):
        return x
    if wire_dtype == torch.uint8:
        return x.to(torch.float32) / 255.0
    return x.to(torch.float32)


def render_orbit_sharded(render_fn: Callable[[Sequence[int]], torch.Tensor], num_views: int,
                         group=None, wire_dtype=None) -> torch.Tensor:
    """``render_fn(view_indices) -> [len(view_indices), C, H, W]`` is called with this rank's shard; returns the
    full ``[num_views, C, H, W]`` stack on every rank in view order.  Shards are padded to equal length so a single
    ``all_gather_into_tensor`` moves everything.  ``wire_dtype`` (fp16 / uint8) shrinks the gathered bytes; the result
    is converted back to fp32 (uint8 only for planes in [0, 1], e.g. clamped RGB + alpha)."""
    if not dist.is_available() or not dist.is_initialized():
        return from_wire(to_wire(render_fn(list(range(num_views))), wire_dtype), wire_dtype)
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    mine = shard_views(num_views, rank, world)
    per = (num_views + world - 1) // world
    padded = mine + [mine[-1] if mine else 0] * (per - len(mine))
    local = render_fn(padded)
    if local.shape[0] != per:
        raise ValueError("render_fn must return one image stack per requested view")
    local = to_wire(local, wire_dtype).contiguous()
    out = torch.empty((world * per,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, local, group=group)
    out = out.view(world, per, *local.shape[1:])
    # view v lives at [v % world, v // world]
    idx_r = torch.arange(num_views, device=local.device) % world
    idx_k = torch.arange(num_views, device=local.device) // world
    return from_wire(out[idx_r, idx_k], wire_dtype)


def shard_range(num_views: int, rank: int, world: int) -> tuple[int, int, int]:
    """Contiguous shards: rank r renders views [r * per, min((r + 1) * per, num_views)); returns (first, count, per).
    With this split a gather of the per-rank stacks IS the view-ordered stack (no permutation)."""
    if not 0 <= rank < world:
        raise ValueError("rank outside [0, world)")
    per = (num_views + world - 1) // world
    first = min(rank * per, num_views)
    return first, max(0, min(num_views - first, per)), per


# wire formats of the overlapped gather: (RGB, depth, alpha) element types
WIRE_EXACT = (torch.float32, torch.float32, torch.float32)       # bitwise equal to a single-GPU render
WIRE_COMPACT = (torch.uint8, torch.float16, torch.float16)       # 7 instead of 20 bytes per pixel (evaluation renders)


def render_orbit_overlapped(render_planes: Callable, num_views: int, height: int, width: int, device, group=None,
                            wire=WIRE_EXACT, chunk: int = 0) -> torch.Tensor:
    """Orbit render of BASELINE config 4 with the gather overlapped with the rendering (mirrors
    ``self.accelerator.gather(out['images_pred'])``, /root/reference/core/loss/eval.py:81-82).

    ``render_planes(view_ids, out) -> (color [n,3,H,W], depth [n,1,H,W], alpha [n,1,H,W])`` renders the given views;
    ``out`` (a tuple of three contiguous float32 tensors of those shapes, or None) is where it should write — they
    are slices of this rank's gather buffer, so with the exact wire format nothing is copied (no ``torch.cat``).
    Views are sharded contiguously; every rank renders its shard in chunks of ``chunk`` views (0 = a third of the
    shard, at least 6: below that a chunk's launch sequence costs the host more than the GPU needs to render it), and while chunk k + 1 renders on the current stream, chunk k is all-gathered (as bytes, one
    collective per chunk) and unpacked into the view-ordered result on a side stream — packing and unpacking are one
    kernel each of the CUDA library (``sgr_wire_pack`` / ``sgr_wire_unpack``).  Returns ``[num_views, 5, H, W]``
    float32 (RGB, depth, alpha) on every rank; with ``wire=WIRE_EXACT`` it is bitwise equal to a single-process
    render, ``WIRE_COMPACT`` sends uint8 RGB (colours in [0, 1]) and fp16 depth / alpha.  CPU tensors (the gloo tests
    of the host logic) take the same route with torch ops instead of the two kernels."""
    distributed = dist.is_available() and dist.is_initialized()
    world = dist.get_world_size(group) if distributed else 1
    rank = dist.get_rank(group) if distributed else 0
    first, count, per = shard_range(num_views, rank, world)
    device = torch.device(device)
    P = height * width
    planes = (3, 1, 1)
    wire = tuple(wire)
    if wire not in (WIRE_EXACT, WIRE_COMPACT):
        raise ValueError("wire must be WIRE_EXACT or WIRE_COMPACT")
    exact = wire == WIRE_EXACT
    sizes = [torch.empty((), dtype=dt).element_size() for dt in wire]
    on_cuda = device.type == "cuda"
    if chunk <= 0:
        chunk = max(6, (per + 2) // 3)
    if on_cuda:
        import ctypes

        from . import _native
        L = _native.lib()
        if P % 4:
            raise ValueError("height * width must be a multiple of 4")
    final = torch.empty((world * per, 5, height, width), dtype=torch.float32, device=device)
    final5 = final.view(world, per, 5, height, width)
    side = torch.cuda.Stream(device=device) if on_cuda else None
    keep = []                                    # buffers stay alive until the side stream has consumed them
    for k0 in range(0, per, chunk):
        c = min(chunk, per - k0)
        # this rank's views of the chunk; shards are padded (with the last valid view) to equal length
        ids = [min(first + k0 + i, max(first + count - 1, 0), num_views - 1) for i in range(c)]
        nbytes = [c * ch * P * es for ch, es in zip(planes, sizes)]
        send = torch.empty((sum(nbytes),), dtype=torch.uint8, device=device)
        offs = [0, nbytes[0], nbytes[0] + nbytes[1]]
        if exact:                                # the renderer writes straight into the send buffer
            outs = tuple(send[o:o + nb].view(torch.float32).view(c, ch, height, width)
                         for o, nb, ch in zip(offs, nbytes, planes))
            got = render_planes(ids, outs)
            for t, o in zip(got, outs):          # a renderer that ignored `out` (CPU tests): copy its result in
                if t.data_ptr() != o.data_ptr():
                    o.copy_(t)
        else:
            got = [t.contiguous() for t in render_planes(ids, None)]
            if on_cuda:
                with torch.cuda.device(device):
                    st = torch.cuda.current_stream(device)
                    _native.check(L.sgr_wire_pack(ctypes.c_void_p(got[0].data_ptr()), ctypes.c_void_p(got[1].data_ptr()),
                                                  ctypes.c_void_p(got[2].data_ptr()), c, P, ctypes.c_void_p(send.data_ptr()),
                                                  ctypes.c_void_p(st.cuda_stream)))
            else:
                for t, o, nb, dt in zip(got, offs, nbytes, wire):
                    send[o:o + nb].view(dt).copy_(to_wire(t, dt).reshape(-1))
        recv = torch.empty((world, send.numel()), dtype=torch.uint8, device=device) if distributed else send.view(1, -1)
        if on_cuda:
            ready = torch.cuda.Event()
            ready.record()
            side.wait_event(ready)
        ctx = torch.cuda.stream(side) if on_cuda else _NullCtx()
        with ctx:
            if distributed:
                dist.all_gather_into_tensor(recv.view(-1), send, group=group)
            if on_cuda:
                with torch.cuda.device(device):
                    _native.check(L.sgr_wire_unpack(ctypes.c_void_p(recv.data_ptr()), world, send.numel(), c, P,
                                                    0 if exact else 1, ctypes.c_void_p(final.data_ptr()), per, k0,
                                                    ctypes.c_void_p(side.cuda_stream)))
            else:
                for pi, (o, nb, ch, dt) in enumerate(zip(offs, nbytes, planes, wire)):
                    blk = recv[:, o:o + nb].view(dt).view(world, c, ch, height, width)
                    c0 = (0, 3, 4)[pi]
                    final5[:, k0:k0 + c, c0:c0 + ch] = from_wire(blk, dt)
        keep.append((send, recv, got))
    if on_cuda:
        torch.cuda.current_stream(device).wait_stream(side)
        for bufs in keep:                        # the caching allocator must not hand them out before the side stream is done
            for t in bufs[:2]:
                t.record_stream(side)
    return final[:num_views]


class _NullCtx:
    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False


def rasterizer_planes(means3D, cov3D, colors, opacities, bg, height, width, tanfov, view_matrices, proj_matrices):
    """``render_planes`` for ``render_orbit_overlapped`` over this package's rasteriser: one subject ([1,N,..] tensors),
    cameras ``view_matrices`` / ``proj_matrices`` [num_views,4,4] (device tensors); RGB is clamped to [0, 1] like
    gs.py:107.  The kernels write straight into ``out`` when it is given."""
    from .rasterizer import rasterize_batch

    def render(view_ids, out):
        ids = list(view_ids)
        if ids == list(range(ids[0], ids[0] + len(ids))):      # a contiguous run: plain slices, no index upload
            vm, pm = view_matrices[ids[0]:ids[0] + len(ids)][None], proj_matrices[ids[0]:ids[0] + len(ids)][None]
        else:                                                  # (padded tail of the last shard)
            idx = torch.as_tensor(ids, device=means3D.device)
            vm, pm = view_matrices[idx][None], proj_matrices[idx][None]
        o = None if out is None else tuple(t.unsqueeze(0) for t in out)
        c, _r, d, a = rasterize_batch(means3D, cov3D, colors, opacities, vm, pm, bg, height, width, tanfov, tanfov,
                                      clamp_color=True, out=o)
        return c[0], d[0], a[0]

    return render


class OrbitRenderer:
    """The overlapped orbit render of ``render_orbit_overlapped`` for a FIXED problem (one subject's tensors, a fixed
    camera orbit, image size, wire format): buffers are allocated once and — on CUDA — the render (+ pack) of every
    chunk is captured in a CUDA graph after one eager warm-up orbit, so that a later ``render()`` costs the host one
    graph launch, one all-gather and one unpack launch per chunk instead of a chunk's whole launch sequence (with
    eight ranks on one host the eager orbit is host-bound).  The subject's tensors are static: refresh them in place
    (``means3D.copy_(...)``) between calls.  Every rank must construct and call it alike (collectives inside).

    ``render()`` returns ``[num_views, 5, H, W]`` float32 (RGB, depth, alpha) on every rank — the same tensor object
    every call; bitwise equal to a single-process render with ``wire=WIRE_EXACT``."""

    def __init__(self, means3D, cov3D, colors, opacities, bg, height, width, tanfov, view_matrices, proj_matrices,
                 group=None, wire=WIRE_EXACT, chunk: int = 0, use_graphs: bool = True):
        import ctypes

        from . import _native
        self._ct, self._native = ctypes, _native
        self.L = _native.lib()
        self.t = (means3D, cov3D, colors, opacities, bg)
        self.H, self.W, self.tanfov = int(height), int(width), float(tanfov)
        self.group = group
        self.wire = tuple(wire)
        if self.wire not in (WIRE_EXACT, WIRE_COMPACT):
            raise ValueError("wire must be WIRE_EXACT or WIRE_COMPACT")
        self.exact = self.wire == WIRE_EXACT
        dev = means3D.device
        if dev.type != "cuda":
            raise ValueError("OrbitRenderer needs CUDA tensors (use render_orbit_overlapped for the CPU test path)")
        self.dev = dev
        self.distributed = dist.is_available() and dist.is_initialized()
        self.world = dist.get_world_size(group) if self.distributed else 1
        self.rank = dist.get_rank(group) if self.distributed else 0
        self.num_views = int(view_matrices.shape[0])
        first, count, per = shard_range(self.num_views, self.rank, self.world)
        self.per = per
        P = self.H * self.W
        if P % 4:
            raise ValueError("height * width must be a multiple of 4")
        if chunk <= 0:
            chunk = max(6, (per + 2) // 3)
        sizes = [torch.empty((), dtype=dt).element_size() for dt in self.wire]
        planes = (3, 1, 1)
        self.final = torch.empty((self.world * per, 5, self.H, self.W), dtype=torch.float32, device=dev)
        self.side = torch.cuda.Stream(device=dev)
        self.chunks = []
        for k0 in range(0, per, chunk):
            c = min(chunk, per - k0)
            ids = [min(first + k0 + i, max(first + count - 1, 0), self.num_views - 1) for i in range(c)]
            idx = torch.as_tensor(ids, device=dev)
            nbytes = [c * ch * P * es for ch, es in zip(planes, sizes)]
            send = torch.empty((sum(nbytes),), dtype=torch.uint8, device=dev)
            offs = [0, nbytes[0], nbytes[0] + nbytes[1]]
            if self.exact:                       # the renderer writes straight into the send buffer
                outs = tuple(send[o:o + nb].view(torch.float32).view(1, c, ch, self.H, self.W)
                             for o, nb, ch in zip(offs, nbytes, planes))
            else:                                # float32 planes, packed into the send buffer by sgr_wire_pack
                outs = tuple(torch.empty((1, c, ch, self.H, self.W), dtype=torch.float32, device=dev) for ch in planes)
            recv = (torch.empty((self.world, send.numel()), dtype=torch.uint8, device=dev) if self.distributed
                    else send.view(1, -1))
            self.chunks.append(dict(k0=k0, c=c, vm=view_matrices[idx][None].contiguous(),
                                    pm=proj_matrices[idx][None].contiguous(), send=send, outs=outs, recv=recv, graph=None))
        self.use_graphs = bool(use_graphs)
        self._warm = False

    def _render_chunk(self, ch):
        """Enqueues the render (+ pack) of one chunk on the current stream."""
        from .rasterizer import rasterize_batch
        m, c6, col, op, bg = self.t
        rasterize_batch(m, c6, col, op, ch["vm"], ch["pm"], bg, self.H, self.W, self.tanfov, self.tanfov,
                        clamp_color=True, out=ch["outs"])
        if not self.exact:
            ct = self._ct
            st = torch.cuda.current_stream(self.dev)
            o = ch["outs"]
            self._native.check(self.L.sgr_wire_pack(ct.c_void_p(o[0].data_ptr()), ct.c_void_p(o[1].data_ptr()),
                                                    ct.c_void_p(o[2].data_ptr()), ch["c"], self.H * self.W,
                                                    ct.c_void_p(ch["send"].data_ptr()), ct.c_void_p(st.cuda_stream)))

    def _gather_chunk(self, ch):
        """On the side stream: all-gather the chunk and scatter it into the view-ordered stack."""
        ct = self._ct
        ready = torch.cuda.Event()
        ready.record()
        self.side.wait_event(ready)
        with torch.cuda.stream(self.side):
            if self.distributed:
                dist.all_gather_into_tensor(ch["recv"].view(-1), ch["send"], group=self.group)
            self._native.check(self.L.sgr_wire_unpack(ct.c_void_p(ch["recv"].data_ptr()), self.world, ch["send"].numel(),
                                                      ch["c"], self.H * self.W, 0 if self.exact else 1,
                                                      ct.c_void_p(self.final.data_ptr()), self.per, ch["k0"],
                                                      ct.c_void_p(self.side.cuda_stream)))

    @torch.no_grad()
    def render(self) -> torch.Tensor:
        with torch.cuda.device(self.dev):
            cur = torch.cuda.current_stream(self.dev)
            self.side.wait_stream(cur)               # the previous result may still be read on the caller's stream
            if not self._warm:
                # eager warm-up orbit (sizes the instance buffers), then one graph per chunk
                for ch in self.chunks:
                    self._render_chunk(ch)
                    self._gather_chunk(ch)
                cur.wait_stream(self.side)
                self._warm = True
                if self.use_graphs:
                    torch.cuda.synchronize(self.dev)
                    for ch in self.chunks:
                        g = torch.cuda.CUDAGraph()
                        with torch.cuda.graph(g):
                            self._render_chunk(ch)
                        ch["graph"] = g
                    torch.cuda.synchronize(self.dev)
                return self.final[:self.num_views]
            for ch in self.chunks:
                if ch["graph"] is not None:
                    ch["graph"].replay()
                else:
                    self._render_chunk(ch)
                self._gather_chunk(ch)
            cur.wait_stream(self.side)
        return self.final[:self.num_views]
