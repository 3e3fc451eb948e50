"""World-size-2 gloo test (CPU) of the multi-GPU orbit sharding: round-robin view shards + one all-gather give the
same stack as a single-process render."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from sigman_release_b200.orbit import render_orbit_sharded, shard_views


def _fake_render(views):
    # an "image" that encodes its view id: [len, 5, 4, 6]  (RGB + depth + alpha planes)
    return torch.stack([torch.full((5, 4, 6), float(v)) + torch.arange(5.0).view(5, 1, 1) * 0.01 for v in views])


def _worker(rank, world, port, num_views, ret, wire=None):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        fn = _fake_render if wire is None else (lambda v: _fake_render(v) / 100.0)
        out = render_orbit_sharded(fn, num_views, wire_dtype=wire)
        ret[rank] = out
    finally:
        dist.destroy_process_group()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def test_shard_views_partition():
    for world in (1, 2, 4, 8):
        for n in (1, 7, 90, 91):
            shards = [shard_views(n, r, world) for r in range(world)]
            assert sorted(sum(shards, [])) == list(range(n))
            assert max(len(s) for s in shards) - min(len(s) for s in shards) <= 1
    with pytest.raises(ValueError):
        shard_views(4, 2, 2)


@pytest.mark.parametrize("num_views", [7, 90])
def test_two_rank_gather_equals_single_process(num_views):
    world = 2
    ret = mp.Manager().dict()
    mp.spawn(_worker, args=(world, _free_port(), num_views, ret), nprocs=world, join=True)
    expect = _fake_render(range(num_views))
    for r in range(world):
        assert torch.equal(ret[r], expect)


def test_single_process_passthrough():
    assert torch.equal(render_orbit_sharded(_fake_render, 5), _fake_render(range(5)))


@pytest.mark.parametrize("wire,tol", [(torch.float16, 1e-3), (torch.uint8, 0.5 / 255 + 1e-6)])
def test_two_rank_gather_with_compact_wire_format(wire, tol):
    world, num_views = 2, 7
    ret = mp.Manager().dict()
    mp.spawn(_worker, args=(world, _free_port(), num_views, ret, wire), nprocs=world, join=True)
    want = _fake_render(list(range(num_views))) / 100.0
    for r in range(world):
        assert ret[r].dtype == torch.float32 and float((ret[r] - want).abs().max()) <= tol


# ---------------------------------------------------------------------------------------------- overlapped gather
def _fake_planes(ids, out):
    """(colour, depth, alpha) that encode the view id; colours in [0, 1] (uint8 wire format)."""
    c = torch.stack([torch.full((3, 4, 6), float(v) / 100.0) + torch.arange(3.0).view(3, 1, 1) * 0.001 for v in ids])
    d = torch.stack([torch.full((1, 4, 6), float(v) * 0.5) for v in ids])
    a = torch.stack([torch.full((1, 4, 6), float(v) / 200.0) for v in ids])
    return c, d, a


def _worker_overlapped(rank, world, port, num_views, chunk, compact, ret):
    from sigman_release_b200.orbit import WIRE_COMPACT, WIRE_EXACT, render_orbit_overlapped
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        ret[rank] = render_orbit_overlapped(_fake_planes, num_views, 4, 6, "cpu", wire=WIRE_COMPACT if compact else WIRE_EXACT,
                                            chunk=chunk)
    finally:
        dist.destroy_process_group()


def test_shard_range_partition():
    from sigman_release_b200.orbit import shard_range
    for world in (1, 2, 4, 8):
        for n in (1, 7, 90, 91):
            got = []
            for r in range(world):
                first, count, per = shard_range(n, r, world)
                assert per == (n + world - 1) // world and count <= per
                got += list(range(first, first + count))
            assert got == list(range(n))                       # contiguous shards in rank order = view order


@pytest.mark.parametrize("num_views,chunk", [(7, 2), (90, 4), (91, 16)])
def test_two_rank_overlapped_gather_equals_single_process(num_views, chunk):
    world = 2
    ret = mp.Manager().dict()
    mp.spawn(_worker_overlapped, args=(world, _free_port(), num_views, chunk, False, ret), nprocs=world, join=True)
    c, d, a = _fake_planes(range(num_views), None)
    expect = torch.cat([c, d, a], dim=1)
    for r in range(world):
        assert torch.equal(ret[r], expect)                     # exact wire format: bitwise


def test_two_rank_overlapped_gather_compact_wire():
    world, num_views = 2, 9
    ret = mp.Manager().dict()
    mp.spawn(_worker_overlapped, args=(world, _free_port(), num_views, 2, True, ret), nprocs=world, join=True)
    c, d, a = _fake_planes(range(num_views), None)
    for r in range(world):
        assert float((ret[r][:, 0:3] - c).abs().max()) <= 0.5 / 255 + 1e-6      # uint8 RGB
        assert float((ret[r][:, 3:4] - d).abs().max()) <= 4e-3                   # fp16 depth
        assert float((ret[r][:, 4:5] - a).abs().max()) <= 1e-4                   # fp16 alpha
