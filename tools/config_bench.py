"""Timings of the BASELINE.json configs other than the bench.py workload (configs[1]): config 1 (10K random Gaussians,
256x256, forward only), config 3 (8 subjects x 4 views, 512x512, gradients through the fused prep;
batched renderer vs the reference-shaped double loop over the drop-in module), config 4 (90-view orbit, sharded over
the ranks of the job + all-gather) and the config-5 stand-in (render-loss driver, 8 subjects x 10 views).
Prints one JSON object; run under torchrun for the multi-GPU numbers of configs 4 / 5."""
import json
import os
import sys
import time
from types import SimpleNamespace

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

from sigman_release_b200 import (GaussianRasterizationSettings, GaussianRasterizer, GaussianRenderer, cameras, rasterizer,
                                 scenes)
from sigman_release_b200.orbit import render_orbit_sharded
from sigman_release_b200.train_driver import RenderLossTrainer

world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0"))
local_rank = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local_rank)
dev = torch.device("cuda", local_rank)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
f32 = lambda a: torch.as_tensor(np.ascontiguousarray(a), dtype=torch.float32, device=dev)
TAN = cameras.tan_half_fov()
out = {"n_gpus": world}


def timed(fn, reps, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    b, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    b.record()
    for _ in range(reps):
        fn()
    e.record()
    torch.cuda.synchronize()
    ms = torch.tensor([b.elapsed_time(e) / reps], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    return float(ms.item())


only = set(sys.argv[1:]) or {"1", "3", "4", "5"}
# ---------------------------------------------------------------------------------------------- config 1
if "1" in only and rank == 0:
    # (bit-exactness of this configuration against the CPU oracle is a test: tests/test_gpu_parity.py::
    # test_config1_10k_random_256; the oracle itself is test infrastructure and is not used here)
    sc = scenes.random_gaussians(10_000, seed=0)
    vm, pm, _ = cameras.orbit_cameras([30])
    t = dict(m=f32(sc["means3D"])[None], c=f32(sc["cov3D"])[None], col=f32(sc["colors"])[None], o=f32(sc["opacities"])[None])
    bg = torch.ones(3, device=dev)
    run = lambda: rasterizer.rasterize_batch(t["m"], t["c"], t["col"], t["o"], f32(vm)[None], f32(pm)[None], bg, 256, 256, TAN, TAN)
    ms = timed(run, 200)
    out["config1"] = {"gpu_ms": ms, "gaussians_per_sec": 10_000 / (ms * 1e-3)}

# ---------------------------------------------------------------------------------------------- config 3
if "3" in only and rank == 0:
    B, V, N, H = 8, 4, 100_000, 512
    rng = np.random.default_rng(3)
    bodies = [scenes.body_gaussians(N, seed=10 + b, jitter=1.0) for b in range(B)]
    mk = lambda a: f32(np.stack(a)).requires_grad_(True)
    g = dict(position=mk([b["means3D"] for b in bodies]), opacity=mk([b["opacities"][:, None] for b in bodies]),
             scale=mk([rng.uniform(-1, 1, (N, 3)).astype(np.float32) for _ in bodies]),
             cov3d=mk([b["rotmats"] for b in bodies]), rgb=mk([b["colors"] for b in bodies]))
    vm, pm, cp = cameras.orbit_cameras([30, 37, 45, 53])
    cam_view = f32(vm)[None].repeat(B, 1, 1, 1); cam_vp = f32(pm)[None].repeat(B, 1, 1, 1); cam_pos = f32(cp)[None].repeat(B, 1, 1)
    renderer = GaussianRenderer(SimpleNamespace(output_size_h=H, output_size_w=H, FoVy=cameras.FOVY))
    target = torch.rand((B, V, 3, H, H), device=dev)

    def batched():
        for v in g.values():
            v.grad = None
        o = renderer.render(g, cam_view, cam_vp, cam_pos)
        (o["image"] - target).abs().mean().backward()

    def fused():
        for v in g.values():
            v.grad = None
        m3, c3, col, op = renderer.prepare(g)
        rasterizer.render_l1_loss(m3, c3, col, op, cam_view, cam_vp, renderer.bg_color, H, H, TAN, TAN, target)[0].backward()

    def looped():
        for v in g.values():
            v.grad = None
        m3, c3, col, op = renderer.prepare(g)
        imgs = []
        for b in range(B):
            for v in range(V):
                s = GaussianRasterizationSettings(image_height=H, image_width=H, tanfovx=TAN, tanfovy=TAN, bg=renderer.bg_color,
                                                  scale_modifier=0.5, viewmatrix=cam_view[b, v], projmatrix=cam_vp[b, v],
                                                  sh_degree=0, campos=cam_pos[b, v], prefiltered=False, debug=False)
                imgs.append(GaussianRasterizer(s)(means3D=m3[b], means2D=torch.zeros_like(m3[b]), shs=None,
                                                  colors_precomp=col[b], opacities=op[b], cov3D_precomp=c3[b])[0].clamp(0, 1))
        (torch.stack(imgs).view(B, V, 3, H, H) - target).abs().mean().backward()

    out["config3"] = {"renders": B * V, "batched_ms": timed(batched, 10), "batched_fused_loss_ms": timed(fused, 10),
                      "reference_shaped_loop_over_dropin_ms": timed(looped, 5)}
    out["config3"]["gaussians_per_sec_fused"] = B * V * N / (out["config3"]["batched_fused_loss_ms"] * 1e-3)
    del g, bodies

# ---------------------------------------------------------------------------------------------- config 4
if "4" in only:
    sc = scenes.body_gaussians(100_000, seed=0)
    t = {k: f32(sc[k])[None] for k in ("means3D", "cov3D", "colors")}
    t["opacities"] = f32(sc["opacities"]).reshape(1, -1)
    bg = torch.ones(3, device=dev)

    def render(views):
        vm, pm, _ = cameras.orbit_cameras(list(views))
        c, r, d, a = rasterizer.rasterize_batch(t["means3D"], t["cov3D"], t["colors"], t["opacities"], f32(vm)[None],
                                                f32(pm)[None], bg, 512, 512, TAN, TAN, clamp_color=True)
        return torch.cat([c[0], d[0], a[0]], dim=1)
    ms = timed(lambda: render_orbit_sharded(render, 90), 10)
    if rank == 0:
        out["config4"] = {"views": 90, "ms": ms, "views_per_sec": 90 / (ms * 1e-3), "gaussians_per_sec": 90 * 100_000 / (ms * 1e-3),
                          "gathered_bytes": 90 * 5 * 512 * 512 * 4}

# ---------------------------------------------------------------------------------------------- config 5 stand-in
if "5" in only:
    tr = RenderLossTrainer(8, 10, 100_000, 512, dev, seed=rank, ddp=world > 1)
    ms = timed(lambda: tr.step(), 10, warm=3)
    if rank == 0:
        out["config5_standin"] = {"subjects_per_gpu": 8, "views": 10, "ms_per_step": ms,
                                  "images_per_sec": 8 * 10 * world / (ms * 1e-3)}
if rank == 0:
    print(json.dumps(out))
if world > 1:
    dist.destroy_process_group()
