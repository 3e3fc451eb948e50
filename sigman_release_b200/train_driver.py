"""Render-loss step driver — the stand-in for BASELINE.json config 5 (``train_vae.py`` end to end).

``/root/reference/train_vae.py`` cannot run here or on the GPU box (no HGS-1M data, SMPL-X assets, diffusers, kiui,
lpips, accelerate, pytorch3d).  This driver keeps the *rasteriser-facing half* of one training step with the same
tensor interface as ``VAE.forward`` -> ``LPIPSWithDiscriminator`` minus the neural networks:

    learnable per-Gaussian features [B, N, 13]  (what ``decode_gaussian_*`` + ``grid_sample`` produce,
                                                 /root/reference/core/modules/autoencoder.py:292-302)
      -> a small shared affine head (the DDP-synchronised parameters; the reference synchronises its VAE weights)
      -> sigmoid / affine attribute maps of autoencoder.py:295-310, Rodrigues rotation on the template frame
         (autoencoder.py:333-337)
      -> GaussianRenderer: kNN scale + fused covariance prep + ONE batched render of B x V views
      -> L1(pred * mask, gt * mask), mean  (/root/reference/core/loss/whole_loss.py:126-130) fused into the blend
      -> backward -> AdamW(lr 3e-6, wd 0.05, betas (0.9, 0.95); train_vae.py:113) [-> DDP all-reduce of the head]

It is a stand-in and says so: no encoder, no LPIPS / discriminator / KL terms, synthetic targets.  Reported number:
images/s = B * V * ranks / step time.

    python -m sigman_release_b200.train_driver --subjects 8 --views 10 --steps 100
    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 -m sigman_release_b200.train_driver ...
"""
from __future__ import annotations

import argparse
import json
import math
import os
import time
from types import SimpleNamespace

import numpy as np
import torch
import torch.distributed as dist

from . import cameras, scenes
from .rasterizer import render_l1_loss
from .renderer import GaussianRenderer

TRAIN_VIEWS = [30, 37, 45, 53, 65, 85, 0, 8]     # /root/reference/core/dataset/dataloader_VAE.py:79
SIGMOID_SATURATION = 0.001                       # autoencoder.py:305-306


def batch_rodrigues(rot_vecs: torch.Tensor) -> torch.Tensor:
    """Axis-angle [n,3] -> rotation matrices [n,3,3] (the ``batch_rodrigues`` SIGMAN takes from SMPL-X)."""
    angle = torch.norm(rot_vecs + 1e-8, dim=1, keepdim=True)
    axis = rot_vecs / angle
    c, s = torch.cos(angle)[:, :, None], torch.sin(angle)[:, :, None]
    rx, ry, rz = axis[:, 0], axis[:, 1], axis[:, 2]
    z = torch.zeros_like(rx)
    K = torch.stack([z, -rz, ry, rz, z, -rx, -ry, rx, z], dim=1).view(-1, 3, 3)
    eye = torch.eye(3, dtype=rot_vecs.dtype, device=rot_vecs.device)[None]
    return eye + s * K + (1 - c) * torch.bmm(K, K)


def _times_skew(M: torch.Tensor, axis: torch.Tensor) -> torch.Tensor:
    """M @ [axis]_x for M [n,3,3], axis [n,3] with elementwise ops (a batched 3x3 bmm runs as 10^5..10^6 tiny SIMT
    GEMMs and would dominate the step)."""
    x, y, z = axis[:, 0:1], axis[:, 1:2], axis[:, 2:3]
    c0, c1, c2 = M[:, :, 0], M[:, :, 1], M[:, :, 2]
    return torch.stack([c1 * z - c2 * y, c2 * x - c0 * z, c0 * y - c1 * x], dim=2)


def compose_rotation(init_rot: torch.Tensor, rot_vecs: torch.Tensor) -> torch.Tensor:
    """``init_rot @ batch_rodrigues(rot_vecs)`` (autoencoder.py:333-334) without bmm:
    A (I + sin K + (1 - cos) K^2) = A + sin (A K) + (1 - cos) (A K) K."""
    angle = torch.norm(rot_vecs + 1e-8, dim=1, keepdim=True)
    axis = rot_vecs / angle
    AK = _times_skew(init_rot, axis)
    AKK = _times_skew(AK, axis)
    return init_rot + torch.sin(angle)[:, :, None] * AK + (1 - torch.cos(angle))[:, :, None] * AKK


class AttributeHead(torch.nn.Module):
    """The shared (DDP-synchronised) parameters of the stand-in: a per-channel affine map of the 13 features."""

    def __init__(self):
        super().__init__()
        self.weight = torch.nn.Parameter(torch.ones(13))
        self.bias = torch.nn.Parameter(torch.zeros(13))

    def forward(self, feats):
        return feats * self.weight + self.bias


def gaussians_from_features(feats: torch.Tensor, init_pcd: torch.Tensor, init_rot: torch.Tensor) -> dict:
    """autoencoder.py:295-310 + 333-337: features [B,N,13] -> the ``gaussians`` dict GaussianRenderer.render takes."""
    B, N = feats.shape[:2]
    opacity, offset, rgb, scale, rot = feats.split([1, 3, 3, 3, 3], dim=2)
    opacity, rgb, scale, rot = torch.sigmoid(opacity), torch.sigmoid(rgb), torch.sigmoid(scale), torch.sigmoid(rot)
    rgb = rgb * (1 + SIGMOID_SATURATION * 2) - SIGMOID_SATURATION
    scale = (scale - 0.5) * 2
    rot = (rot - 0.5) * math.pi
    R = compose_rotation(init_rot.reshape(-1, 3, 3), rot.reshape(-1, 3)).reshape(B, N, 3, 3)
    return {"position": init_pcd + offset, "opacity": opacity, "scale": scale, "cov3d": R, "rgb": rgb}


class RenderLossTrainer:
    def __init__(self, subjects: int, views: int, num_gaussians: int, size: int, device, seed: int = 0,
                 lr: float = 3e-6, weight_decay: float = 0.05, ddp: bool = False):
        self.B, self.V, self.N, self.H, self.W = subjects, views, num_gaussians, size, size
        self.device = device
        g = torch.Generator().manual_seed(seed)
        bodies = [scenes.body_gaussians(num_gaussians, seed=seed * 1000 + b, jitter=1.0 if b else 0.0) for b in range(subjects)]
        f32 = lambda a: torch.as_tensor(np.ascontiguousarray(a), dtype=torch.float32, device=device)
        self.init_pcd = torch.stack([f32(b["means3D"]) for b in bodies])
        self.init_rot = torch.stack([f32(b["rotmats"]) for b in bodies])
        feats = torch.randn((subjects, num_gaussians, 13), generator=g) * 0.5
        feats[:, :, 0] += 2.0                                     # mostly opaque, like a trained model
        feats[:, :, 1:4] *= 0.002                                 # offsets of a few millimetres
        self.feats = torch.nn.Parameter(feats.to(device))
        self.head = AttributeHead().to(device)
        self.model = self.head
        if ddp:
            self.model = torch.nn.parallel.DistributedDataParallel(self.head, device_ids=[device.index])
        self.opt = torch.optim.AdamW([self.feats] + list(self.head.parameters()), lr=lr, weight_decay=weight_decay,
                                     betas=(0.9, 0.95))
        ids = [TRAIN_VIEWS[v % len(TRAIN_VIEWS)] if v < len(TRAIN_VIEWS) else (11 * v) % 89 for v in range(views)]
        vm, pm, cp = cameras.orbit_cameras(ids)
        self.cam_view = f32(vm)[None].repeat(subjects, 1, 1, 1)
        self.cam_view_proj = f32(pm)[None].repeat(subjects, 1, 1, 1)
        self.cam_pos = f32(cp)[None].repeat(subjects, 1, 1)
        opt_ns = SimpleNamespace(output_size_h=size, output_size_w=size, FoVy=cameras.FOVY)
        self.renderer = GaussianRenderer(opt_ns)
        gd = torch.Generator(device=device).manual_seed(seed + 1)
        self.gt_images = torch.rand((subjects, views, 3, size, size), device=device, generator=gd)
        self.gt_masks = (torch.rand((subjects, views, 1, size, size), device=device, generator=gd) > 0.2).float()

    def loss(self) -> torch.Tensor:
        gaussians = gaussians_from_features(self.model(self.feats), self.init_pcd, self.init_rot)
        means3D, cov3D, rgbs, opacity = self.renderer.prepare(gaussians)
        tan = self.renderer.tan_half_fov
        return render_l1_loss(means3D, cov3D, rgbs, opacity, self.cam_view, self.cam_view_proj, self.renderer.bg_color,
                              self.H, self.W, tan, tan, self.gt_images, self.gt_masks)[0]

    def step(self) -> torch.Tensor:
        self.opt.zero_grad(set_to_none=True)
        loss = self.loss()
        loss.backward()
        self.opt.step()
        return loss.detach()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--subjects", type=int, default=8)           # batch_size 8 per rank (README / BASELINE config 5)
    ap.add_argument("--views", type=int, default=10)             # VAE.py:118 num_views
    ap.add_argument("--gaussians", type=int, default=100_000)
    ap.add_argument("--size", type=int, default=512)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--lr", type=float, default=3e-6)
    args = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("the render-loss driver needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    tr = RenderLossTrainer(args.subjects, args.views, args.gaussians, args.size, dev, seed=rank, lr=args.lr,
                           ddp=world > 1)
    first = None
    for _ in range(args.warmup):
        l = tr.step()
        first = l if first is None else first
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    beg, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    beg.record()
    for _ in range(args.steps):
        last = tr.step()
    end.record()
    torch.cuda.synchronize()
    ms = torch.tensor([beg.elapsed_time(end)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_step = float(ms.item()) / args.steps
    if rank == 0:
        print(json.dumps({
            "driver": "render-loss step (stand-in for train_vae.py; no encoder / LPIPS / GAN / KL)",
            "images_per_sec": args.subjects * args.views * world / (ms_step * 1e-3), "ms_per_step": ms_step,
            "n_gpus": world, "subjects_per_gpu": args.subjects, "views": args.views, "gaussians": args.gaussians,
            "image": [args.size, args.size], "steps": args.steps, "loss_first": float(first) if first is not None else None,
            "loss_last": float(last), "wall_s": time.perf_counter() - t0,
        }))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
