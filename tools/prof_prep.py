"""Minimal driver of the per-subject preparation (gs.py:69-73) for ncu captures: batched kNN (distCUDA2) and the fused
scale / covariance kernel with its backward, 8 subjects x 100 K Gaussians (BASELINE config 3 shape)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from sigman_release_b200 import scenes
from sigman_release_b200.renderer import distCUDA2_batched, prep_cov3d

B, N = 8, 100_000
pts = torch.stack([torch.as_tensor(scenes.body_gaussians(N, seed=s)["means3D"]) for s in range(B)]).float().cuda()
g = torch.Generator(device="cuda").manual_seed(0)
scale_raw = (torch.rand((B, N, 3), device="cuda", generator=g) * 0.5).requires_grad_(True)
rot = torch.randn((B, N, 3, 3), device="cuda", generator=g).requires_grad_(True)
for it in range(3):
    d2 = distCUDA2_batched(pts)
    cov = prep_cov3d(scale_raw, rot, d2)
    cov.square().sum().backward()
torch.cuda.synchronize()
