"""Minimal forward-only driver of the bench workload for ncu captures."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from sigman_release_b200 import scenes
from common import gpu_forward
VIEWS = [30, 37, 45, 53, 65, 85, 0, 8]
sc = scenes.body_gaussians(100_000, seed=0)
bwd = len(sys.argv) > 1 and sys.argv[1] == "bwd"
for it in range(3):
    out, t, _ = gpu_forward(sc, VIEWS, 512, 512, requires_grad=bwd)
    if bwd:
        (out[0].clamp(0, 1) - 0.5).abs().mean().backward()
torch.cuda.synchronize()
