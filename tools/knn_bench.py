"""distCUDA2 timing (CUDA events, mean of 20 runs): one subject and 8 subjects batched, 100 K points each."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from sigman_release_b200 import scenes
from sigman_release_b200.renderer import distCUDA2, distCUDA2_batched
pts = torch.stack([torch.as_tensor(scenes.body_gaussians(100_000, seed=s)["means3D"]) for s in range(8)]).float().cuda()
for name, fn in (("1 x 100K", lambda: distCUDA2(pts[0])), ("8 x 100K batched", lambda: distCUDA2_batched(pts))):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    b, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    b.record()
    for _ in range(20): fn()
    e.record(); torch.cuda.synchronize()
    print(f"{name}: {b.elapsed_time(e) / 20 * 1e3:.1f} us per call (all kNN kernels)")
