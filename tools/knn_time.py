import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from sigman_release_b200 import scenes, distCUDA2, prep_cov3d
from torch.profiler import profile, ProfilerActivity
sc = scenes.body_gaussians(100_000, seed=0)
p = torch.as_tensor(sc["means3D"]).cuda()
for _ in range(3): distCUDA2(p)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    d = distCUDA2(p)
    torch.cuda.synchronize()
for e in sorted([e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA], key=lambda e: e.time_range.start):
    print(f"{e.time_range.elapsed_us():8.1f} us  {e.name[:100]}")
