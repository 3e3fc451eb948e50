"""Kernel timeline of a few bench steps (torch.profiler / CUPTI): prints every GPU activity of the last step with its
start offset, duration and the idle gap before it — used to find launch bubbles between the stages."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from torch.profiler import profile, ProfilerActivity
from sigman_release_b200 import cameras, rasterizer, scenes

VIEWS = [30, 37, 45, 53, 65, 85, 0, 8]
H = W = 512
sc = scenes.body_gaussians(100_000, seed=0)
f32 = lambda a: torch.as_tensor(np.ascontiguousarray(a), dtype=torch.float32).cuda()
t = dict(means3D=f32(sc["means3D"])[None], cov3D=f32(sc["cov3D"])[None], colors=f32(sc["colors"])[None], opacities=f32(sc["opacities"])[None])
for v in t.values():
    v.requires_grad_(True)
vm, pm, _ = cameras.orbit_cameras(VIEWS)
vmt, pmt = f32(vm)[None], f32(pm)[None]
tan = cameras.tan_half_fov()
bg = torch.ones(3, device="cuda")
target = torch.rand((1, len(VIEWS), 3, H, W), device="cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

def step():
    for v in t.values():
        v.grad = None
    loss = rasterizer.render_l1_loss(t["means3D"], t["cov3D"], t["colors"], t["opacities"], vmt, pmt, bg, H, W, tan, tan, target)[0]
    loss.backward()

if len(sys.argv) > 1 and sys.argv[1] == "driver":      # the config-5 stand-in step instead of the bench step
    from sigman_release_b200.train_driver import RenderLossTrainer
    tr = RenderLossTrainer(8, 10, 100_000, 512, torch.device("cuda", 0), seed=0)
    step = tr.step

for _ in range(5):
    step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for _ in range(4):
        flush.zero_()
        step()
    torch.cuda.synchronize()
ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
ev.sort(key=lambda e: e.time_range.start)
# last step = after the last big fill
fills = [i for i, e in enumerate(ev) if "FillFunctor<unsigned char>" in e.name or ("Memset" in e.name and e.time_range.elapsed_us() > 30)]
start = fills[-1]
t0 = ev[start].time_range.end
prev_end = t0
busy = 0.0
for e in ev[start + 1:]:
    s, d = e.time_range.start, e.time_range.elapsed_us()
    print(f"{s - t0:9.1f} us  dur {d:8.1f}  gap {s - prev_end:7.1f}  {e.name[:90]}")
    prev_end = max(prev_end, e.time_range.end)
    busy += d
print(f"span {prev_end - t0:.1f} us, sum of durations {busy:.1f} us")
