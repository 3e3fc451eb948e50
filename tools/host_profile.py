"""Host-side (enqueue) profile of the bench step: cProfile over 300 steps, GPU work left to run asynchronously."""
import cProfile, pstats, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from sigman_release_b200 import cameras, rasterizer, scenes
VIEWS = [30, 37, 45, 53, 65, 85, 0, 8]; H = W = 512
sc = scenes.body_gaussians(100_000, seed=0)
f32 = lambda a: torch.as_tensor(np.ascontiguousarray(a), dtype=torch.float32).cuda()
t = dict(means3D=f32(sc["means3D"])[None], cov3D=f32(sc["cov3D"])[None], colors=f32(sc["colors"])[None], opacities=f32(sc["opacities"])[None])
for v in t.values(): v.requires_grad_(True)
vm, pm, _ = cameras.orbit_cameras(VIEWS); vmt, pmt = f32(vm)[None], f32(pm)[None]
tan = cameras.tan_half_fov(); bg = torch.ones(3, device="cuda"); target = torch.rand((1, 8, 3, H, W), device="cuda")
def step():
    for v in t.values(): v.grad = None
    rasterizer.render_l1_loss(t["means3D"], t["cov3D"], t["colors"], t["opacities"], vmt, pmt, bg, H, W, tan, tan, target)[0].backward()
for _ in range(20): step()
torch.cuda.synchronize()
ts = []
for rep in range(10):                       # 20 steps from an empty queue: no launch-queue back-pressure
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(20): step()
    ts.append((time.perf_counter() - t0) / 20); torch.cuda.synchronize()
print(f"host enqueue without back-pressure: {1e3*min(ts):.3f} ms/step (min of 10 x 20 steps), median {1e3*sorted(ts)[5]:.3f}")
import ctypes
from sigman_release_b200 import _native
L = _native.lib()
t0 = time.perf_counter()
for _ in range(2000): L.sgr_abi_version()
print(f"ctypes call overhead {1e6*(time.perf_counter()-t0)/2000:.2f} us")
t0 = time.perf_counter()
for _ in range(2000): torch.empty((1000,), device="cuda")
print(f"torch.empty {1e6*(time.perf_counter()-t0)/2000:.2f} us")
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(300): step()
t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
print(f"enqueue {1e3*(t1-t0)/300:.3f} ms/step, wall {1e3*(t2-t0)/300:.3f} ms/step")
pr = cProfile.Profile(); pr.enable()
for _ in range(300): step()
pr.disable(); torch.cuda.synchronize()
pstats.Stats(pr).sort_stats("tottime").print_stats(22)
