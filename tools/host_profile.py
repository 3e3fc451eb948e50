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
t0 = time.perf_counter()
for _ in range(300): step()
t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
print(f"enqueue {1e3*(t1-t0)/300:.3f} ms/step, wall {1e3*(t2-t0)/300:.3f} ms/step")
pr = cProfile.Profile(); pr.enable()
for _ in range(300): step()
pr.disable(); torch.cuda.synchronize()
pstats.Stats(pr).sort_stats("tottime").print_stats(22)
