"""Host-side profile of the reference-shaped per-view loop (gs.py:62-109) over the drop-in module: cProfile over 30
steps of 8 views (forward + backward)."""
import cProfile, pstats, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from sigman_release_b200 import GaussianRasterizationSettings, GaussianRasterizer, cameras, scenes
VIEWS = [30, 37, 45, 53, 65, 85, 0, 8]; H = W = 512
sc = scenes.body_gaussians(100_000, seed=0)
f32 = lambda a: torch.as_tensor(np.ascontiguousarray(a), dtype=torch.float32).cuda()
d = dict(means3D=f32(sc["means3D"]), cov3D=f32(sc["cov3D"]), colors=f32(sc["colors"]), opacities=f32(sc["opacities"])[:, None])
for v in d.values(): v.requires_grad_(True)
vm, pm, cp = cameras.orbit_cameras(VIEWS); vmt, pmt, cpt = f32(vm), f32(pm), f32(cp)
tan = cameras.tan_half_fov(); bg = torch.ones(3, device="cuda"); target = torch.rand((8, 3, H, W), device="cuda")
def step():
    for v in d.values(): v.grad = None
    imgs = []
    for v in range(8):
        st = GaussianRasterizationSettings(image_height=H, image_width=W, tanfovx=tan, tanfovy=tan, bg=bg, scale_modifier=0.5,
                                           viewmatrix=vmt[v], projmatrix=pmt[v], sh_degree=0, campos=cpt[v], prefiltered=False, debug=False)
        c = GaussianRasterizer(st)(means3D=d["means3D"], means2D=torch.zeros_like(d["means3D"]), shs=None,
                                   colors_precomp=d["colors"], opacities=d["opacities"], cov3D_precomp=d["cov3D"])[0]
        imgs.append(c.clamp(0, 1))
    (torch.stack(imgs) - target).abs().mean().backward()
for _ in range(5): step()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(30): step()
t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
print(f"enqueue {1e3*(t1-t0)/30:.3f} ms/step, wall {1e3*(t2-t0)/30:.3f} ms/step")
pr = cProfile.Profile(); pr.enable()
for _ in range(30): step()
pr.disable(); torch.cuda.synchronize()
pstats.Stats(pr).sort_stats("tottime").print_stats(24)
