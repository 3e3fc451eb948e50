"""Host / device time of the orbit render (BASELINE config 4) on one GPU for several chunk sizes:
`python tools/orbit_profile.py` (single process; the gather is a no-op, the unpack kernel still runs)."""
import cProfile, pstats, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from sigman_release_b200 import cameras, orbit, rasterizer, scenes

H = W = 512
dev = torch.device("cuda:0")
sc = scenes.body_gaussians(100_000, seed=0)
f32 = lambda a: torch.as_tensor(np.ascontiguousarray(a), dtype=torch.float32).to(dev)
t = [f32(sc[k])[None] for k in ("means3D", "cov3D", "colors")] + [f32(sc["opacities"]).reshape(1, -1)]
vm, pm, _ = cameras.orbit_cameras(list(range(90)))
bg = torch.ones(3, device=dev)
fn = orbit.rasterizer_planes(t[0], t[1], t[2], t[3], bg, H, W, cameras.tan_half_fov(), f32(vm), f32(pm))
views = int(sys.argv[1]) if len(sys.argv) > 1 else 90
with torch.no_grad():
    for chunk in (4, 6, 12, 30):
        for wire in (orbit.WIRE_EXACT, orbit.WIRE_COMPACT):
            for _ in range(3):
                orbit.render_orbit_overlapped(fn, views, H, W, dev, wire=wire, chunk=chunk)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t0 = time.perf_counter(); e0.record()
            for _ in range(5):
                orbit.render_orbit_overlapped(fn, views, H, W, dev, wire=wire, chunk=chunk)
            e1.record(); t1 = time.perf_counter(); torch.cuda.synchronize()
            print(f"views {views} chunk {chunk:2d} wire {'exact' if wire == orbit.WIRE_EXACT else 'compact'}: "
                  f"host enqueue {1e3 * (t1 - t0) / 5:.3f} ms, device {e0.elapsed_time(e1) / 5:.3f} ms")
    pr = cProfile.Profile(); pr.enable()
    for _ in range(20):
        orbit.render_orbit_overlapped(fn, views, H, W, dev, wire=orbit.WIRE_EXACT, chunk=4)
    pr.disable(); torch.cuda.synchronize()
    pstats.Stats(pr).sort_stats("tottime").print_stats(18)
rasterizer.check_status()
