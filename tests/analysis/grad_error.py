"""Gradient accuracy of the CUDA backward at BASELINE config 2 against the fp64 oracle, next to the accuracy of the
fp32 oracle itself (the straightforward fp32 evaluation of the same formulas): max abs error relative to max|g| and
percentiles of the per-element relative error.  Test infrastructure (uses oracle/)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from gpu_utils import gpu_forward, oracle_forward, to_dev
from sigman_release_b200 import scenes

views = [30, 65]
H = W = int(os.environ.get("HW", "512"))
N = int(os.environ.get("N", "100000"))
body = scenes.body_gaussians(N, seed=0)
rng = np.random.default_rng(0)
target = rng.uniform(0, 1, (len(views), 3, H, W)).astype(np.float32)
for exact in (True, False):
    out, t, (vm, pm) = gpu_forward(body, views, H, W, requires_grad=True, exact_exp=exact)
    color = out[0]
    loss = (color[0].clamp(0, 1) - to_dev(target)).abs().mean()
    loss.backward()
    ref32 = ref64 = None
    for v in range(len(views)):
        r32, o32 = oracle_forward(body, vm[v], pm[v], H, W)
        r64, o64 = oracle_forward(body, vm[v], pm[v], H, W, dtype=np.float64)
        c = o32.color
        gc = (np.sign(np.clip(c, 0, 1) - target[v]) * ((c >= 0) & (c <= 1)) / target.size).astype(np.float32)
        g32 = r32.backward(gc)
        g64 = r64.backward(gc.astype(np.float64))
        ref32 = g32 if ref32 is None else {k: ref32[k] + g32[k] for k in g32}
        ref64 = g64 if ref64 is None else {k: ref64[k] + g64[k] for k in g64}
    print(f"exact_exp={exact}")
    for k in ("means3D", "cov3D", "colors", "opacities"):
        got = t[k].grad[0].cpu().numpy().astype(np.float64)
        w64 = ref64[k].astype(np.float64)
        w32 = ref32[k].astype(np.float64)
        s = np.abs(w64).max()
        def stats(a, b):
            e = np.abs(a - b)
            big = np.abs(b) > 1e-3 * s
            rel = e[big] / np.abs(b[big])
            return f"max|err|/max|g| {e.max() / s:.2e}  rel(p50 {np.percentile(rel, 50):.1e} p99 {np.percentile(rel, 99):.1e} max {rel.max():.1e})"
        print(f"  {k:10s} max|g| {s:.3e}  gpu-vs-f64: {stats(got, w64)} | oracle32-vs-f64: {stats(w32, w64)} | gpu-vs-oracle32: {stats(got, w32)}")
