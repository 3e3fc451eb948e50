"""CPU-side statistics over a whole view: (Gaussian, 4x2 quarter) pairs kept by the bbox cull vs. pairs in which at
least one pixel passes the alpha test (the floor of any conservative cull) — sizing tool for the quarter masks."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import oracle
from sigman_release_b200 import cameras, scenes
VIEWS = [30, 37, 45, 53, 65, 85, 0, 8]
r_idx = int(sys.argv[1]) if len(sys.argv) > 1 else 0
sc = scenes.body_gaussians(100_000, seed=0)
vm, pm, _ = cameras.orbit_cameras(VIEWS)
tan = cameras.tan_half_fov()
R = oracle.Rasterizer(np.float32)
o = R.forward(sc["means3D"], sc["cov3D"], sc["colors"], sc["opacities"], vm[r_idx].reshape(-1), pm[r_idx].reshape(-1), tan, tan, (1,1,1), 512, 512)
g = R.geom(); b = R.binning()
pl = b["point_list"]; rg = b["ranges"]
tile_of = np.zeros(len(pl), np.int64)
for t, (s, e) in enumerate(rg):
    tile_of[s:e] = t
x, y = g["xy"][pl, 0].astype(np.float64), g["xy"][pl, 1].astype(np.float64)
A, Bc, C = (g["conic"][pl].T).astype(np.float64)
op = g["opacity"][pl].astype(np.float64)
det = A * C - Bc * Bc
ca, cc = C / det, A / det
tau = np.log(255 * np.maximum(op, 1e-9))
ex = np.where(op >= 1/255, np.sqrt(np.maximum(2 * tau * ca, 0)) + 0.01, -1e30)
ey = np.where(op >= 1/255, np.sqrt(np.maximum(2 * tau * cc, 0)) + 0.01, -1e30)
X0 = (tile_of % 32) * 16.0; Y0 = (tile_of // 32) * 16.0
n = len(pl)
bbox_pairs = 0; exact_pairs = 0; pix_pass = 0; exact2x2 = 0; bbox2x2 = 0; exact_rowseg = 0
CH = 20000
for s in range(0, n, CH):
    sl = slice(s, min(n, s + CH))
    px = X0[sl, None, None] + np.arange(16)[None, None, :]
    py = Y0[sl, None, None] + np.arange(16)[None, :, None]
    dx = x[sl, None, None] - px; dy = y[sl, None, None] - py
    power = -0.5 * (A[sl, None, None] * dx * dx + C[sl, None, None] * dy * dy) - Bc[sl, None, None] * dx * dy
    ok = (power <= 0) & (op[sl, None, None] * np.exp(power) >= 1 / 255)
    pix_pass += ok.sum()
    q = ok.reshape(-1, 8, 2, 4, 4).any(axis=(2, 4))      # [inst, qrow(8), qcol(4)]
    exact_pairs += q.sum()
    q2 = ok.reshape(-1, 8, 2, 8, 2).any(axis=(2, 4))
    exact2x2 += q2.sum()
    rs = ok.reshape(-1, 16, 4, 4).any(axis=3)            # 4x1 row segments
    exact_rowseg += rs.sum()
    inb = (np.abs(dx) <= ex[sl, None, None]) & (np.abs(dy) <= ey[sl, None, None])
    bbox_pairs += inb.reshape(-1, 8, 2, 4, 4).any(axis=(2, 4)).sum()
    bbox2x2 += inb.reshape(-1, 8, 2, 8, 2).any(axis=(2, 4)).sum()
print(f"view {r_idx}: instances {n}, pixel passes {pix_pass} ({pix_pass/n:.1f}/inst)")
print(f" 4x2 quarters: bbox pairs {bbox_pairs} ({bbox_pairs*8/1e6:.1f}M lane evals), exact pairs {exact_pairs} ({exact_pairs*8/1e6:.1f}M), ratio {exact_pairs/bbox_pairs:.2f}")
print(f" 2x2 groups:   bbox pairs {bbox2x2} ({bbox2x2*4/1e6:.1f}M lane evals), exact pairs {exact2x2} ({exact2x2*4/1e6:.1f}M)")
print(f" 4x1 segments: exact pairs {exact_rowseg} ({exact_rowseg*4/1e6:.1f}M lane evals)")
