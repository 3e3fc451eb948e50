"""CPU-side statistics of one tile's list (oracle geometry): how many records survive the per-warp bbox cull,
how many pass the alpha test, when warps finish."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import oracle
from sigman_release_b200 import cameras, scenes
VIEWS = [30, 37, 45, 53, 65, 85, 0, 8]
r_idx, tile = int(sys.argv[1]), int(sys.argv[2])
sc = scenes.body_gaussians(100_000, seed=0)
vm, pm, _ = cameras.orbit_cameras(VIEWS)
tan = cameras.tan_half_fov()
R = oracle.Rasterizer(np.float32)
o = R.forward(sc["means3D"], sc["cov3D"], sc["colors"], sc["opacities"], vm[r_idx].reshape(-1), pm[r_idx].reshape(-1), tan, tan, (1,1,1), 512, 512)
g = R.geom(); b = R.binning()
s, e = b["ranges"][tile]
ids = b["point_list"][s:e]
x, y = g["xy"][ids, 0], g["xy"][ids, 1]
A, Bc, C = g["conic"][ids].T
op = g["opacity"][ids]
det = A * C - Bc * Bc
ca, cc = C / det, A / det      # cov2D diag (inverse of conic)
tau = np.log(255 * np.maximum(op, 1e-9))
ex = np.where(op >= 1/255, np.sqrt(np.maximum(2 * tau * ca, 0)) + 0.01, -1e30)
ey = np.where(op >= 1/255, np.sqrt(np.maximum(2 * tau * cc, 0)) + 0.01, -1e30)
print("tile", tile, "n", len(ids), "radius mean", g["xy"].shape, "ex mean %.2f ey mean %.2f" % (ex[ex>0].mean(), ey[ey>0].mean()))
tx, ty = tile % 32, tile // 32
ncon = b["n_contrib"][ty*16:(ty+1)*16, tx*16:(tx+1)*16]
print("n_contrib min/mean/max", ncon.min(), ncon.mean(), ncon.max())
tot_surv = 0
for w in range(8):
    bx0 = tx*16 + (w & 1)*8; by0 = ty*16 + (w >> 1)*4
    surv = (x + ex >= bx0) & (x - ex <= bx0 + 7) & (y + ey >= by0) & (y - ey <= by0 + 3)
    # pixel evals
    px = np.arange(bx0, bx0+8)[None, :, None]; py = np.arange(by0, by0+4)[:, None, None]
    dx = x[None, None, surv] - px; dy = y[None, None, surv] - py
    power = -0.5*(A[surv]*dx*dx + C[surv]*dy*dy) - Bc[surv]*dx*dy
    alpha = np.minimum(0.99, op[surv]*np.exp(power))
    passed = (power <= 0) & (alpha >= 1/255)
    wl = ncon[(w>>1)*4:(w>>1)*4+4, (w&1)*8:(w&1)*8+8]
    anypass = passed.any(axis=(0,1))
    print(f" warp {w}: survivors {surv.sum():5d} ({surv.mean()*100:4.1f}%), survivors with any passing lane {anypass.sum():5d}, lane-pass rate {passed.mean()*100:4.1f}%, n_contrib max {wl.max()} min {wl.min()}")
    tot_surv += surv.sum()
print("total warp-survivors", tot_surv, "per entry", tot_surv/len(ids))
