"""Dense (non-tiled) PyTorch fp64 re-derivation of the splatting maths, differentiable by autograd.

Independent of oracle/sgr_oracle.cpp: no sorted instance lists, no per-tile ranges, no hand-written
backward.  Every pixel looks at every Gaussian; the tile-rectangle rule of the published algorithm
("a Gaussian is only considered by pixels of the 16x16 tiles its 3-sigma square overlaps") is applied
as a mask.  Used by tests/test_oracle.py to cross-check the oracle's forward and analytic backward on
small scenes (tens of Gaussians, <= 48x48 pixels).
"""
import torch


def render_dense(means3D, cov6, colors, opac, view, proj, tanfovx, tanfovy, bg, H, W, ndc_offset=None):
    """All tensors float64.  view/proj are the flat column-major [16] arrays the rasteriser receives.
    Returns color [3,H,W], depth [1,H,W], alpha [1,H,W]."""
    dt = means3D.dtype
    N = means3D.shape[0]
    Wm = view.reshape(4, 4).t()           # W2C (row-major maths matrix)
    PV = proj.reshape(4, 4).t()           # P @ W2C
    ones = torch.ones(N, 1, dtype=dt)
    ph = torch.cat([means3D, ones], 1)
    p_view = ph @ Wm.t()
    p_hom = ph @ PV.t()
    pw = 1.0 / (p_hom[:, 3] + 1e-7)
    ndc = p_hom[:, :2] * pw[:, None]
    if ndc_offset is not None:
        ndc = ndc + ndc_offset            # zero leaf used to read dL/dmean2D (NDC units)
    px = ((ndc[:, 0] + 1.0) * W - 1.0) * 0.5
    py = ((ndc[:, 1] + 1.0) * H - 1.0) * 0.5
    tz = p_view[:, 2]
    fx = W / (2.0 * tanfovx)
    fy = H / (2.0 * tanfovy)
    limx, limy = 1.3 * tanfovx, 1.3 * tanfovy
    txc = torch.clamp(p_view[:, 0] / tz, -limx, limx) * tz
    tyc = torch.clamp(p_view[:, 1] / tz, -limy, limy) * tz
    J = torch.zeros(N, 2, 3, dtype=dt)
    J[:, 0, 0] = fx / tz
    J[:, 0, 2] = -fx * txc / (tz * tz)
    J[:, 1, 1] = fy / tz
    J[:, 1, 2] = -fy * tyc / (tz * tz)
    Rw = Wm[:3, :3]
    M = J @ Rw
    S = torch.stack([cov6[:, 0], cov6[:, 1], cov6[:, 2], cov6[:, 1], cov6[:, 3], cov6[:, 4], cov6[:, 2], cov6[:, 4],
                     cov6[:, 5]], 1).reshape(N, 3, 3)
    cov2 = M @ S @ M.transpose(1, 2)
    a = cov2[:, 0, 0] + 0.3
    b = cov2[:, 0, 1]
    c = cov2[:, 1, 1] + 0.3
    det = a * c - b * b
    cA, cB, cC = c / det, -b / det, a / det
    # --- non-differentiable visibility: near cull, 3-sigma radius, tile rectangle (16x16 tiles)
    with torch.no_grad():
        mid = 0.5 * (a + c)
        lam = mid + torch.sqrt(torch.clamp(mid * mid - det, min=0.1))
        radius = torch.ceil(3.0 * torch.sqrt(lam))
        gx, gy = (W + 15) // 16, (H + 15) // 16
        rminx = torch.clamp(torch.trunc((px - radius) / 16), 0, gx)
        rminy = torch.clamp(torch.trunc((py - radius) / 16), 0, gy)
        rmaxx = torch.clamp(torch.trunc((px + radius + 15) / 16), 0, gx)
        rmaxy = torch.clamp(torch.trunc((py + radius + 15) / 16), 0, gy)
        visible = (tz > 0.2) & ((rmaxx - rminx) * (rmaxy - rminy) > 0)
        order = torch.argsort(tz.float(), stable=True)       # depth keys are fp32 bit patterns upstream
    ys, xs = torch.meshgrid(torch.arange(H, dtype=dt), torch.arange(W, dtype=dt), indexing="ij")
    tile_x = (xs // 16)[None]
    tile_y = (ys // 16)[None]
    in_rect = ((tile_x >= rminx[:, None, None]) & (tile_x < rmaxx[:, None, None]) &
               (tile_y >= rminy[:, None, None]) & (tile_y < rmaxy[:, None, None]) & visible[:, None, None])
    dx = px[:, None, None] - xs[None]
    dy = py[:, None, None] - ys[None]
    power = -0.5 * (cA[:, None, None] * dx * dx + cC[:, None, None] * dy * dy) - cB[:, None, None] * dx * dy
    alpha = torch.clamp(opac[:, None, None] * torch.exp(power), max=0.99)
    keep = in_rect & (power <= 0) & (alpha >= 1.0 / 255.0)
    alpha = torch.where(keep, alpha, torch.zeros_like(alpha))
    alpha = alpha[order]
    # transmittance before each Gaussian; early termination when T*(1-alpha) < 1e-4
    one_minus = 1.0 - alpha
    T_after = torch.cumprod(one_minus, 0)
    T_before = torch.cat([torch.ones(1, H, W, dtype=dt), T_after[:-1]], 0)
    with torch.no_grad():
        stop = (T_after < 1e-4) & (alpha > 0)
        alive = (torch.cumsum(stop.to(torch.int32), 0) == 0)
    w = torch.where(alive, alpha * T_before, torch.zeros_like(alpha))
    col = colors[order]
    z = tz[order]
    C = (w[:, None] * col[:, :, None, None]).sum(0)
    D = (w * z[:, None, None]).sum(0, keepdim=True)
    A = w.sum(0, keepdim=True)
    T_fin = 1.0 - A
    color = C + T_fin * bg[:, None, None]
    return color, D, A
