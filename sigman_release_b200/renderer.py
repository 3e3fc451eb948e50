"""Batched drop-in for ``GaussianRenderer`` of ``/root/reference/core/gaussians/gs.py:41-117`` and for
``simple_knn._C.distCUDA2`` (gs.py:6,70).

Same constructor argument (an options object with ``output_size_h``, ``output_size_w``, ``FoVy``), same
``render(gaussians, cam_view, cam_view_proj, cam_pos, bg_color=None, scale_modifier=0.5)`` signature and the same
``{"image": [B,V,3,H,W], "alpha": [B,V,1,H,W]}`` result — but the B x V Python double loop (one full rasteriser
pipeline, one host sync and >= 8 launches per view) is replaced by ONE batched launch set of the sm_100a library.
"""
from __future__ import annotations

import ctypes

import numpy as np
import torch

from . import _native
from .rasterizer import _ptr, rasterize_batch

__all__ = ["GaussianRenderer", "distCUDA2", "distCUDA2_batched", "get_covariance", "strip_lowerdiag", "prep_cov3d"]


def distCUDA2(points: torch.Tensor) -> torch.Tensor:
    """Mean squared distance to the 3 nearest other points, float32 [N] (``simple_knn._C.distCUDA2``, gs.py:70)."""
    L = _native.lib()
    if not points.is_cuda:
        raise ValueError("distCUDA2 needs a CUDA tensor (no CPU path)")
    pts = points.detach().contiguous().float()
    if pts.dim() != 2 or pts.shape[1] != 3:
        raise ValueError("points must be [N,3]")
    N = int(pts.shape[0])
    out = torch.empty((N,), dtype=torch.float32, device=pts.device)
    nbytes = int(L.sgr_knn_scratch_bytes(N))
    with torch.cuda.device(pts.device):
        scratch = torch.empty((nbytes,), dtype=torch.uint8, device=pts.device)
        st = torch.cuda.current_stream(pts.device)
        _native.check(L.sgr_knn_mean_dist2(_ptr(pts), N, _ptr(out), _ptr(scratch), nbytes,
                                           ctypes.c_void_p(st.cuda_stream)))
    return out


def distCUDA2_batched(points: torch.Tensor) -> torch.Tensor:
    """``distCUDA2`` for B independent point sets [B,N,3] -> [B,N] in one launch set (gs.py:62-70 calls it per subject)."""
    L = _native.lib()
    if not points.is_cuda:
        raise ValueError("distCUDA2_batched needs a CUDA tensor (no CPU path)")
    pts = points.detach().contiguous().float()
    if pts.dim() != 3 or pts.shape[2] != 3:
        raise ValueError("points must be [B,N,3]")
    B, N = int(pts.shape[0]), int(pts.shape[1])
    out = torch.empty((B, N), dtype=torch.float32, device=pts.device)
    nbytes = int(L.sgr_knn_scratch_bytes_batched(B, N))
    with torch.cuda.device(pts.device):
        scratch = torch.empty((nbytes,), dtype=torch.uint8, device=pts.device)
        st = torch.cuda.current_stream(pts.device)
        _native.check(L.sgr_knn_mean_dist2_batched(_ptr(pts), B, N, _ptr(out), _ptr(scratch), nbytes,
                                                   ctypes.c_void_p(st.cuda_stream)))
    return out


def strip_lowerdiag(L: torch.Tensor) -> torch.Tensor:
    """[..., 3, 3] -> [..., 6] in the order (xx, xy, xz, yy, yz, zz) of gs.py:29-38."""
    return torch.stack([L[..., 0, 0], L[..., 0, 1], L[..., 0, 2], L[..., 1, 1], L[..., 1, 2], L[..., 2, 2]], dim=-1)


def get_covariance(scaling: torch.Tensor, rotation: torch.Tensor, scaling_modifier: float = 1) -> torch.Tensor:
    """R diag(s)^2 R^T packed to 6 values (gs.py:17-23); works on [N,...] and [B,N,...]."""
    RS = rotation * (scaling * scaling).unsqueeze(-2)            # R diag(s^2)
    return strip_lowerdiag(RS @ rotation.transpose(-1, -2))


class _PrepCov3D(torch.autograd.Function):
    """gs.py:69-73 in one kernel: cov3D [.., 6] from raw scale [.., 3], rotation-like matrix [.., 3, 3] and the
    (detached) kNN mean squared distance [..]; backward to scale and rotation."""

    @staticmethod
    def forward(ctx, scale_raw, rotation, dist2, bf16_autocast):
        L = _native.lib()
        n = scale_raw.numel() // 3
        cov = torch.empty(tuple(scale_raw.shape[:-1]) + (6,), dtype=torch.float32, device=scale_raw.device)
        with torch.cuda.device(scale_raw.device):
            st = torch.cuda.current_stream(scale_raw.device)
            _native.check(L.sgr_prep_cov3d(_ptr(scale_raw), _ptr(rotation), _ptr(dist2), n, int(bool(bf16_autocast)),
                                           _ptr(cov), ctypes.c_void_p(st.cuda_stream)))
        ctx.save_for_backward(scale_raw, rotation, dist2)
        ctx.bf16 = int(bool(bf16_autocast))
        return cov

    @staticmethod
    def backward(ctx, g_cov):
        L = _native.lib()
        scale_raw, rotation, dist2 = ctx.saved_tensors
        n = scale_raw.numel() // 3
        g_cov = g_cov.contiguous().float()
        ds = torch.empty_like(scale_raw)
        dr = torch.empty_like(rotation)
        with torch.cuda.device(scale_raw.device):
            st = torch.cuda.current_stream(scale_raw.device)
            _native.check(L.sgr_prep_cov3d_backward(_ptr(scale_raw), _ptr(rotation), _ptr(dist2), n, ctx.bf16,
                                                    _ptr(g_cov), _ptr(ds), _ptr(dr), ctypes.c_void_p(st.cuda_stream)))
        return ds, dr, None, None


def prep_cov3d(scale_raw: torch.Tensor, rotation: torch.Tensor, dist2: torch.Tensor, bf16_autocast: bool = False):
    """Fused ``(s + 1) * sqrt(max(d2, 1e-7))`` -> ``R diag(scale^2) R^T`` -> 6-pack of gs.py:69-73 (and its backward).
    scale_raw [..,3], rotation [..,3,3], dist2 [..] (treated as a constant, like the ``.detach()`` of gs.py:71)."""
    lead = tuple(scale_raw.shape[:-1])
    if tuple(rotation.shape) != lead + (3, 3) or tuple(dist2.shape) != lead:
        raise ValueError("prep_cov3d: scale_raw [..,3], rotation [..,3,3], dist2 [..] must agree")
    for name, t in (("scale_raw", scale_raw), ("rotation", rotation), ("dist2", dist2)):
        if not t.is_cuda or t.dtype != torch.float32:
            raise ValueError(f"{name} must be a float32 CUDA tensor")
    return _PrepCov3D.apply(scale_raw.contiguous(), rotation.contiguous(), dist2.detach().contiguous(), bf16_autocast)


class GaussianRenderer:
    def __init__(self, opt):
        self.opt = opt
        self.bg_color = torch.tensor([1, 1, 1], dtype=torch.float32, device="cuda")
        self.tan_half_fov = np.tan(0.5 * self.opt.FoVy)
        # True reproduces the bf16 rounding the reference's get_covariance bmm's see under accelerate's autocast
        self.bf16_autocast = False
        # True: exp(power) as the oracle's fixed IEEE sequence (bit-exact parity mode, include/sgr.h SGR_FLAG_EXACT_EXP)
        self.exact_exp = False

    def prepare(self, gaussians):
        """Per-subject preparation of gs.py:64-73, batched over B: returns (means3D, cov3D [B,N,6], rgb, opacity)."""
        means3D = gaussians["position"].contiguous().float()
        opacity = gaussians["opacity"].contiguous().float()
        scales = gaussians["scale"].contiguous().float()
        rot = gaussians["cov3d"].contiguous().float()
        rgbs = gaussians["rgb"].contiguous().float()
        B = means3D.shape[0]
        with torch.no_grad():
            dist2 = distCUDA2_batched(means3D)                  # detached kNN factor (gs.py:70-71), all subjects
        cov3D = prep_cov3d(scales, rot, dist2, bf16_autocast=self.bf16_autocast)
        return means3D, cov3D, rgbs, opacity

    def render(self, gaussians, cam_view, cam_view_proj, cam_pos, bg_color=None, scale_modifier=0.5):
        device = gaussians["position"].device
        B, V = cam_view.shape[:2]
        H, W = int(self.opt.output_size_h), int(self.opt.output_size_w)
        means3D, cov3D, rgbs, opacity = self.prepare(gaussians)
        bg = self.bg_color if bg_color is None else bg_color
        bg = bg.to(device=device, dtype=torch.float32)
        # scale_modifier and cam_pos do not enter the cov3D_precomp / colors_precomp path.
        # gs.py:107 `rendered_image.clamp(0, 1)` is fused into the blend epilogue; its gradient mask lives in the
        # rasteriser state and is applied inside the backward kernel (no clamp kernel, no saved image).
        image, _radii, _depth, alpha = rasterize_batch(
            means3D, cov3D, rgbs, opacity, cam_view.float(), cam_view_proj.float(), bg, H, W,
            self.tan_half_fov, self.tan_half_fov, clamp_color=True, exact_exp=self.exact_exp)
        return {"image": image, "alpha": alpha}
