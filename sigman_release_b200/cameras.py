"""Camera conventions of the SIGMAN render path (host side, numpy).

Restates, without reading any reference file at run time:
  * the 90-view orbit rig of ``/root/reference/core/dataset/camera_full_calibration.json``
    (three rings of 30 views at elevation -30/0/+45 degrees, radius 2.5, OpenCV axes, K = 1100 px
    focal on a 1024^2 sensor) — procedural, checked against a golden subset in ``tests/golden``;
  * ``getProjectionMatrix`` and the matrix layout handed to the rasteriser,
    ``/root/reference/core/dataset/dataloader_VAE.py:207-246``:
    ``cam_view = W2C^T``, ``cam_view_proj = W2C^T @ P^T`` (row-major storage of the transposes, i.e.
    the flat arrays are column-major W2C and P*W2C), ``cam_pos = inv(W2C)[:3, 3]``.
"""
from __future__ import annotations

import math

import numpy as np

# /root/reference/core/model_config/VAE.py:32-37
FOVY = 0.8712626851529752
ZNEAR = 0.1
ZFAR = 100.0
# training / eval view lists, /root/reference/core/dataset/dataloader_VAE.py:77-79
TRAIN_FIXED_VIEWS = (30, 37, 45, 53, 65, 85)
EVAL_VIEWS = (30, 37, 45, 53, 65, 85, 0, 8, 82, 60)

_ELEVATIONS_DEG = (-30.0, 0.0, 45.0)
ORBIT_RADIUS = 2.5
NUM_ORBIT_VIEWS = 90


def orbit_w2c(view_id: int) -> np.ndarray:
    """World-to-camera 4x4 (float64) of orbit view ``view_id`` in [0, 90)."""
    if not 0 <= view_id < NUM_ORBIT_VIEWS:
        raise ValueError(f"view_id {view_id} outside [0, {NUM_ORBIT_VIEWS})")
    el = math.radians(_ELEVATIONS_DEG[view_id // 30])
    az = math.radians(12.0 * (view_id % 30))
    pos = ORBIT_RADIUS * np.array([math.sin(az) * math.cos(el), math.sin(el), math.cos(az) * math.cos(el)])
    z_cam = -pos / np.linalg.norm(pos)
    x_cam = np.cross(z_cam, np.array([0.0, 1.0, 0.0]))
    x_cam /= np.linalg.norm(x_cam)
    y_cam = np.cross(z_cam, x_cam)
    R = np.stack([x_cam, y_cam, z_cam], axis=0)
    w2c = np.eye(4)
    w2c[:3, :3] = R
    w2c[:3, 3] = -R @ pos
    return w2c


def projection_matrix(znear: float = ZNEAR, zfar: float = ZFAR, fovx: float = FOVY, fovy: float = FOVY,
                      K: np.ndarray | None = None, img_h: int | None = None, img_w: int | None = None) -> np.ndarray:
    """P (4x4, float64) as built by dataloader_VAE.py:218-246 (z_sign = +1, P[3,2] = 1)."""
    if K is None:
        top = math.tan(fovy / 2) * znear
        bottom = -top
        right = math.tan(fovx / 2) * znear
        left = -right
    else:
        near_fx = znear / K[0, 0]
        near_fy = znear / K[1, 1]
        left = -(img_w - K[0, 2]) * near_fx
        right = K[0, 2] * near_fx
        bottom = (K[1, 2] - img_h) * near_fy
        top = K[1, 2] * near_fy
    P = np.zeros((4, 4))
    P[0, 0] = 2.0 * znear / (right - left)
    P[1, 1] = 2.0 * znear / (top - bottom)
    P[0, 2] = (right + left) / (right - left)
    P[1, 2] = (top + bottom) / (top - bottom)
    P[3, 2] = 1.0
    P[2, 2] = zfar / (zfar - znear)
    P[2, 3] = -(zfar * znear) / (zfar - znear)
    return P


def sigman_projection() -> np.ndarray:
    """P for the shipped intrinsics K=[[1100,0,512],[0,1100,512]] on 1024^2 (P[0,0] = 2.1484375)."""
    K = np.array([[1100.0, 0.0, 512.0], [0.0, 1100.0, 512.0], [0.0, 0.0, 1.0]])
    return projection_matrix(K=K, img_h=1024, img_w=1024)


def tan_half_fov() -> float:
    """``np.tan(0.5 * opt.FoVy)`` of /root/reference/core/gaussians/gs.py:47 (= 512/1100)."""
    return float(np.tan(0.5 * FOVY))


def rasterizer_matrices(w2c: np.ndarray, P: np.ndarray | None = None):
    """(viewmatrix, projmatrix, campos) as float32 arrays in the layout gs.py:78-80 passes on."""
    if P is None:
        P = sigman_projection()
    cam_view = w2c.T
    cam_view_proj = cam_view @ P.T
    cam_pos = np.linalg.inv(w2c)[:3, 3]
    return cam_view.astype(np.float32), cam_view_proj.astype(np.float32), cam_pos.astype(np.float32)


def orbit_cameras(view_ids) -> tuple[np.ndarray, np.ndarray, np.ndarray]:
    """Stacked (cam_view [V,4,4], cam_view_proj [V,4,4], cam_pos [V,3]) float32 for the given orbit views."""
    P = sigman_projection()
    vs, ps, cs = zip(*(rasterizer_matrices(orbit_w2c(int(v)), P) for v in view_ids))
    return np.stack(vs), np.stack(ps), np.stack(cs)
