"""Synthetic Gaussian sets for the parity tests and the benchmark (host side, numpy, seeded).

No dataset, checkpoint or licensed SMPL-X asset is available, so the workloads named in
BASELINE.json are generated procedurally:

  * ``random_gaussians``  — BASELINE config 1 (10K random Gaussians in the unit cube).
  * ``body_gaussians``    — a ~100K-Gaussian human-shaped surface set standing in for the SMPL-X
    template of SIGMAN (one Gaussian per surface sample; head and hands are sampled more densely,
    like the SMPL-X mesh; attribute distributions follow
    ``/root/reference/core/modules/autoencoder.py:295-310`` and the scale composition of
    ``/root/reference/core/gaussians/gs.py:70-73``).

Everything here is input preparation; none of it is on the timed path.
"""
from __future__ import annotations

import numpy as np


def _sigmoid(x):
    return 1.0 / (1.0 + np.exp(-x))


def quat_to_rotmat(q: np.ndarray) -> np.ndarray:
    """(r, x, y, z) quaternions [N,4] -> rotation matrices [N,3,3] (normalised first)."""
    q = q / np.linalg.norm(q, axis=1, keepdims=True)
    r, x, y, z = q[:, 0], q[:, 1], q[:, 2], q[:, 3]
    R = np.empty((q.shape[0], 3, 3), q.dtype)
    R[:, 0, 0] = 1 - 2 * (y * y + z * z); R[:, 0, 1] = 2 * (x * y - r * z); R[:, 0, 2] = 2 * (x * z + r * y)
    R[:, 1, 0] = 2 * (x * y + r * z); R[:, 1, 1] = 1 - 2 * (x * x + z * z); R[:, 1, 2] = 2 * (y * z - r * x)
    R[:, 2, 0] = 2 * (x * z - r * y); R[:, 2, 1] = 2 * (y * z + r * x); R[:, 2, 2] = 1 - 2 * (x * x + y * y)
    return R


def covariance6(scale: np.ndarray, rot: np.ndarray) -> np.ndarray:
    """R diag(s^2) R^T packed (xx,xy,xz,yy,yz,zz) — gs.py:17-38 (get_covariance + strip_lowerdiag)."""
    L = rot * scale[:, None, :]
    S = L @ np.transpose(L, (0, 2, 1))
    return np.stack([S[:, 0, 0], S[:, 0, 1], S[:, 0, 2], S[:, 1, 1], S[:, 1, 2], S[:, 2, 2]], axis=1)


def random_gaussians(n: int = 10_000, seed: int = 0) -> dict:
    """BASELINE config 1: xyz~U(-.5,.5)^3, s~U(.005,.03), random rotations, rgb~U(0,1), o~U(.05,1)."""
    rng = np.random.default_rng(seed)
    xyz = rng.uniform(-0.5, 0.5, (n, 3))
    scale = rng.uniform(0.005, 0.03, (n, 3))
    quat = rng.normal(size=(n, 4))
    quat /= np.linalg.norm(quat, axis=1, keepdims=True)
    rot = quat_to_rotmat(quat)
    rgb = rng.uniform(0.0, 1.0, (n, 3))
    opacity = rng.uniform(0.05, 1.0, (n,))
    f = np.float32
    return dict(means3D=xyz.astype(f), cov3D=covariance6(scale, rot).astype(f), colors=rgb.astype(f),
                opacities=opacity.astype(f), scales=scale.astype(f), rotations=quat.astype(f))


# ----------------------------------------------------------------------------- procedural body
# (p0, p1, radius, (sx, sz) cross-section stretch, sampling-density weight)
_PARTS = [
    ("torso",   (0.00, -0.05, 0.00), (0.00, 0.45, 0.00), 0.135, (1.25, 0.80), 1.0),
    ("neck",    (0.00, 0.50, 0.00),  (0.00, 0.62, 0.01), 0.050, (1.0, 1.0), 1.0),
    ("head",    (0.00, 0.70, 0.02),  (0.00, 0.76, 0.02), 0.095, (0.90, 1.05), 9.0),
    ("uarm_l",  (0.20, 0.44, 0.00),  (0.45, 0.20, 0.00), 0.045, (1.0, 1.0), 1.0),
    ("uarm_r",  (-0.20, 0.44, 0.00), (-0.45, 0.20, 0.00), 0.045, (1.0, 1.0), 1.0),
    ("farm_l",  (0.45, 0.20, 0.00),  (0.62, -0.02, 0.02), 0.035, (1.0, 1.0), 1.0),
    ("farm_r",  (-0.45, 0.20, 0.00), (-0.62, -0.02, 0.02), 0.035, (1.0, 1.0), 1.0),
    ("hand_l",  (0.64, -0.05, 0.02), (0.70, -0.13, 0.03), 0.035, (1.0, 0.5), 16.0),
    ("hand_r",  (-0.64, -0.05, 0.02), (-0.70, -0.13, 0.03), 0.035, (1.0, 0.5), 16.0),
    ("thigh_l", (0.09, -0.08, 0.00), (0.12, -0.48, 0.00), 0.075, (1.0, 1.0), 1.0),
    ("thigh_r", (-0.09, -0.08, 0.00), (-0.12, -0.48, 0.00), 0.075, (1.0, 1.0), 1.0),
    ("shin_l",  (0.12, -0.48, 0.00), (0.13, -0.82, -0.01), 0.050, (1.0, 1.0), 1.0),
    ("shin_r",  (-0.12, -0.48, 0.00), (-0.13, -0.82, -0.01), 0.050, (1.0, 1.0), 1.0),
    ("foot_l",  (0.13, -0.85, -0.02), (0.13, -0.86, 0.13), 0.038, (1.0, 1.0), 2.0),
    ("foot_r",  (-0.13, -0.85, -0.02), (-0.13, -0.86, 0.13), 0.038, (1.0, 1.0), 2.0),
]


def _capsule_samples(rng, n, p0, p1, r, stretch):
    """n points on a capsule surface (uniform by area) + outward normals."""
    p0 = np.asarray(p0, float); p1 = np.asarray(p1, float)
    axis = p1 - p0
    L = np.linalg.norm(axis)
    a = axis / L
    # orthonormal frame (u, w, a)
    ref = np.array([0.0, 0.0, 1.0]) if abs(a[2]) < 0.9 else np.array([1.0, 0.0, 0.0])
    u = np.cross(a, ref); u /= np.linalg.norm(u)
    w = np.cross(a, u)
    area_cyl = 2 * np.pi * r * L
    area_cap = 4 * np.pi * r * r
    on_cyl = rng.uniform(size=n) < area_cyl / (area_cyl + area_cap)
    phi = rng.uniform(0, 2 * np.pi, n)
    t = rng.uniform(0, 1, n)
    # spherical caps
    cz = rng.uniform(-1, 1, n)
    cr = np.sqrt(np.maximum(0.0, 1 - cz * cz))
    nx = np.where(on_cyl, np.cos(phi), cr * np.cos(phi))
    ny = np.where(on_cyl, np.sin(phi), cr * np.sin(phi))
    nz = np.where(on_cyl, 0.0, cz)
    along = np.where(on_cyl, t * L, np.where(cz > 0, L, 0.0))
    sx, sz = stretch
    local_n = nx[:, None] * u[None] * sx + ny[:, None] * w[None] * sz + nz[:, None] * a[None]
    pts = p0[None] + along[:, None] * a[None] + r * local_n
    nrm = nx[:, None] * u[None] / sx + ny[:, None] * w[None] / sz + nz[:, None] * a[None]
    nrm /= np.linalg.norm(nrm, axis=1, keepdims=True)
    return pts, nrm


def _tangent_frames(nrm):
    ref = np.where(np.abs(nrm[:, 1:2]) < 0.9, np.array([[0.0, 1.0, 0.0]]), np.array([[1.0, 0.0, 0.0]]))
    t1 = np.cross(ref, nrm); t1 /= np.linalg.norm(t1, axis=1, keepdims=True)
    t2 = np.cross(nrm, t1)
    return np.stack([t1, t2, nrm], axis=2)          # columns = (t1, t2, n)


def _axis_angle_to_rotmat(v):
    th = np.linalg.norm(v, axis=1, keepdims=True)
    k = v / np.maximum(th, 1e-12)
    K = np.zeros((v.shape[0], 3, 3))
    K[:, 0, 1] = -k[:, 2]; K[:, 0, 2] = k[:, 1]; K[:, 1, 0] = k[:, 2]
    K[:, 1, 2] = -k[:, 0]; K[:, 2, 0] = -k[:, 1]; K[:, 2, 1] = k[:, 0]
    s = np.sin(th)[:, :, None]; c = np.cos(th)[:, :, None]
    return np.eye(3)[None] + s * K + (1 - c) * (K @ K)


def knn3_mean_dist2(xyz: np.ndarray) -> np.ndarray:
    """Host-side stand-in for ``distCUDA2`` during scene preparation (scipy KD-tree, exact)."""
    from scipy.spatial import cKDTree

    d, _ = cKDTree(xyz).query(xyz, k=4)
    return (d[:, 1:] ** 2).mean(axis=1)


def body_gaussians(n: int = 100_000, seed: int = 0, jitter: float = 0.0) -> dict:
    """A human-shaped Gaussian set (height ~1.7 m, centred at the origin, facing +z).

    ``jitter`` > 0 applies a per-subject random rigid perturbation (yaw, small translation) so that
    different seeds give different subjects for the batch configs.
    """
    rng = np.random.default_rng(seed)
    weights = []
    for _, p0, p1, r, _, dens in _PARTS:
        L = np.linalg.norm(np.asarray(p1) - np.asarray(p0))
        weights.append((2 * np.pi * r * L + 4 * np.pi * r * r) * dens)
    weights = np.asarray(weights); weights /= weights.sum()
    counts = np.floor(weights * n).astype(int)
    counts[0] += n - counts.sum()
    pts, nrms = [], []
    for (_, p0, p1, r, st, _), c in zip(_PARTS, counts):
        p, q = _capsule_samples(rng, int(c), p0, p1, r, st)
        pts.append(p); nrms.append(q)
    xyz = np.concatenate(pts); nrm = np.concatenate(nrms)
    perm = rng.permutation(n)                 # template order is not spatially sorted
    xyz, nrm = xyz[perm], nrm[perm]
    if jitter > 0:
        yaw = rng.uniform(-np.pi, np.pi) * jitter
        Ry = np.array([[np.cos(yaw), 0, np.sin(yaw)], [0, 1, 0], [-np.sin(yaw), 0, np.cos(yaw)]])
        xyz = xyz @ Ry.T + rng.normal(scale=0.03 * jitter, size=(1, 3))
        nrm = nrm @ Ry.T
    base = np.sqrt(np.maximum(knn3_mean_dist2(xyz), 1e-7))                    # gs.py:70-71
    scale = base[:, None] * rng.uniform(0.0, 2.0, (n, 3))                    # (s + 1) * sqrt(d2), s in (-1, 1)
    rot = _tangent_frames(nrm) @ _axis_angle_to_rotmat(rng.uniform(-np.pi / 2, np.pi / 2, (n, 3)) * 0.25)
    rgb = _sigmoid(rng.normal(size=(n, 3))) * 1.002 - 0.001                    # autoencoder.py:305-306
    opacity = _sigmoid(rng.normal(loc=2.0, size=(n,)))
    f = np.float32
    return dict(means3D=xyz.astype(f), cov3D=covariance6(scale, rot).astype(f), colors=rgb.astype(f),
                opacities=opacity.astype(f), scales=scale.astype(f), rotmats=rot.astype(f))
