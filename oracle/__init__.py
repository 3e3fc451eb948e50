"""CPU oracle for the Gaussian-splat rasteriser hot path — TEST INFRASTRUCTURE, PARITY UNPINNED.

ctypes front end of ``oracle/sgr_oracle.cpp`` (read that file's header first).  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import
this package; the product package ``sigman_release_b200`` never does.

Restates what ``/root/reference/core/gaussians/gs.py:99-106`` executes (the third-party
``diff_gaussian_rasterization`` forward/backward) and
``gs.py:70`` (``simple_knn.distCUDA2``, Appendix B).
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from dataclasses import dataclass

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def _cpu_has_v3() -> bool:
    try:
        with open("/proc/cpuinfo") as f:
            for line in f:
                if line.startswith("flags"):
                    fl = set(line.split(":", 1)[1].split())
                    return {"avx2", "fma", "bmi2"} <= fl
    except OSError:
        pass
    return False


def build(force: bool = False) -> None:
    """Compile the oracle with the committed Makefile (g++ only; seconds)."""
    src = os.path.join(_HERE, "sgr_oracle.cpp")
    out = os.path.join(_HERE, "libsgr_oracle.so")
    if force or not os.path.exists(out) or (
        os.path.exists(src) and os.path.getmtime(out) < os.path.getmtime(src)
    ):
        subprocess.check_call(["make", "-C", _HERE, "-s", "-B"])


def lib() -> ctypes.CDLL:
    global _LIB
    if _LIB is not None:
        return _LIB
    generic = os.path.join(_HERE, "libsgr_oracle.so")
    v3 = os.path.join(_HERE, "libsgr_oracle_v3.so")
    if not os.path.exists(generic):
        build()
    path = v3 if (_cpu_has_v3() and os.path.exists(v3)) else generic
    L = ctypes.CDLL(path)
    L.oracle_expf.restype = ctypes.c_float
    L.oracle_expf.argtypes = [ctypes.c_float]
    L.oracle_num_threads.restype = ctypes.c_int
    L.oracle_create_f32.restype = ctypes.c_void_p
    L.oracle_create_f64.restype = ctypes.c_void_p
    L.oracle_destroy_f32.argtypes = [ctypes.c_void_p]
    L.oracle_destroy_f64.argtypes = [ctypes.c_void_p]
    L.oracle_num_instances_f32.restype = ctypes.c_uint64
    L.oracle_num_instances_f32.argtypes = [ctypes.c_void_p]
    _LIB = L
    return L


def _p(a: np.ndarray):
    return a.ctypes.data_as(ctypes.c_void_p)


def expf(x) -> np.ndarray:
    """exp_spec() of the oracle, elementwise (fp32)."""
    L = lib()
    x = np.asarray(x, dtype=np.float32)
    out = np.empty_like(x)
    flat_in, flat_out = x.reshape(-1), out.reshape(-1)
    for i in range(flat_in.size):
        flat_out[i] = L.oracle_expf(ctypes.c_float(float(flat_in[i])))
    return out


def num_threads() -> int:
    return int(lib().oracle_num_threads())


def set_num_threads(n: int) -> None:
    lib().oracle_set_num_threads(ctypes.c_int(int(n)))


def set_exp_mode(mode: str) -> None:
    """"spec": the fixed IEEE sequence exp_spec (default; what SGR_FLAG_EXACT_EXP reproduces bit for bit);
    "libm": the C library's expf — an independent exponential, used to measure how many alpha >= 1/255 /
    T >= 1e-4 decisions depend on the last bits of exp()."""
    lib().oracle_set_exp_mode(ctypes.c_int({"spec": 0, "libm": 1}[mode]))


@dataclass
class ForwardResult:
    color: np.ndarray          # [3,H,W]
    depth: np.ndarray          # [1,H,W]
    alpha: np.ndarray          # [1,H,W]
    radii: np.ndarray          # [N] int32
    num_instances: int
    evals: int                 # pixel*Gaussian evaluations walked
    blends: int                # accepted blends


class Rasterizer:
    """One render (one view of one Gaussian set) on the CPU.  dtype float32 (bit-exact spec) or float64."""

    def __init__(self, dtype=np.float32):
        self.dtype = np.dtype(dtype)
        assert self.dtype in (np.dtype(np.float32), np.dtype(np.float64))
        self._sfx = "f32" if self.dtype == np.float32 else "f64"
        self._L = lib()
        self._h = ctypes.c_void_p(getattr(self._L, f"oracle_create_{self._sfx}")())
        self._real = ctypes.c_float if self.dtype == np.float32 else ctypes.c_double
        self._saved = None

    def __del__(self):
        try:
            getattr(self._L, f"oracle_destroy_{self._sfx}")(self._h)
        except Exception:
            pass

    def _c(self, a, shape=None):
        a = np.ascontiguousarray(np.asarray(a, dtype=self.dtype))
        if shape is not None:
            a = a.reshape(shape)
        return a

    def forward(self, means3D, cov3D, colors, opacities, viewmatrix, projmatrix, tanfovx, tanfovy, bg, H, W) -> ForwardResult:
        N = int(np.asarray(means3D).shape[0])
        m = self._c(means3D, (N, 3)); c6 = self._c(cov3D, (N, 6)); col = self._c(colors, (N, 3))
        op = self._c(opacities, (N,)); vm = self._c(viewmatrix, (16,)); pm = self._c(projmatrix, (16,))
        bgc = self._c(bg, (3,))
        color = np.zeros((3, H, W), self.dtype); depth = np.zeros((1, H, W), self.dtype)
        alpha = np.zeros((1, H, W), self.dtype); radii = np.zeros((N,), np.int32)
        stats = np.zeros((3,), np.uint64)
        fn = getattr(self._L, f"oracle_forward_{self._sfx}")
        rc = fn(self._h, N, int(H), int(W), _p(m), _p(c6), _p(col), _p(op), _p(vm), _p(pm),
                self._real(float(tanfovx)), self._real(float(tanfovy)), _p(bgc), _p(color), _p(depth), _p(alpha),
                _p(radii), _p(stats))
        assert rc == 0
        self._saved = dict(N=N, H=H, W=W, m=m, c6=c6, col=col, vm=vm, pm=pm, bg=bgc, alpha=alpha,
                           tanfovx=float(tanfovx), tanfovy=float(tanfovy))
        return ForwardResult(color, depth, alpha, radii, int(stats[0]), int(stats[1]), int(stats[2]))

    def backward(self, dL_dcolor, dL_ddepth=None, dL_dalpha=None):
        """Returns dict(means3D [N,3], means2D [N,3], cov3D [N,6], colors [N,3], opacities [N])."""
        s = self._saved
        N, H, W = s["N"], s["H"], s["W"]
        gc = self._c(dL_dcolor, (3, H, W))
        gd = self._c(np.zeros((1, H, W)) if dL_ddepth is None else dL_ddepth, (1, H, W))
        ga = self._c(np.zeros((1, H, W)) if dL_dalpha is None else dL_dalpha, (1, H, W))
        o = dict(means3D=np.zeros((N, 3), self.dtype), means2D=np.zeros((N, 3), self.dtype),
                 cov3D=np.zeros((N, 6), self.dtype), colors=np.zeros((N, 3), self.dtype),
                 opacities=np.zeros((N,), self.dtype))
        fn = getattr(self._L, f"oracle_backward_{self._sfx}")
        rc = fn(self._h, _p(s["m"]), _p(s["c6"]), _p(s["col"]), _p(s["vm"]), _p(s["pm"]),
                self._real(s["tanfovx"]), self._real(s["tanfovy"]), _p(s["bg"]), _p(s["alpha"]), _p(gc), _p(gd),
                _p(ga), _p(o["means3D"]), _p(o["means2D"]), _p(o["cov3D"]), _p(o["colors"]), _p(o["opacities"]))
        assert rc == 0
        return o

    # ---- stage-by-stage state (fp32 only) for GPU parity checks
    def geom(self):
        assert self.dtype == np.float32
        N = self._saved["N"]
        geom = np.zeros((N, 7), np.float32); rect = np.zeros((N, 4), np.int32); tt = np.zeros((N,), np.uint32)
        self._L.oracle_get_geom_f32(self._h, _p(geom), _p(rect), _p(tt))
        return dict(depth=geom[:, 0], xy=geom[:, 1:3], conic=geom[:, 3:6], opacity=geom[:, 6], rect=rect,
                    tiles_touched=tt)

    def binning(self):
        assert self.dtype == np.float32
        s = self._saved
        n = int(self._L.oracle_num_instances_f32(self._h))
        tiles = ((s["W"] + 15) // 16) * ((s["H"] + 15) // 16)
        pl = np.zeros((max(n, 1),), np.uint32); rg = np.zeros((tiles, 2), np.uint32)
        nc = np.zeros((s["H"] * s["W"],), np.uint32)
        self._L.oracle_get_binning_f32(self._h, _p(pl), _p(rg), _p(nc))
        return dict(point_list=pl[:n], ranges=rg, n_contrib=nc.reshape(s["H"], s["W"]))


def cov3d_from_scale_rot(scales, rots, mod=1.0, dtype=np.float32) -> np.ndarray:
    L = lib()
    dt = np.dtype(dtype)
    s = np.ascontiguousarray(np.asarray(scales, dt)); r = np.ascontiguousarray(np.asarray(rots, dt))
    N = s.shape[0]
    out = np.zeros((N, 6), dt)
    if dt == np.float32:
        L.oracle_cov3d_from_scale_rot_f32(N, _p(s), _p(r), ctypes.c_float(mod), _p(out))
    else:
        L.oracle_cov3d_from_scale_rot_f64(N, _p(s), _p(r), ctypes.c_double(mod), _p(out))
    return out


def knn_mean_dist2(points) -> np.ndarray:
    """Brute-force mean squared distance to the 3 nearest other points (distCUDA2 semantics)."""
    L = lib()
    p = np.ascontiguousarray(np.asarray(points, np.float32))
    out = np.zeros((p.shape[0],), np.float32)
    L.oracle_knn_mean_dist2(p.shape[0], _p(p), _p(out))
    return out


def sh_colors(means3D, shs, campos, degree, dtype=np.float32):
    """upstream computeColorFromSH: (colors [N,3] clamped at 0, clamped flags [N,3] bool).  shs [N,K,3]."""
    L = lib()
    dt = np.dtype(dtype)
    m = np.ascontiguousarray(np.asarray(means3D, dt)); sh = np.ascontiguousarray(np.asarray(shs, dt))
    cp = np.ascontiguousarray(np.asarray(campos, dt).reshape(3))
    N, K = m.shape[0], sh.shape[1]
    assert (degree + 1) ** 2 <= K
    out = np.zeros((N, 3), dt); cl = np.zeros((N, 3), np.uint8)
    fn = L.oracle_sh_colors_f32 if dt == np.float32 else L.oracle_sh_colors_f64
    fn(ctypes.c_int(N), ctypes.c_int(int(degree)), ctypes.c_int(K), _p(m), _p(sh), _p(cp), _p(out), _p(cl))
    return out, cl.astype(bool)


def prep_cov3d(scale_raw, rot, dist2, dtype=np.float32) -> np.ndarray:
    """gs.py:69-73: Sigma = R diag(((s + 1) sqrt(max(d2, 1e-7)))^2) R^T packed to 6 values.  rot [n,3,3]."""
    L = lib()
    dt = np.dtype(dtype)
    s = np.ascontiguousarray(np.asarray(scale_raw, dt)).reshape(-1, 3)
    r = np.ascontiguousarray(np.asarray(rot, dt)).reshape(-1, 9)
    d = np.ascontiguousarray(np.asarray(dist2, dt)).reshape(-1)
    n = s.shape[0]
    out = np.zeros((n, 6), dt)
    fn = L.oracle_prep_cov3d_f32 if dt == np.float32 else L.oracle_prep_cov3d_f64
    fn(ctypes.c_int64(n), _p(s), _p(r), _p(d), _p(out))
    return out
