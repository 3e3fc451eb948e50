"""ctypes binding of ``libsgr_b200.so`` (C ABI declared in ``include/sgr.h``).

The library is built in-tree by ``sigman_release_b200/csrc/build.sh`` (``__graft_entry__.build()``).  There is no
CPU or PyTorch fallback: if the library is missing or a CUDA device is absent, the calls raise.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("SGR_LIB_PATH") or os.path.join(_HERE, "libsgr_b200.so")
ABI_VERSION = 3

SGR_OK = 0
SGR_E_INVALID_ARGUMENT = -1
SGR_E_BUFFER_TOO_SMALL = -2
SGR_E_CUDA = -3
SGR_E_INSTANCE_OVERFLOW = -4

FLAG_SIMPLE_BLEND = 1
FLAG_CLAMP_COLOR = 2
FLAG_FORWARD_ONLY = 4
FLAG_TILE_TIMING = 8
FLAG_EXACT_EXP = 16

_vp = ctypes.c_void_p


class SgrProblem(ctypes.Structure):
    _fields_ = [
        ("num_subjects", ctypes.c_int32),
        ("views_per_subject", ctypes.c_int32),
        ("num_gaussians", ctypes.c_int32),
        ("image_height", ctypes.c_int32),
        ("image_width", ctypes.c_int32),
        ("tanfovx", ctypes.c_float),
        ("tanfovy", ctypes.c_float),
        ("means3D", _vp),
        ("cov3D", _vp),
        ("colors", _vp),
        ("opacities", _vp),
        ("viewmatrix", _vp),
        ("projmatrix", _vp),
        ("bg", _vp),
        ("max_instances", ctypes.c_uint64),
        ("max_block_records", ctypes.c_uint64),
        ("renders_per_chunk", ctypes.c_int32),
        ("flags", ctypes.c_int32),
        ("max_tile_instances_hint", ctypes.c_uint32),
        ("reserved", ctypes.c_uint32),
    ]


class SgrForwardArgs(ctypes.Structure):
    _fields_ = [
        ("p", SgrProblem),
        ("out_color", _vp),
        ("out_depth", _vp),
        ("out_alpha", _vp),
        ("radii", _vp),
        ("state", _vp),
        ("state_bytes", ctypes.c_uint64),
        ("scratch", _vp),
        ("scratch_bytes", ctypes.c_uint64),
        ("stream", _vp),
        ("loss_target", _vp),
        ("loss_mask", _vp),
        ("loss_dL_dcolor", _vp),
        ("loss_out", _vp),
        ("loss_scale", ctypes.c_float),
        ("out_lpips_feed", _vp),
    ]


class SgrBackwardArgs(ctypes.Structure):
    _fields_ = [
        ("p", SgrProblem),
        ("out_alpha", _vp),
        ("radii", _vp),
        ("dL_dcolor", _vp),
        ("dL_ddepth", _vp),
        ("dL_dalpha", _vp),
        ("dL_dmeans3D", _vp),
        ("dL_dcov3D", _vp),
        ("dL_dcolors", _vp),
        ("dL_dopacities", _vp),
        ("dL_dmeans2D", _vp),
        ("state", _vp),
        ("state_bytes", ctypes.c_uint64),
        ("scratch", _vp),
        ("scratch_bytes", ctypes.c_uint64),
        ("stream", _vp),
        ("loss_dL_dcolor", _vp),
        ("dL_dcolor_scale", _vp),
        ("dL_dlpips_feed", _vp),
        ("fused_clamp", ctypes.c_int32),
        ("reserved", ctypes.c_int32),
    ]


class SgrStatus(ctypes.Structure):
    _fields_ = [
        ("instances_required", ctypes.c_uint64),
        ("instances_capacity", ctypes.c_uint64),
        ("overflow", ctypes.c_uint32),
        ("max_tile_instances", ctypes.c_uint32),
        ("nonempty_tiles", ctypes.c_uint32),
        ("reserved", ctypes.c_uint32),
        ("block_records_required", ctypes.c_uint64),
        ("block_records_capacity", ctypes.c_uint64),
    ]


# every symbol include/sgr.h declares: name -> (restype, argtypes)
_i32, _u64, _f32 = ctypes.c_int32, ctypes.c_uint64, ctypes.c_float
SYMBOLS = {
    "sgr_abi_version": (ctypes.c_int, []),
    "sgr_last_error": (ctypes.c_char_p, []),
    "sgr_state_bytes": (_u64, [_i32, _i32, _i32, _i32, _i32, _u64, _u64, _i32]),
    "sgr_scratch_bytes": (_u64, [_i32, _i32, _i32, _i32, _i32, _u64, _i32]),
    "sgr_forward": (ctypes.c_int, [ctypes.POINTER(SgrForwardArgs)]),
    "sgr_backward": (ctypes.c_int, [ctypes.POINTER(SgrBackwardArgs)]),
    "sgr_read_status": (ctypes.c_int, [_vp, _vp, ctypes.POINTER(SgrStatus)]),
    "sgr_mark_visible": (ctypes.c_int, [_vp, _i32, _vp, _vp, _vp, _vp]),
    "sgr_cov3d_from_scale_rot": (ctypes.c_int, [_vp, _vp, _f32, _i32, _vp, _vp]),
    "sgr_cov3d_from_scale_rot_backward": (ctypes.c_int, [_vp, _vp, _f32, _i32, _vp, _vp, _vp, _vp]),
    "sgr_prep_cov3d": (ctypes.c_int, [_vp, _vp, _vp, ctypes.c_int64, _i32, _vp, _vp]),
    "sgr_prep_cov3d_backward": (ctypes.c_int, [_vp, _vp, _vp, ctypes.c_int64, _i32, _vp, _vp, _vp, _vp]),
    "sgr_sh_colors": (ctypes.c_int, [_vp, _vp, _vp, _i32, _i32, _i32, _vp, _vp, _vp]),
    "sgr_sh_colors_backward": (ctypes.c_int, [_vp, _vp, _vp, _i32, _i32, _i32, _vp, _vp, _vp, _vp, _vp]),
    "sgr_knn_scratch_bytes": (_u64, [_i32]),
    "sgr_knn_mean_dist2": (ctypes.c_int, [_vp, _i32, _vp, _vp, _u64, _vp]),
    "sgr_knn_scratch_bytes_batched": (_u64, [_i32, _i32]),
    "sgr_knn_mean_dist2_batched": (ctypes.c_int, [_vp, _i32, _i32, _vp, _vp, _u64, _vp]),
    "sgr_wire_chunk_bytes": (_u64, [_i32, ctypes.c_int64, _i32]),
    "sgr_wire_pack": (ctypes.c_int, [_vp, _vp, _vp, _i32, ctypes.c_int64, _vp, _vp]),
    "sgr_wire_unpack": (ctypes.c_int, [_vp, _i32, ctypes.c_int64, _i32, ctypes.c_int64, _i32, _vp, ctypes.c_int64,
                                       ctypes.c_int64, _vp]),
    "sgr_profile_enable": (None, [ctypes.c_int]),
    "sgr_profile_collect": (ctypes.c_int, [ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_uint32)]),
    "sgr_launch_count": (_u64, []),
    "sgr_debug_copy_state": (ctypes.c_int, [_vp, _i32, _i32, _i32, _i32, _i32, _u64, _u64, _i32, _i32, _vp, _vp, _vp, _u64,
                                            _vp, _vp]),
}

STAGES = ("preprocess", "plan", "scatter", "sort", "blend_forward", "blend_backward", "preprocess_backward")

_LIB = None


class SgrError(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__(f"libsgr_b200 error {code}: {message}")
        self.code = code


def build(force: bool = False) -> str:
    """Compile ``libsgr_b200.so`` for sm_100a with nvcc (cross-compiles without a GPU)."""
    src_dir = os.path.join(_HERE, "csrc")
    srcs = [os.path.join(src_dir, f) for f in sorted(os.listdir(src_dir)) if f.endswith((".cu", ".cuh", ".sh"))]
    srcs.append(os.path.join(os.path.dirname(_HERE), "include", "sgr.h"))
    if not force and os.path.exists(LIB_PATH) and all(os.path.getmtime(LIB_PATH) >= os.path.getmtime(s) for s in srcs):
        return LIB_PATH
    subprocess.check_call(["bash", os.path.join(src_dir, "build.sh"), LIB_PATH])
    return LIB_PATH


def lib() -> ctypes.CDLL:
    """The loaded library; raises if it has not been built (no fallback path exists)."""
    global _LIB
    if _LIB is not None:
        return _LIB
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(or sigman_release_b200/csrc/build.sh).  sigman_release_b200 has no CPU/PyTorch fallback."
        )
    L = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(L, name)          # AttributeError if the library lacks a declared symbol
        fn.restype = res
        fn.argtypes = args
    if L.sgr_abi_version() != ABI_VERSION:
        raise ImportError(f"libsgr_b200.so ABI {L.sgr_abi_version()} != expected {ABI_VERSION}; rebuild")
    _LIB = L
    return L


def check(rc: int) -> None:
    if rc != SGR_OK:
        raise SgrError(rc, lib().sgr_last_error().decode("utf-8", "replace"))
