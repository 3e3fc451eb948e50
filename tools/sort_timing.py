"""Per-tile phase times of the tile sorts for the bench workload (experiment build:
`SGR_NVCC_EXTRA=-DSGR_SORT_TIMING bash sigman_release_b200/csrc/build.sh build_variants/sorttiming.so`, then
`SGR_LIB_PATH=$PWD/build_variants/sorttiming.so python tools/sort_timing.py`)."""
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import numpy as np
import torch

from sigman_release_b200 import _native, scenes
from common import gpu_forward

VIEWS = [30, 37, 45, 53, 65, 85, 0, 8]
sc = scenes.body_gaussians(100_000, seed=0)
L = _native.lib()
M = 12
buf = np.zeros((4096, M), dtype=np.uint64)
cnt = ctypes.c_uint(0)
for it in range(3):
    gpu_forward(sc, VIEWS, 512, 512, requires_grad=True)
    torch.cuda.synchronize()
    L.sgr_debug_sort_marks(buf.ctypes.data_as(ctypes.c_void_p), ctypes.byref(cnt))
n = min(cnt.value, 4095)
print("instances with an empty / non-empty quarter mask (all 3 runs):", int(buf[4095, 0]), int(buf[4095, 1]))
rows = buf[:n].astype(np.int64)
t0 = rows[:, 0].min()
print("tiles timed", n, "span us", (rows[:, 9].max() - t0) / 1e3)
for thr in (1024, 256):
    sel = rows[rows[:, 11] == thr]
    if not len(sel):
        continue
    print(f"--- {thr}-thread CTAs: {len(sel)} tiles, keys {sel[:, 10].sum()}, first start {(sel[:, 0].min() - t0) / 1e3:.1f} us, "
          f"last end {(sel[:, 9].max() - t0) / 1e3:.1f} us")
    order = np.argsort(-sel[:, 10])
    for i in list(order[:6]) + list(order[len(order) // 2: len(order) // 2 + 3]):
        r = sel[i]
        d = [(r[k + 1] - r[k]) / 1e3 for k in range(9)]
        print(f"  n={r[10]:6d} start={(r[0] - t0) / 1e3:7.1f} total={(r[9] - r[0]) / 1e3:7.1f} us | " +
              " ".join(f"{nm}={x:.1f}" for nm, x in zip(["load+minmax", "hist", "scan", "scatter", "rank", "mask+count", "-", "blkscan", "blkemit"], d)))
    tot = [(sel[:, k + 1] - sel[:, k]).sum() / 1e3 for k in range(9)]
    print("  sum over tiles (us):", " ".join(f"{x:.0f}" for x in tot))
