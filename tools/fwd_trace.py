"""Per-item trace of the forward blend for the bench workload (experiment build:
`SGR_NVCC_EXTRA=-DSGR_FWD_TRACE bash sigman_release_b200/csrc/build.sh build_variants/fwdtrace.so`, then
`SGR_LIB_PATH=$PWD/build_variants/fwdtrace.so python tools/fwd_trace.py`): how many warps are busy over time, which
items finish last, duration against list length."""
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import numpy as np
import torch

from sigman_release_b200 import _native, rasterizer, scenes
from common import gpu_forward

VIEWS = [30, 37, 45, 53, 65, 85, 0, 8]
sc = scenes.body_gaussians(100_000, seed=0)
L = _native.lib()
FUSED = len(sys.argv) > 1 and sys.argv[1] == "loss"      # the bench's forward: fused clamp + L1 loss epilogue
if FUSED:
    from sigman_release_b200 import cameras
    from common import TAN, to_dev
    t = dict(means3D=to_dev(sc["means3D"])[None], cov3D=to_dev(sc["cov3D"])[None], colors=to_dev(sc["colors"])[None],
             opacities=to_dev(sc["opacities"]).reshape(1, -1))
    for v in t.values():
        v.requires_grad_(True)
    vm, pm, _ = cameras.orbit_cameras(VIEWS)
    vmt, pmt, bg = to_dev(vm)[None], to_dev(pm)[None], torch.ones(3, device="cuda")
    target = torch.rand((1, len(VIEWS), 3, 512, 512), device="cuda")

    def forward():
        rasterizer.render_l1_loss(t["means3D"], t["cov3D"], t["colors"], t["opacities"], vmt, pmt, bg, 512, 512, TAN, TAN, target)
else:
    def forward():
        gpu_forward(sc, VIEWS, 512, 512, requires_grad=True)
for it in range(3):
    forward()
torch.cuda.synchronize()
L.sgr_profile_enable(1)
for it in range(5):
    forward()
torch.cuda.synchronize()
ms = (ctypes.c_double * len(_native.STAGES))(); cnt = (ctypes.c_uint32 * len(_native.STAGES))()
L.sgr_profile_collect(ms, cnt)
L.sgr_profile_enable(0)
print("stage times (us):", {k: round(ms[i] / max(cnt[i], 1) * 1e3, 1) for i, k in enumerate(_native.STAGES) if cnt[i]})
n_items = len(VIEWS) * 1024 * 8                 # rows are indexed by (render, tile, block)
buf = np.zeros((n_items, 4), dtype=np.uint64)
L.sgr_debug_fwd_trace(buf.ctypes.data_as(ctypes.c_void_p), ctypes.c_uint(n_items))
rows = buf.astype(np.int64)
rows = rows[rows[:, 1] > 0]
t0 = rows[:, 0].min()
beg, end, n, sm = (rows[:, 0] - t0) / 1e3, (rows[:, 1] - t0) / 1e3, rows[:, 2], rows[:, 3] >> 8
span = end.max()
print(f"items {len(rows)}, span {span:.1f} us, sum of item durations {np.sum(end - beg) / 1e3:.1f} ms "
      f"(= {np.sum(end - beg) / span / 2368 * 100:.0f} % of 2368 warp slots)")
print("busy warps over time (items in flight):")
for t in np.linspace(0, span, 17)[:-1]:
    busy = int(((beg <= t) & (end > t)).sum())
    sms = len(np.unique(sm[(beg <= t) & (end > t)]))
    print(f"  t={t:6.1f} us  busy warps {busy:5d}  SMs with work {sms:4d}")
order = np.argsort(-end)
print("last finishing items: n, begin, end, duration, ns/entry")
for i in order[:12]:
    print(f"  n={n[i]:5d} begin={beg[i]:7.1f} end={end[i]:7.1f} dur={end[i] - beg[i]:7.1f} ns/entry={(end[i] - beg[i]) * 1e3 / max(n[i], 1):6.1f}")
print("longest items: n, begin, end, duration")
for i in np.argsort(-(end - beg))[:12]:
    print(f"  n={n[i]:5d} begin={beg[i]:7.1f} end={end[i]:7.1f} dur={end[i] - beg[i]:7.1f} ns/entry={(end[i] - beg[i]) * 1e3 / max(n[i], 1):6.1f}")
print("by list length: range, items, total entries, mean duration us, ns/entry")
for lo, hi in [(0, 1), (1, 32), (32, 128), (128, 512), (512, 1024), (1024, 2048), (2048, 1 << 20)]:
    m = (n >= lo) & (n < hi)
    if m.any():
        print(f"  [{lo},{hi}) items={int(m.sum()):6d} entries={int(n[m].sum()):8d} mean dur={np.mean((end - beg)[m]):7.2f} "
              f"ns/entry={np.sum((end - beg)[m]) * 1e3 / max(n[m].sum(), 1):6.1f}")
