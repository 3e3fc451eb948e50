"""Load-balance diagnostics: per-tile (list length, start, duration) of the forward blend for the bench workload."""
import ctypes
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from sigman_release_b200 import _native, scenes
from common import gpu_forward, saved_state

VIEWS = [30, 37, 45, 53, 65, 85, 0, 8]
sc = scenes.body_gaussians(100_000, seed=0)
for it in range(3):
    out, t, _ = gpu_forward(sc, VIEWS, 512, 512, requires_grad=True)
torch.cuda.synchronize()
state, dims = saved_state(out[0])
B, V, N, H, W = dims[:5]; cap, cap_b, flags = dims[7]
L = _native.lib()
T = 1024
rows = []
for r in range(V):
    ranges = torch.zeros((T, 2), dtype=torch.int32, device="cuda")
    tt = torch.zeros((T, 2), dtype=torch.int32, device="cuda")
    st = torch.cuda.current_stream()
    _native.check(L.sgr_debug_copy_state(ctypes.c_void_p(state.data_ptr()), B, V, N, H, W, int(cap), int(cap_b), int(flags), r,
                                         ctypes.c_void_p(ranges.data_ptr()), None, None, 0,
                                         ctypes.c_void_p(tt.data_ptr()), ctypes.c_void_p(st.cuda_stream)))
    torch.cuda.synchronize()
    rg = ranges.cpu().numpy().astype(np.int64); tm = tt.cpu().numpy().astype(np.uint32).astype(np.int64)
    for k in range(T):
        n = rg[k, 1] - rg[k, 0]
        if n > 0:
            rows.append((r, k, n, tm[k, 0], tm[k, 1]))
rows = np.array(rows)
t0 = rows[:, 3].min()
end = (rows[:, 3] + rows[:, 4]).max()
print("tiles", len(rows), "kernel span us", (end - t0) / 1e3, "sum tile us", rows[:, 4].sum() / 1e3)
order = np.argsort(-rows[:, 4])
print("top tiles by duration: render tile n start_us dur_us ns/entry")
for i in order[:25]:
    r, k, n, s, d = rows[i]
    print(f"  r{r} t{k:4d} n={n:6d} start={(s - t0) / 1e3:8.1f} dur={d / 1e3:8.1f} ns/entry={d / n:7.1f}")
print("by size class: n range, count, mean dur us, mean ns/entry")
for lo, hi in [(1, 64), (64, 256), (256, 1024), (1024, 4096), (4096, 1 << 20)]:
    m = (rows[:, 2] >= lo) & (rows[:, 2] < hi)
    if m.any():
        print(f"  [{lo},{hi}) cnt={m.sum()} dur={rows[m, 4].mean() / 1e3:.1f} ns/entry={(rows[m, 4] / rows[m, 2]).mean():.1f}")
late = order[np.argsort(-(rows[order, 3] + rows[order, 4]))][:10]
print("last finishing tiles:")
for i in late:
    r, k, n, s, d = rows[i]
    print(f"  r{r} t{k:4d} n={n:6d} start={(s - t0) / 1e3:8.1f} end={(s + d - t0) / 1e3:8.1f}")
