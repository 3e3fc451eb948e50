"""Per-stage device times of forward + backward for a given number of views of the bench subject:
`python tools/stage_times.py [views]` (CUDA events around every stage inside the library)."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import torch
from sigman_release_b200 import _native, scenes
from common import gpu_forward

VIEWS = [30, 37, 45, 53, 65, 85, 0, 8]
nv = int(sys.argv[1]) if len(sys.argv) > 1 else 8
sc = scenes.body_gaussians(100_000, seed=0)
L = _native.lib()


def step():
    out, t, _ = gpu_forward(sc, VIEWS[:nv], 512, 512, requires_grad=True)
    (out[0].clamp(0, 1) - 0.5).abs().mean().backward()


for _ in range(5):
    step()
torch.cuda.synchronize()
L.sgr_profile_enable(1)
for _ in range(20):
    step()
torch.cuda.synchronize()
ms = (ctypes.c_double * len(_native.STAGES))(); cnt = (ctypes.c_uint32 * len(_native.STAGES))()
L.sgr_profile_collect(ms, cnt)
L.sgr_profile_enable(0)
st = {k: round(ms[i] / max(cnt[i], 1) * 1e3, 1) for i, k in enumerate(_native.STAGES) if cnt[i]}
print(f"{nv} view(s):", st, "sum", round(sum(st.values()), 1), "us")
