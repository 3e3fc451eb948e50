"""Experiment: per-phase cycle counters of the forward blend (needs a library built with -DSGR_PHASE_TIMING)."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from sigman_release_b200 import _native, scenes
from common import gpu_forward
VIEWS = [30, 37, 45, 53, 65, 85, 0, 8]
sc = scenes.body_gaussians(100_000, seed=0)
L = _native.lib()
out = (ctypes.c_ulonglong * 16)()
for it in range(3):
    gpu_forward(sc, VIEWS, 512, 512, requires_grad=True)
L.sgr_debug_phase_counters(out, 1)
gpu_forward(sc, VIEWS, 512, 512, requires_grad=True)
L.sgr_debug_phase_counters(out, 1)
v = list(out)
names = ["tma wait", "cull", "trips", "refine+ckpt"]
tot = sum(v[:4])
print("quarter survivors (sum over quarters)", v[6], "trips (max over quarters)", v[4], "mean/max", v[6] / 4 / max(v[4], 1))
print("all items: cycles", {n: f"{x/1e6:.1f}M ({100*x/tot:.0f}%)" for n, x in zip(names, v[:4])}, "trips", v[4], "batches", v[5],
      f"cycles/trip {v[2]/max(v[4],1):.0f} cull cycles/batch {v[1]/max(v[5],1):.0f} refine/batch {v[3]/max(v[5],1):.0f} wait/batch {v[0]/max(v[5],1):.0f}")
tot = sum(v[8:12])
print("longest tile blk 7: n", v[14], "batches", v[13], "trips", v[12], {n: f"{x/1e3:.0f}K ({100*x/max(tot,1):.0f}%)" for n, x in zip(names, v[8:12])},
      f"cycles/trip {v[10]/max(v[12],1):.0f} cull/batch {v[9]/max(v[13],1):.0f} refine/batch {v[11]/max(v[13],1):.0f} wait/batch {v[8]/max(v[13],1):.0f}")

# backward (the same counters are reused: run one backward after a reset)
out_t, t_, _ = gpu_forward(sc, VIEWS, 512, 512, requires_grad=True)
torch.cuda.synchronize()
L.sgr_debug_phase_counters(out, 1)
(out_t[0].clamp(0, 1) - 0.5).abs().mean().backward()
L.sgr_debug_phase_counters(out, 1)
v = list(out)
tot = sum(v[:4])
print("backward, all items: cycles", {n: f"{x/1e6:.1f}M ({100*x/tot:.0f}%)" for n, x in zip(["tma wait", "cull", "phases A+B", "phase C"], v[:4])},
      "trips", v[4], "batches", v[5], f"A+B cycles/trip {v[2]/max(v[4],1):.0f}  C cycles/trip {v[3]/max(v[4],1):.0f}  cull/batch {v[1]/max(v[5],1):.0f}  wait/batch {v[0]/max(v[5],1):.0f}  trips/batch {v[4]/max(v[5],1):.1f}")
