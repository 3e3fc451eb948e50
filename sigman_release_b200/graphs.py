"""CUDA-graph capture of a whole rasteriser step (forward + backward).

The launch sequence of ``libsgr_b200.so`` depends on host-side shapes only — no device->host read and no blocking wait
sits on it — so a training or rendering step can be captured once and replayed: ~20 launches and their Python /
autograd dispatch collapse into one ``cudaGraphLaunch`` (host cost per step: tens of microseconds instead of
~0.3 ms; matters when several ranks share the host cores, and removes the 1-3 us gaps between kernels).

    def fn():
        for leaf in static_leaves: leaf.grad = None          # so that backward assigns the .grad tensors
        loss = render_l1_loss(*static_inputs, target)[0]; loss.backward(); return loss
    step = GraphedStep(fn)
    ...
    static_means.copy_(new_means)        # refresh the static inputs in place
    step.replay()                        # gradients land in the (static) .grad tensors of the captured leaves

Rules (those of ``torch.cuda.graph``): the callable must run at least once eagerly before the capture (``GraphedStep``
does that: it also sizes the instance buffers), the tensors it reads and the ``.grad`` tensors it writes are static,
and shapes must not change.  The instance capacity is frozen at capture time; ``rasterizer.graph_status()`` reports an
overflow of the latest replay.
"""
from __future__ import annotations

from typing import Callable

import torch

__all__ = ["GraphedStep"]


class GraphedStep:
    def __init__(self, fn: Callable[[], object], warmup: int = 3, device=None):
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        # the warm-up runs on a side stream on purpose while the leaves were created on the caller's stream: silence
        # autograd's stream-mismatch warning for the warm-up only and restore the setting afterwards
        setter = getattr(torch.autograd.graph, "set_warn_on_accumulate_grad_stream_mismatch", None)
        if setter is not None:
            setter(False)
        try:
            side = torch.cuda.Stream(device=self.device)
            side.wait_stream(torch.cuda.current_stream(self.device))
            with torch.cuda.stream(side):                   # eager warm-up on a side stream (torch's capture recipe)
                for _ in range(max(1, warmup)):
                    fn()
            torch.cuda.current_stream(self.device).wait_stream(side)
            torch.cuda.synchronize(self.device)
        finally:
            if setter is not None:
                setter(True)                                # (the default)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.result = fn()

    def replay(self):
        self.graph.replay()
        return self.result
