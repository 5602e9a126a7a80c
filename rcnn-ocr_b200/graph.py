"""CUDA-graph replay of a whole hot-path step.

A cfg-B train step is ~70 launches (casts, packs, GEMMs, the two cooperative recurrent kernels per
block, the fused CTC kernel, the weight-gradient GEMMs, fused Adam).  Each launch through
Python + ctypes costs a few microseconds of host time; after a host sync (``loss.item()`` in the
reference's loop, training/train.py:507-508) the GPU runs out of queued work.  ``GraphedStep``
captures one invocation of a step function into a ``torch.cuda.CUDAGraph`` (every kernel of the
library is launched on torch's current stream and none allocates or synchronises, so the capture
needs nothing special; cooperative launches are captured as cooperative kernel nodes) and replays
it with fresh inputs copied into the captured (static) input tensors.

Requirements on ``fn``: shapes fixed across calls; optimizers inside must be ``capturable=True``;
no host synchronisation (``.item()``, ``.cpu()``) inside ``fn``.
"""
from __future__ import annotations

import torch


class GraphedStep:
    def __init__(self, fn, example_inputs, warmup: int = 3):
        """``fn(*tensors) -> tensor | tuple of tensors``; ``example_inputs`` are CUDA tensors
        whose shapes/dtypes every later call must match.  Runs ``warmup`` eager invocations on a
        side stream (lazy initialisation: library workspaces, optimizer state), then captures."""
        assert all(t.is_cuda for t in example_inputs), "GraphedStep needs CUDA tensors"
        self.static_in = [t.clone() for t in example_inputs]
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(max(1, warmup)):
                fn(*self.static_in)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.static_out = fn(*self.static_in)
        torch.cuda.synchronize()

    def replay(self):
        """Replay on whatever the static input tensors (``self.static_in``) hold -- for callers that stage the
        next inputs straight into them (e.g. H2D copies on a copy stream, two GraphedSteps used in turn)."""
        self.graph.replay()
        return self.static_out

    def __call__(self, *inputs):
        """Copy ``inputs`` (device, or pinned host) into the static tensors and replay.  The result
        tensors are overwritten by the next call: clone what must outlive it."""
        for dst, src in zip(self.static_in, inputs):
            if src is not dst:
                dst.copy_(src, non_blocking=True)
        self.graph.replay()
        return self.static_out
