"""CUDA-graph replay of one forward+backward step of a module built from STC cells.

At the reference's own batch size (32) one SF-shape training step is ~400 kernel launches of a few microseconds
each (24 cell steps x (2 convolutions + 3 support products) forward, twice that backward): the step is bound by
launch latency, not by the kernels.  Every launch of libstc_b200.so goes to the current
stream with no host synchronisation and no allocation inside the library, so the whole step can be captured once
and replayed: inputs are copied into static buffers, gradients land in static tensors.
"""
from __future__ import annotations

from typing import Callable, Sequence

import torch


class GraphedStep:
    """Capture ``loss = loss_fn(*static_inputs); loss.backward()`` once; ``replay(*inputs)`` copies new inputs in.

    ``params`` are the leaves whose ``.grad`` the step produces (their ``.grad`` tensors stay the same objects
    across replays, as torch.cuda.graphs requires).  ``warmup`` eager iterations run on a side stream first.
    """

    def __init__(self, loss_fn: Callable[..., torch.Tensor], example_inputs: Sequence[torch.Tensor],
                 params: Sequence[torch.Tensor], warmup: int = 3):
        if not torch.cuda.is_available():
            raise RuntimeError("GraphedStep needs a CUDA device (there is no CPU path)")
        self.static_inputs = [t.clone() for t in example_inputs]
        self.params = list(params)
        self.loss_fn = loss_fn
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(max(1, warmup)):
                self._zero()
                loss_fn(*self.static_inputs).backward()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self._zero()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.loss = loss_fn(*self.static_inputs)
            self.loss.backward()
        torch.cuda.synchronize()

    def _zero(self):
        for p in self.params:
            p.grad = None

    def replay(self, *inputs: torch.Tensor) -> torch.Tensor:
        for dst, src in zip(self.static_inputs, inputs):
            if src is not dst:
                dst.copy_(src, non_blocking=True)
        self.graph.replay()
        return self.loss
