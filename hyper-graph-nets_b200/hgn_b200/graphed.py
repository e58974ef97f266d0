"""Whole-step CUDA-graph capture for the launch-bound small-mesh shapes.

A flag_simple-sized training step (1.6 k nodes x batch 21, 15 layers) is ~1 100 kernel launches of a few microseconds each;
eager PyTorch spends more time launching than the GPU spends computing.  ``GraphedStep`` captures one call of a step function
(forward, backward and -- if it is part of the function -- a capturable optimizer step) and replays it.  The library side is
capture-safe by construction: kernels run on the current (capturing) stream, workspaces come from torch's allocator (the graph's
private pool), segment plans are built during warm-up, and the weight-pack kernels are always recorded while capturing
(``ops._pack_weights``), so a replay re-reads the master weights an optimizer step inside the graph has changed.

    step = GraphedStep(lambda: loss_of(static_inputs), params)     # warm-up + capture
    static_inputs.copy_(batch); step()                              # replay

Inputs must live in static tensors that the caller overwrites in place between replays; gradients are (re)allocated inside the
graph's pool and stay valid until the next replay.
"""
from __future__ import annotations

from typing import Callable, Iterable, Optional

import torch

from . import _cabi


class GraphedStep:
    def __init__(self, step_fn: Callable[[], Optional[torch.Tensor]], params: Iterable[torch.nn.Parameter], warmup: int = 3):
        _cabi.require_cuda()
        self.params = list(params)
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(max(1, warmup)):          # lazy parameters, plans, kernel attributes, allocator warm-up
                step_fn()
        torch.cuda.current_stream().wait_stream(side)
        for p in self.params:
            p.grad = None                            # the captured backward allocates the gradients in the graph's pool
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.result = step_fn()

    def __call__(self):
        self.graph.replay()
        return self.result
