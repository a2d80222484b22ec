"""Whole-step CUDA graph capture.

The small configs (ACM/DBLP/IMDB shapes) are launch-bound: a step is ~70 kernel launches of a few
microseconds each, and when the large graph is sharded over 8 GPUs every kernel is 8x shorter while
the host work per launch is not.  Capturing forward + backward (+ the NCCL collectives, which torch
can capture) once and replaying it removes the host from the step entirely.

Everything the step does must be capture-safe: no host synchronisation (``.item()``), static input
tensors, every per-graph structure (transposed view, chunk tables, empty-row flags) built beforehand
-- ``GraphedStep`` runs the step eagerly a few times first, which warms all of those caches.
"""
from __future__ import annotations

from typing import Callable

import torch


class GraphedStep:
    """Captures ``fn()`` (which reads static tensors and returns tensors) into one CUDA graph.

        step = GraphedStep(lambda: train_step(X, graphs))   # warm-up + capture
        loss = step()                                        # replay; returns the same static outputs
    """

    def __init__(self, fn: Callable[[], object], warmup: int = 3):
        self.fn = fn
        try:   # the parameters' AccumulateGrad nodes predate the capture stream; that is intended here
            torch.autograd.graph.set_warn_on_accumulate_grad_stream_mismatch(False)
        except AttributeError:
            pass
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(warmup):
                fn()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.out = fn()
        torch.cuda.synchronize()

    def __call__(self):
        self.graph.replay()
        return self.out
