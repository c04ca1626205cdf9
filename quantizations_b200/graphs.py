"""CUDA-graph helper: capture a callable that runs Linear4bit / core.* calls and replay it with one launch.

The reference launches every kernel on the legacy default stream (ops.cu:170), which cannot be captured; every kernel of
this library takes the current torch stream, so a whole decode step (hundreds of GEMVs) can be one graph launch.
"""
from __future__ import annotations

import torch


class CapturedStep:
    def __init__(self, graph: torch.cuda.CUDAGraph, result):
        self.graph = graph
        self.result = result  # whatever the callable returned (static output tensors)

    def replay(self):
        self.graph.replay()
        return self.result


def capture(fn, warmup: int = 2) -> CapturedStep:
    """Warm `fn` up on a side stream, capture one call into a CUDA graph, return a replayable handle.
    Inputs must live in static tensors that the caller overwrites in place between replays."""
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for _ in range(warmup):
            fn()
    torch.cuda.current_stream().wait_stream(s)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        result = fn()
    return CapturedStep(g, result)
