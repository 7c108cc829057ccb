"""CUDA-graph capture of a whole loss step (forward + backward).

The FCAM step is a chain of ~25 short kernels (word-region forward / backward, the B x B cross entropies, the
sentence-loss products).  At B = 128 the device needs ~0.75 ms for it while the Python / autograd launch path
needs ~0.9 ms, so an eager step is host bound.  `GraphedStep` captures one call of a step function -- every
kernel of libtgfr_b200.so is launched on the caller's stream and none of them synchronises or allocates, so the
whole chain is capturable -- and replays it with a single launch.

    step = GraphedStep(lambda: compute_losses_and_backward(static_inputs))
    static_inputs[0].copy_(new_batch)          # refresh the static input buffers in place
    loss = step()                              # replay; gradients land in the static .grad tensors

The callable must read only tensors that stay alive and in place between replays (the usual CUDA-graph
contract) and must include the backward pass if gradients are wanted.
"""
from __future__ import annotations

import torch

__all__ = ["GraphedStep"]


class GraphedStep:
    def __init__(self, fn, warmup: int = 3, pool=None):
        """pool: a torch.cuda.graph_pool_handle() shared by graphs that are never replayed concurrently (their
        private allocations -- e.g. the forward->backward attention records -- then reuse the same memory)."""
        self._fn = fn
        dev = torch.cuda.current_device()
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):                      # warm-up off the capturing stream (allocator, lazy inits)
            for _ in range(max(1, warmup)):
                fn()
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph, pool=pool):
            self.out = fn()

    def __call__(self):
        self.graph.replay()
        return self.out
