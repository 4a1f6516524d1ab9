"""Training-step plumbing on B200: whole-step CUDA-graph capture.

At the reference's small per-GPU batches (16 clouds in pretrain.yaml / part segmentation) the training step is
launch-bound in eager mode: ~1200 kernel launches and ~650 dtype casts per step cost 23 ms of host time for 11 ms of
GPU work at config C3.  Every kernel of this library launches on the current stream with by-value tensor maps, the MAE
index maps need no host sync once ``n_vis`` is passed, and the only host-side randomness (the MAE mask, the HLT
tie-break noise - both drawn on the CPU by the reference too) enters through static device tensors, so forward +
backward + optimizer can be captured once and replayed.
"""

from __future__ import annotations

from typing import Callable

import torch

from .autograd import invalidate_param_cache


class GraphedStep:
    """Capture ``step_fn()`` (forward + backward [+ optimizer.step()]) into one CUDA graph after ``warmup`` eager runs on
    a side stream; ``replay()`` runs it.  ``step_fn`` must read its inputs from tensors that stay alive and are updated
    in place between replays (``.copy_``), must not sync with the host, and must zero gradients with
    ``set_to_none=True`` BEFORE capture only (gradients are then static graph outputs).  Optimizers need
    ``capturable=True``."""

    def __init__(self, step_fn: Callable[[], torch.Tensor], warmup: int = 3):
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(warmup):
                step_fn()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.output = step_fn()

    def replay(self):
        self.graph.replay()
        # the replay updated parameters / BatchNorm statistics in place without bumping their version counters
        invalidate_param_cache()
        return self.output
