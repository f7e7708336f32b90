"""CUDA-graph capture of a whole training step.

At training sizes (batch 256 .. 4096) every kernel of the hot path runs for 10 .. 300 us and a step issues 60 .. 150
launches: an eager step is bound by launch overhead (2 .. 3 ms), not by any kernel (profiles/r01_kernel_table.txt).
Every libemk entry point is capturable by construction -- no synchronisation, stream-ordered scratch from a private
pool, no host read-back unless ``check_finite=True`` is asked for -- so a step (forward, backward, clipping, optimiser)
is captured ONCE and replayed with new batch contents; libemk's own NCCL collectives (``parallel.init_comm``) are
captured with it, which covers the data-parallel step as well.

This is the product-side home of what round 1 kept in ``tools/train_harness.py``.  The reference has no counterpart:
it relies on ``tf.function`` tracing of ``train_step`` (encodermap/models/models.py:3367-3401), which removes Python
overhead but still launches every kernel separately.
"""
from __future__ import annotations

from typing import Callable, Sequence

import torch


class GraphedStep:
    """``step_fn(*batch) -> tensor or tuple of tensors`` captured into one CUDA graph.

    ``example_batch`` fixes shapes and dtypes; its tensors are cloned into static buffers that every later ``__call__``
    copies the new batch into before replaying.  ``step_fn`` must be side-effect free on the host (the usual rules of
    graph capture): parameters and optimiser state are updated in place on the device, which replay repeats.  Optimisers
    need ``capturable=True`` (torch.optim.Adam) so that their step counter lives on the device.

    ``warmup`` eager iterations run on a side stream first (allocator warm-up, lazy initialisation of cuBLAS / NCCL / the
    per-device kernel attributes of libemk); they DO update the parameters, like ordinary training steps."""

    def __init__(self, step_fn: Callable, example_batch: Sequence[torch.Tensor], warmup: int = 3):
        self._fn = step_fn
        self._static = [b.clone() for b in example_batch]
        dev = self._static[0].device
        side = torch.cuda.Stream(dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            for _ in range(max(0, warmup)):
                step_fn(*self._static)
        torch.cuda.current_stream(dev).wait_stream(side)
        self._graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self._graph):
            self._out = step_fn(*self._static)

    @property
    def outputs(self):
        """static output tensor(s): valid after each replay until the next one"""
        return self._out

    def __call__(self, *batch: torch.Tensor):
        if len(batch) != len(self._static):
            raise ValueError(f"expected {len(self._static)} batch tensors, got {len(batch)}")
        for dst, src in zip(self._static, batch):
            if dst.shape != src.shape or dst.dtype != src.dtype:
                raise ValueError(f"batch tensor {tuple(src.shape)} {src.dtype} differs from the captured {tuple(dst.shape)} {dst.dtype}")
            dst.copy_(src, non_blocking=True)
        self._graph.replay()
        return self._out


def graphed_train_step(model: torch.nn.Module, loss_fn: Callable, optimizer: torch.optim.Optimizer, example_batch: Sequence[torch.Tensor],
                       clip_value: float | None = 1.0, grad_sync: Callable | None = None, warmup: int = 3) -> GraphedStep:
    """The step the reference's ``train_step`` performs (models/models.py:3367-3401, 2462-2521; optimiser
    ``Adam(lr, clipvalue=1.0)``, autoencoder/autoencoder.py:741-743) as one replayable graph:
    zero grads -> ``loss_fn(*batch)`` -> backward -> [``grad_sync(parameters)``: data-parallel averaging] -> clip by value ->
    ``optimizer.step()``.  Returns the GraphedStep; calling it yields the (static) detached loss tensor."""
    params = [p for p in model.parameters() if p.requires_grad]

    def step(*batch):
        optimizer.zero_grad(set_to_none=True)
        loss = loss_fn(*batch)
        loss.backward()
        if grad_sync is not None:
            grad_sync(params)
        if clip_value is not None:
            torch.nn.utils.clip_grad_value_(params, clip_value)
        optimizer.step()
        return loss.detach()

    return GraphedStep(step, example_batch, warmup)
