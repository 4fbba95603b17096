"""CUDA-graph capture of a whole render / training step (SURVEY 7 step 4 "no host sync" made it possible: the
binning never reads the intersection count back, all buffers are capacity-sized).

A step of the fused path is ~15 C-ABI launches plus autograd's bookkeeping; enqueued from Python that is ~0.3 ms of
host work per step, more than the 0.1 ms the binning takes on the device.  `CapturedStep` records the step once
and replays it with ONE host call: the launch-bound front of the step disappears and an end-to-end step (host
inputs in, host results out) costs the device time plus its copies.

Rules for the captured function (the usual CUDA-graph ones): it reads its inputs from tensors that live as long as
the capture and are UPDATED IN PLACE between replays (parameters by FusedAdam, cameras by ViewBatch.update_, targets
by copy_), it must not synchronise with the host, and the problem sizes (Gaussian count, image size, views) are
frozen -- after a densification capture again.
"""
from __future__ import annotations

from typing import Callable, Optional

import torch

from . import _lib, ops


class CapturedStep:
    """graph = CapturedStep(fn, device): runs fn() a few times on a side stream (allocator and binning capacities
    settle), captures one more run, and `replay()` re-issues it.  `outputs` is whatever fn returned during the
    capture (tensors at fixed addresses, rewritten by every replay)."""

    def __init__(self, fn: Callable[[], object], device, warmup: int = 3, capture_context=None, pool=None):
        """capture_context: optional context-manager factory entered around the captured run only (e.g. event
        timing of the captured launches); pool: memory pool of another capture to share (graph.pool())."""
        self.device = torch.device(device)
        cur = torch.cuda.current_stream(self.device)
        side = torch.cuda.Stream(device=self.device)
        side.wait_stream(cur)
        with torch.cuda.stream(side):
            for _ in range(max(0, warmup)):      # 0: fn must not run eagerly (e.g. a backward of a captured forward)
                fn()
        cur.wait_stream(side)
        torch.cuda.synchronize(self.device)
        ws = ops.workspace(self.device)
        ws.poll()                                   # an overflow of the warm-up runs would surface here
        self._slot_before = ws.info_next
        self.graph = torch.cuda.CUDAGraph()
        l0 = _lib.launch_count()
        import contextlib
        with torch.cuda.graph(self.graph, pool=pool):
            with (capture_context() if capture_context is not None else contextlib.nullcontext()) as self.context:
                self.outputs = fn()
        self.launches = _lib.launch_count() - l0     # kernels of this library that one replay launches
        # the binning calls of the captured run wrote their {M, overflow, longest} records into these pinned slots;
        # every replay rewrites them
        n_slots = (ws.info_next - self._slot_before) % ops._INFO_SLOTS
        self._slots = [(self._slot_before + i) % ops._INFO_SLOTS for i in range(n_slots)]
        self._ws = ws

    def pool(self):
        return self.graph.pool()

    def replay(self):
        self.graph.replay()
        return self.outputs

    def check(self) -> int:
        """Synchronise and look at the binning records of the last replay: raises if an intersection list outgrew
        the capacity frozen into the graph (capture again: the capacity has been raised); returns the intersection
        count of the last binning call."""
        torch.cuda.current_stream(self.device).synchronize()
        m = 0
        for s in self._slots:
            m, overflow, longest, _ = (int(x) for x in self._ws.info_host[s])
            if overflow:
                for key, cap in list(self._ws.capacity.items()):
                    self._ws.capacity[key] = max(cap, ops._capacity_for(m))
                raise _lib.GGError(f"captured step: {m} intersections exceed the capacity frozen into the graph; "
                                   "the images of this replay are invalid.  Capture the step again")
        return m
