"""View-sharded data parallelism over one 8xB200 box.

The Gaussian parameters are replicated on every rank; camera views are partitioned by rank
(view v -> rank v mod G); every rank renders its own views with no data-path collective.  For a
training step the leaf gradients of the local views are summed across ranks with ONE all-reduce
over a flat, persistently allocated fp32 buffer (NCCL over NVLink 5 / NVSwitch; gloo in the CPU
tests).  Forward-only render batches need no collective at all.

The reference has no working multi-GPU path for this model (its DDP wrapper cannot survive the
densification's parameter replacement); this module is what BASELINE.json's
north_star asks for instead.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence

import torch
import torch.distributed as dist

PARAM_ORDER = ("means", "log_scales", "quats", "opacity_logit", "sh_coeffs", "features")


def views_for_rank(n_views: int, rank: int, world: int) -> List[int]:
    """Round-robin partition: global view v is rendered by rank v mod world."""
    if world < 1 or not 0 <= rank < world:
        raise ValueError(f"bad rank/world {rank}/{world}")
    return list(range(rank, n_views, world))


def owner_of_view(view: int, world: int) -> int:
    return view % world


class GradientBucket:
    """One flat fp32 buffer holding every leaf gradient, all-reduced in a single collective.

    Layout: parameters in PARAM_ORDER, each flattened row-major; N x (3+3+4+1+3K+D) floats
    (408 MB at N = 1M, K = 25, D = 16)."""

    def __init__(self, params: Dict[str, torch.Tensor], group: Optional[dist.ProcessGroup] = None):
        self.names = [k for k in PARAM_ORDER if k in params]
        extra = [k for k in params if k not in PARAM_ORDER]
        if extra:
            raise ValueError(f"unknown parameter names {extra}")
        self.shapes = {k: tuple(params[k].shape) for k in self.names}
        self.sizes = {k: int(params[k].numel()) for k in self.names}
        self.offsets = {}
        off = 0
        for k in self.names:
            self.offsets[k] = off
            off += self.sizes[k]
        self.numel = off
        p0 = params[self.names[0]]
        self.flat = torch.zeros(off, dtype=torch.float32, device=p0.device)
        self.group = group

    def view(self, name: str) -> torch.Tensor:
        o = self.offsets[name]
        return self.flat[o:o + self.sizes[name]].view(self.shapes[name])

    def pack(self, grads: Dict[str, Optional[torch.Tensor]]) -> torch.Tensor:
        """Copy (or zero, for a missing gradient) every gradient into the flat buffer."""
        for k in self.names:
            g = grads.get(k)
            if g is None:
                self.view(k).zero_()
            else:
                self.view(k).copy_(g)
        return self.flat

    def all_reduce(self, async_op: bool = False):
        """Sum over ranks (a no-op without an initialised process group)."""
        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(self.group) == 1:
            return None
        return dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=self.group, async_op=async_op)

    def unpack(self) -> Dict[str, torch.Tensor]:
        return {k: self.view(k) for k in self.names}


def all_reduce_gradients(params: Dict[str, torch.Tensor], bucket: Optional[GradientBucket] = None,
                         group: Optional[dist.ProcessGroup] = None) -> GradientBucket:
    """Sum `.grad` of every parameter over the ranks and write the result back into `.grad`."""
    if bucket is None:
        bucket = GradientBucket(params, group)
    bucket.pack({k: params[k].grad for k in bucket.names})
    bucket.all_reduce()
    for k, g in bucket.unpack().items():
        if params[k].grad is None:
            params[k].grad = g.clone()
        else:
            params[k].grad.copy_(g)
    return bucket


def render_views_sharded(params: Dict[str, torch.Tensor], cameras: Sequence, device, **kw):
    """Render this rank's share of `cameras` (a list indexed by global view id).

    Returns (local_view_ids, outputs of render_views for those views, in that order)."""
    from .render import ViewBatch, render_views
    world = dist.get_world_size() if dist.is_initialized() else 1
    rank = dist.get_rank() if dist.is_initialized() else 0
    mine = views_for_rank(len(cameras), rank, world)
    if not mine:
        return mine, None
    batch = ViewBatch.from_cameras([cameras[v] for v in mine], device)
    out = render_views(params["means"], params["log_scales"], params["quats"], params["opacity_logit"],
                       params["sh_coeffs"], params["features"], batch, **kw)
    return mine, out
