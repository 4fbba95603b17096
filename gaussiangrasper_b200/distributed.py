"""View-sharded data parallelism over one 8xB200 box.

The Gaussian parameters are replicated on every rank; camera views are partitioned by rank
(view v -> rank v mod G); every rank renders its own views with no data-path collective.  For a
training step the leaf gradients of the local views are summed across ranks with ONE all-reduce
over a flat, persistently allocated fp32 buffer (NCCL over NVLink 5 / NVSwitch; gloo in the CPU
tests).  Forward-only render batches need no collective at all.  `FactoredExchange` cuts the bytes of
that step 2x by exchanging the SH gradient as its per-view factors (an all-gather of 3 floats per view
and Gaussian) instead of the 75-float product.

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

    Layout: parameters in PARAM_ORDER, each flattened row-major and starting on a 16-byte boundary;
    N x (3+3+4+1+3K+D) floats of payload (408 MB at N = 1M, K = 25, D = 16) plus at most 3 padding
    floats per segment (they stay zero)."""

    def __init__(self, params: Dict[str, torch.Tensor], group: Optional[dist.ProcessGroup] = None, allocator=None):
        """allocator(numel) -> zero-filled fp32 CUDA tensor of at least numel elements (e.g. symmetric memory);
        default torch.zeros."""
        self.names = [k for k in PARAM_ORDER if k in params]
        extra = [k for k in params if k not in PARAM_ORDER]
        if extra:
            raise ValueError(f"unknown parameter names {extra}")
        self.shapes = {k: tuple(params[k].shape) for k in self.names}
        self.sizes = {k: int(params[k].numel()) for k in self.names}
        self.offsets = {}
        off = 0
        for k in self.names:
            off = (off + 3) // 4 * 4   # every segment starts 16-byte aligned (the backward stores rows in bulk)
            self.offsets[k] = off
            off += self.sizes[k]
        self.numel = off
        self.payload = sum(self.sizes.values())
        p0 = params[self.names[0]]
        if allocator is None:
            self.flat = torch.zeros(off, dtype=torch.float32, device=p0.device)
        else:
            self.flat = allocator(off)
            if self.flat.numel() < off or self.flat.dtype != torch.float32 or not self.flat.is_contiguous():
                raise ValueError("GradientBucket allocator must return a contiguous fp32 tensor of >= numel elements")
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


class FactoredExchange:
    """Gradient exchange of a view-sharded training step that sends the SH gradient as its factors.

    The gradient of a Gaussian's [K,3] SH coefficients is sum_v Y(dir_v) (x) v_rgb_v -- one outer
    product per view -- so instead of all-reducing 3K floats per Gaussian (75 of the 102 at degree 4,
    D = 16) the ranks all-gather the 3-float colour gradient of each of their views plus the camera
    centres, and every rank rebuilds the sum locally (`gg_sh_grad_from_views`).  The remaining leaf
    gradients (N x (11 + D) floats) go through one all-reduce as before.  Per-rank traffic at 8 ranks,
    N = 500k: 54 MB all-reduced + 48 MB gathered instead of 204 MB all-reduced; the rebuild kernel runs
    while the all-reduce is in flight.  Results equal the plain all-reduce up to reassociation.

    Usage per step:  holder = ex.holder();  render_views(..., views, holder=holder);  loss.backward();
                     grads = ex.exchange(params["means"], views.positions, degree, degrees_to_use)
    (one render_views call with `views_per_rank` views per step and rank)."""

    def __init__(self, params: Dict[str, torch.Tensor], views_per_rank: int,
                 group: Optional[dist.ProcessGroup] = None, reconstruct=None):
        self.group = group
        self.on = dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1
        self.world = dist.get_world_size(group) if self.on else 1
        self.V = int(views_per_rank)
        self.bucket = GradientBucket({k: v for k, v in params.items() if k != "sh_coeffs"}, group)
        sh = params["sh_coeffs"]
        n, dev = sh.shape[0], sh.device
        self.n = n
        self.sh_grad = torch.zeros(tuple(sh.shape), dtype=torch.float32, device=dev)
        self.rgb_all = torch.zeros((self.world * self.V, n, 3), dtype=torch.float32, device=dev)
        self.pos_all = torch.zeros((self.world * self.V, 3), dtype=torch.float32, device=dev)
        # with a single rank the "gathered" table is the send buffer itself
        self.rgb_send = torch.zeros((self.V, n, 3), dtype=torch.float32, device=dev) if self.on else self.rgb_all
        self._reconstruct = reconstruct

    def rebuild(self, params: Dict[str, torch.Tensor]) -> "FactoredExchange":
        """Adopt a refined parameter set (the Gaussian count changed): fresh bucket and factor tables."""
        self.__init__(params, self.V, self.group, self._reconstruct)
        return self

    def holder(self) -> dict:
        go = getattr(self, "_go", None)
        if go is None or go["v_rgb_views"].data_ptr() != self.rgb_send.data_ptr():
            go = dict(self.bucket.unpack())      # views of the flat buffer: built once, they do not change
            go["v_rgb_views"] = self.rgb_send
            self._go = go
        return {"grad_out": go, "defer_sh_grad": True}

    def exchange(self, means: torch.Tensor, positions: torch.Tensor, degree: int, degrees_to_use: int,
                 holder: Optional[dict] = None, all_positions: Optional[torch.Tensor] = None) -> Dict[str, torch.Tensor]:
        """positions: centres [V,3] of this rank's views.  all_positions: centres [world*V,3] of every
        rank's views in rank order, when the caller knows them (a shared sampler does) -- saves the small
        all-gather."""
        if holder is not None and holder.get("v_rgb_views") is not None and \
                holder["v_rgb_views"].data_ptr() != self.rgb_send.data_ptr():
            self.rgb_send.copy_(holder["v_rgb_views"].view_as(self.rgb_send))
        positions = positions.detach().to(torch.float32).contiguous()
        if positions.shape != (self.V, 3):
            raise ValueError(f"expected the centres of {self.V} views, got {tuple(positions.shape)}")
        reduce_work = None
        if all_positions is not None:
            if tuple(all_positions.shape) != tuple(self.pos_all.shape):
                raise ValueError(f"all_positions must be {tuple(self.pos_all.shape)}")
            self.pos_all.copy_(all_positions)
        if self.on:
            w_pos = None
            if all_positions is None:
                w_pos = dist.all_gather_into_tensor(self.pos_all, positions, group=self.group, async_op=True)
            w_rgb = dist.all_gather_into_tensor(self.rgb_all, self.rgb_send, group=self.group, async_op=True)
            reduce_work = self.bucket.all_reduce(async_op=True)
            if w_pos is not None:
                w_pos.wait()
            w_rgb.wait()
        elif all_positions is None:
            self.pos_all.copy_(positions)
        # rebuilds the SH gradient of ALL views while the all-reduce of the other leaves is in flight
        rec = self._reconstruct
        if rec is None:
            from . import ops
            rec = ops.sh_grad_from_views
        rec(int(degree), int(degrees_to_use), means.detach(), self.pos_all, self.rgb_all, out=self.sh_grad)
        if reduce_work is not None:
            reduce_work.wait()
        grads = dict(self.bucket.unpack())
        grads["sh_coeffs"] = self.sh_grad
        return grads


class NvlsExchange(FactoredExchange):
    """FactoredExchange whose two collectives are ONE kernel of this library over symmetric memory
    (csrc/exchange.cu, gg_nvls_exchange): the bucket and the gathered factor table live in
    torch.distributed._symmetric_memory allocations (peer-mapped, with an NVLS multicast mapping on an NVSwitch
    box); per step: barrier -> [multimem.st all-gather of the rgb slots + multimem.ld_reduce / multimem.st two-shot
    all-reduce of the bucket] -> barrier -> SH rebuild.  No NCCL call on the data path, no staging copy, the
    reduction arithmetic happens in the switch.  Same interface and same results (up to the order of the
    world-term sums) as FactoredExchange; needs an initialised NCCL process group with more than one rank."""

    def __init__(self, params: Dict[str, torch.Tensor], views_per_rank: int,
                 group: Optional[dist.ProcessGroup] = None, reconstruct=None):
        import torch.distributed._symmetric_memory as symm
        if not (dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1):
            raise RuntimeError("NvlsExchange needs an initialised process group with more than one rank")
        self._symm = symm
        self.group = group
        self.on = True
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self.V = int(views_per_rank)
        pg = group if group is not None else dist.group.WORLD
        sh = params["sh_coeffs"]
        n, dev = sh.shape[0], sh.device
        self.n = n

        def symmetric(numel):
            numel = (int(numel) + 3) // 4 * 4
            t = symm.empty(numel, dtype=torch.float32, device=dev)
            t.zero_()
            return t

        self.bucket = GradientBucket({k: v for k, v in params.items() if k != "sh_coeffs"}, group, allocator=symmetric)
        self.slot = (self.V * n * 3 + 3) // 4 * 4            # floats per rank slot, 16-byte multiple
        self.rgb_sym = symmetric(self.world * self.slot)
        self.h_bucket = symm.rendezvous(self.bucket.flat, pg)
        self.h_rgb = symm.rendezvous(self.rgb_sym, pg)
        self.multicast = bool(self.h_bucket.multicast_ptr) and bool(self.h_rgb.multicast_ptr)
        self.sh_grad = torch.zeros(tuple(sh.shape), dtype=torch.float32, device=dev)
        self.pos_all = torch.zeros((self.world * self.V, 3), dtype=torch.float32, device=dev)
        self.rgb_send = self.rgb_sym[self.rank * self.slot:self.rank * self.slot + self.V * n * 3].view(self.V, n, 3)
        # the reconstruction reads the views of all ranks: [world, V, n, 3] when the slots are unpadded
        if self.slot != self.V * n * 3:
            raise ValueError("NvlsExchange: views_per_rank * n * 3 must be a multiple of 4 floats")
        self.rgb_all = self.rgb_sym.view(self.world * self.V, n, 3)
        self._reconstruct = reconstruct
        import ctypes as C
        self._bucket_peers = (C.c_void_p * self.world)(*[int(p) for p in self.h_bucket.buffer_ptrs])
        self._rgb_peers = (C.c_void_p * self.world)(*[int(p) for p in self.h_rgb.buffer_ptrs])

    def rebuild(self, params: Dict[str, torch.Tensor]) -> "NvlsExchange":
        self.__init__(params, self.V, self.group, self._reconstruct)
        return self

    def exchange(self, means: torch.Tensor, positions: torch.Tensor, degree: int, degrees_to_use: int,
                 holder: Optional[dict] = None, all_positions: Optional[torch.Tensor] = None) -> Dict[str, torch.Tensor]:
        from . import _lib, ops
        if holder is not None and holder.get("v_rgb_views") is not None and \
                holder["v_rgb_views"].data_ptr() != self.rgb_send.data_ptr():
            self.rgb_send.copy_(holder["v_rgb_views"].view_as(self.rgb_send))
        if all_positions is not None:
            key = (all_positions.data_ptr(), all_positions._version)
            if getattr(self, "_pos_key", None) != key:     # the same table step after step: copied once
                self.pos_all.copy_(all_positions)
                self._pos_key = key
        else:
            dist.all_gather_into_tensor(self.pos_all, positions.detach().to(torch.float32).contiguous(), group=self.group)
        dev = self.bucket.flat.device
        main = torch.cuda.current_stream(dev)
        if getattr(self, "_side", None) is None:
            self._side = torch.cuda.Stream(device=dev)
        side = self._side
        rec = self._reconstruct or ops.sh_grad_from_views

        def launch(parts):
            _lib.call("gg_nvls_exchange", self.rank, self.world,
                      int(self.h_bucket.multicast_ptr) if self.multicast else None, self._bucket_peers,
                      int(self.bucket.flat.numel()),
                      int(self.h_rgb.multicast_ptr) if self.multicast else None, self._rgb_peers, int(self.slot),
                      parts, _lib.stream_ptr(dev))

        with _lib.device_guard(dev):
            self.h_bucket.barrier(channel=0)      # every rank's backward has written its bucket and its rgb slot
            ready = main.record_event()
            # the reduction of the bucket on a side stream ...
            side.wait_event(ready)
            with torch.cuda.stream(side):
                launch(2)
                self.h_bucket.barrier(channel=1)  # every rank's slice of the bucket has landed everywhere
                reduced = side.record_event()
            # ... while the main stream gathers the factors and rebuilds the SH gradient from them
            launch(1)
            self.h_rgb.barrier(channel=0)         # every rank's rgb slot has landed here
            rec(int(degree), int(degrees_to_use), means.detach(), self.pos_all, self.rgb_all, out=self.sh_grad)
            main.wait_event(reduced)
        grads = dict(self.bucket.unpack())
        grads["sh_coeffs"] = self.sh_grad
        return grads


class SymmetricBucket(GradientBucket):
    """GradientBucket in symmetric memory whose all_reduce() is this library's two-shot NVLS kernel
    (gg_nvls_exchange, reduction only: multimem.ld_reduce of the rank's slice, multimem.st back to every rank)
    between two cross-rank barriers, instead of ncclAllReduce.  For steps whose SH gradient is not factored (many
    views per rank: BASELINE configs[3]).  Needs an initialised process group with more than one rank."""

    def __init__(self, params: Dict[str, torch.Tensor], group: Optional[dist.ProcessGroup] = None):
        import ctypes as C
        import torch.distributed._symmetric_memory as symm
        if not (dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1):
            raise RuntimeError("SymmetricBucket needs an initialised process group with more than one rank")
        dev = next(iter(params.values())).device

        def symmetric(numel):
            t = symm.empty((int(numel) + 3) // 4 * 4, dtype=torch.float32, device=dev)
            t.zero_()
            return t
        super().__init__(params, group, allocator=symmetric)
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        self.handle = symm.rendezvous(self.flat, group if group is not None else dist.group.WORLD)
        self.multicast = bool(self.handle.multicast_ptr)
        self._peers = (C.c_void_p * self.world)(*[int(p) for p in self.handle.buffer_ptrs])

    def all_reduce(self, async_op: bool = False):
        from . import _lib
        dev = self.flat.device
        mc = int(self.handle.multicast_ptr) if self.multicast else None
        with _lib.device_guard(dev):
            self.handle.barrier(channel=0)        # every rank's backward has written its bucket
            _lib.call("gg_nvls_exchange", self.rank, self.world, mc, self._peers, int(self.flat.numel()), mc, self._peers,
                      0, 2, _lib.stream_ptr(dev))
            self.handle.barrier(channel=1)        # every slice has landed everywhere
        return None


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
