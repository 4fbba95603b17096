"""Fused multi-view entry point: one projection, ONE binning and ONE blend pass per batch of
views for everything GaussianGrasper renders per view -- RGB (SH), depth, normal and the D-channel
latent language feature -- instead of the reference's 1 projection + 4 rasterizations with 4
identical binning passes (nerfstudio/models/gaussian_splatting.py:699-784).

Channel layout of the blended image: [rgb(3) | depth(1) | normal(3) | feature(D)], padded to a
multiple of 4 floats per Gaussian row so that rows are gathered with 16-byte cp.async.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence

import torch
from torch.autograd import Function

from . import _lib, ops
from ._torch_impl import quat_to_rotmat
from .sh import SphericalHarmonics


_pinned_ring = {}
_bg_cache = {}


def _background(cp: int, depth_background: float, dev) -> torch.Tensor:
    """[0,0,0, depth_background, 0, ...]: the model's backgrounds (gaussian_splatting.py:745,757,769,783)."""
    key = (cp, depth_background, dev.type, dev.index)
    t = _bg_cache.get(key)
    if t is None:
        host = torch.zeros(cp, dtype=torch.float32)
        host[3] = depth_background
        t = _bg_cache[key] = host.to(dev)
    return t


@dataclass
class ViewBatch:
    """Cameras of one batch of views, already on the device (same image size for all)."""
    viewmats: torch.Tensor   # [V, 12] rows 0..2 of world->camera
    fullmats: torch.Tensor   # [V, 16] projmat @ viewmat
    intrins: torch.Tensor    # [V, 4]  fx, fy, cx, cy
    positions: torch.Tensor  # [V, 3]  camera centres (for SH view directions)
    H: int
    W: int

    @property
    def n_views(self) -> int:
        return self.viewmats.shape[0]

    @staticmethod
    def from_cameras(cams: Sequence, device, out_packed: Optional[torch.Tensor] = None) -> "ViewBatch":
        # one packed row of 35 floats per camera, built once and kept on the camera object
        # The row is cached on the camera together with the identity and version counters of the tensors it
        # was built from: a camera optimiser that updates the pose in place, or assigns a new one, invalidates it.
        rows = []
        for c in cams:
            stamp = (id(c.viewmat), c.viewmat._version, id(c.fullmat), c.fullmat._version, id(c.position),
                     c.position._version, float(c.fx), float(c.fy), float(c.cx), float(c.cy))
            cached = getattr(c, "_gg_row", None)
            if cached is not None and cached[0] == stamp:
                r = cached[1]
            else:
                r = torch.cat([c.viewmat[:3].reshape(-1).float().cpu(), c.fullmat.reshape(-1).float().cpu(),
                               torch.tensor([c.fx, c.fy, c.cx, c.cy], dtype=torch.float32),
                               c.position.reshape(-1).float().cpu()])
                try:
                    c._gg_row = (stamp, r)
                except AttributeError:
                    pass
            rows.append(r)
        V = len(rows)
        H, W = cams[0].H, cams[0].W
        if device.type == "cuda":
            stage, ring, slot = _stage_cameras(rows, device)
            packed = out_packed if out_packed is not None else torch.empty(35 * V, dtype=torch.float32, device=device)
            packed.copy_(stage, non_blocking=True)                  # ONE host->device copy per batch, any V
            ev = torch.cuda.Event()
            ev.record(torch.cuda.current_stream(device))
            ring[2][slot] = ev
        else:
            packed = _field_major(rows, torch.empty(35 * V, dtype=torch.float32))
        return _views_of(packed, V, H, W)


def _field_major(rows, out: torch.Tensor) -> torch.Tensor:
    """[V x 12 viewmat | V x 16 fullmat | V x 4 intrinsics | V x 3 centre]: every field contiguous over the views,
    so one flat buffer serves the kernels' four pointers."""
    V = len(rows)
    if V == 1:
        out.copy_(rows[0])      # a single row is field-major already
        return out
    stacked = torch.stack(rows)                                   # [V, 35]
    o = 0
    for lo, hi in ((0, 12), (12, 28), (28, 32), (32, 35)):
        w = hi - lo
        out[o:o + V * w].view(V, w).copy_(stacked[:, lo:hi])
        o += V * w
    return out


def _views_of(packed: torch.Tensor, V: int, H: int, W: int) -> "ViewBatch":
    vb = ViewBatch(packed[:12 * V].view(V, 12), packed[12 * V:28 * V].view(V, 16), packed[28 * V:32 * V].view(V, 4),
                   packed[32 * V:35 * V].view(V, 3), H, W)
    vb._packed = packed
    return vb


def _stage_cameras(rows, device):
    """Field-major rows in a pinned staging buffer from a small ring per (device, batch size): pinning memory per
    call costs more than the render's whole prepare stage.  Every slot carries the event recorded behind the
    asynchronous copy that last read it; a slot is rewritten only after that event has completed (normally long
    ago -- the wait is free unless a caller stages many batches without rendering)."""
    V = len(rows)
    dev_index = device.index if device.index is not None else torch.cuda.current_device()
    key = (dev_index, V)
    ring = _pinned_ring.get(key)
    if ring is None:
        ring = _pinned_ring[key] = [0, [torch.empty(35 * V).pin_memory() for _ in range(8)], [None] * 8]
    ring[0] = (ring[0] + 1) % len(ring[1])
    slot = ring[0]
    pending = ring[2][slot]
    if pending is not None:
        pending.synchronize()
    return _field_major(rows, ring[1][slot]), ring, slot


def _update_views(self: ViewBatch, cams: Sequence) -> ViewBatch:
    """Refresh the cameras of this batch IN PLACE (same device tensors: what a captured CUDA graph reads) from a
    new list of the same length and image size: one pinned staging row set, asynchronous copies."""
    if len(cams) != self.n_views or cams[0].H != self.H or cams[0].W != self.W:
        raise ValueError("update_ needs the same number of views and the same image size")
    packed = getattr(self, "_packed", None)
    if packed is None:
        raise ValueError("update_ needs a batch made by ViewBatch.from_cameras")
    ViewBatch.from_cameras(cams, packed.device, out_packed=packed)     # one copy into the same storage
    return self


ViewBatch.update_ = _update_views


class _ProjectViews(Function):
    @staticmethod
    def forward(ctx, means, scales, quats, viewmats, fullmats, intrins, H, W, clip_thresh):
        dev = _lib.require_cuda(means, scales, quats, viewmats, fullmats, intrins)
        means, scales, quats = ops.f32c(means), ops.f32c(scales), ops.f32c(quats)
        n, V = means.shape[0], viewmats.shape[0]
        tb = ops.tile_bounds_for(H, W)
        cov3d = torch.empty((V, n, 6), dtype=torch.float32, device=dev)
        xys = torch.empty((V, n, 2), dtype=torch.float32, device=dev)
        depths = torch.empty((V, n), dtype=torch.float32, device=dev)
        radii = torch.empty((V, n), dtype=torch.int32, device=dev)
        conics = torch.empty((V, n, 3), dtype=torch.float32, device=dev)
        nth = torch.empty((V, n), dtype=torch.int32, device=dev)
        with _lib.device_guard(dev):
            _lib.call("gg_project_fwd_views", 
                n, V, ops.ptr(means), ops.ptr(scales), 1.0, ops.ptr(quats), ops.ptr(viewmats), ops.ptr(fullmats),
                ops.ptr(intrins), 0.0, 0.0, 0.0, 0.0, int(H), int(W), tb[0], tb[1], float(clip_thresh),
                ops.ptr(cov3d), ops.ptr(xys), ops.ptr(depths), ops.ptr(radii), ops.ptr(conics), ops.ptr(nth),
                ops.stream_ptr(dev))
        ctx.size = (int(H), int(W))
        ctx.save_for_backward(means, scales, quats, viewmats, fullmats, intrins, radii, conics)
        ctx.mark_non_differentiable(radii, nth)
        return xys, depths, radii, conics, nth

    @staticmethod
    def backward(ctx, v_xys, v_depths, v_radii, v_conics, v_nth):
        means, scales, quats, viewmats, fullmats, intrins, radii, conics = ctx.saved_tensors
        dev = means.device
        n, V = means.shape[0], viewmats.shape[0]
        H, W = ctx.size
        v_xys = ops.f32c(v_xys) if v_xys is not None else torch.zeros((V, n, 2), device=dev)
        v_conics = ops.f32c(v_conics) if v_conics is not None else torch.zeros((V, n, 3), device=dev)
        v_depths = ops.f32c(v_depths) if v_depths is not None else None
        v_means = torch.empty((n, 3), dtype=torch.float32, device=dev)
        v_scales = torch.empty((n, 3), dtype=torch.float32, device=dev)
        v_quats = torch.empty((n, 4), dtype=torch.float32, device=dev)
        with _lib.device_guard(dev):
            _lib.call("gg_project_bwd_views", 
                n, V, ops.ptr(means), ops.ptr(scales), 1.0, ops.ptr(quats), ops.ptr(viewmats), ops.ptr(fullmats),
                ops.ptr(intrins), 0.0, 0.0, 0.0, 0.0, H, W, ops.ptr(radii), ops.ptr(conics), ops.ptr(v_xys),
                ops.ptr(v_depths), ops.ptr(v_conics), 0, ops.ptr(v_means), ops.ptr(v_scales), ops.ptr(v_quats),
                ops.stream_ptr(dev))
        return v_means, v_scales, v_quats, None, None, None, None, None, None


class _RenderViews(Function):
    """prepare (activations + projection + SH + packing) -> bin -> blend, and the exact backward,
    as three + two kernel groups with no torch glue in between."""

    @staticmethod
    def forward(ctx, means, log_scales, quats, opacity_logit, sh_coeffs, features, views, degrees_to_use,
                depth_background, clip_thresh, stats, holder):
        dev = _lib.require_cuda(means, log_scales, quats, opacity_logit, sh_coeffs, features)
        means, log_scales, quats = ops.f32c(means), ops.f32c(log_scales), ops.f32c(quats)
        opacity_logit, sh_coeffs, features = ops.f32c(opacity_logit), ops.f32c(sh_coeffs), ops.f32c(features)
        n, D = means.shape[0], features.shape[1]
        V, H, W = views.n_views, views.H, views.W
        degree = ops.sh_degree_from_bases(sh_coeffs.shape[-2])
        C = 7 + D
        CP = (C + 3) // 4 * 4
        tb = ops.tile_bounds_for(H, W)
        geo = torch.empty((V * n, 8), dtype=torch.float32, device=dev)
        chan = torch.empty((V * n, CP), dtype=torch.float32, device=dev)
        depths = torch.empty((V * n,), dtype=torch.float32, device=dev)
        radii = torch.empty((V * n,), dtype=torch.int32, device=dev)
        nth = torch.empty((V * n,), dtype=torch.int32, device=dev)
        dbg = holder.get("debug_activations") if holder is not None else None
        s_out = torch.empty((n, 3), dtype=torch.float32, device=dev) if dbg else None
        q_out = torch.empty((n, 4), dtype=torch.float32, device=dev) if dbg else None
        def prepare(phase):
            with _lib.device_guard(dev):
                _lib.call("gg_prepare_views", n, V, D, CP, degree, int(degrees_to_use), ops.ptr(means),
                          ops.ptr(log_scales), ops.ptr(quats), ops.ptr(opacity_logit), ops.ptr(sh_coeffs),
                          ops.ptr(features), ops.ptr(views.viewmats), ops.ptr(views.fullmats), ops.ptr(views.intrins),
                          ops.ptr(views.positions), H, W, tb[0], tb[1], float(clip_thresh), ops.ptr(geo),
                          ops.ptr(chan), ops.ptr(depths), ops.ptr(radii), ops.ptr(nth), ops.ptr(s_out), ops.ptr(q_out),
                          phase, ops.stream_ptr(dev))

        # one view: geometry and channel rows in one launch.  Several views: geometry per (view, Gaussian),
        # then the channel rows of all views from one read of the SH / feature rows (prepare_chan_kernel).
        # The binning never waits for the host (ops.bin_views_tiles).
        if V == 1:
            prepare(0)
        else:
            prepare(1)
        binning = ops.bin_views(n, V, geo, depths, radii, nth, tb, xy_from_geo=True,
                                sync_free=not (holder or {}).get("exact_binning", False))
        if V > 1:
            prepare(2)
        bg = _background(CP, float(depth_background), dev)
        out, final_T, final_idx, hit_words = ops.blend_fwd(binning, geo, chan, bg, H, W, colors_per_view=True,
                                                           pair_counter=stats,
                                                           record_hits=any(ctx.needs_input_grad))
        ctx.hit_words = hit_words
        ctx.binning, ctx.views, ctx.dims = binning, views, (n, V, D, CP, degree, int(degrees_to_use), H, W)
        ctx.holder = holder
        ctx.save_for_backward(means, log_scales, quats, opacity_logit, features, geo, chan, radii, bg, final_T,
                              final_idx)
        if holder is not None:
            holder.update(geo=geo, radii=radii.view(V, n), num_tiles_hit=nth.view(V, n), depths=depths.view(V, n),
                          binning=binning, scales=s_out, quats=q_out)
        ctx.mark_non_differentiable(final_T)
        return out, final_T

    @staticmethod
    def backward(ctx, v_out, _v_T):
        means, log_scales, quats, opacity_logit, features, geo, chan, radii, bg, final_T, final_idx = ctx.saved_tensors
        n, V, D, CP, degree, deg_use, H, W = ctx.dims
        views, dev = ctx.views, means.device
        v_geo, v_chan = ops.blend_bwd(ctx.binning, geo, chan, bg, final_T, final_idx, v_out, H, W,
                                      colors_per_view=True, hit_words=ctx.hit_words)
        nb = (degree + 1) ** 2
        go = (ctx.holder or {}).get("grad_out")  # optional caller-owned buffers (views of a flat bucket)

        direct = set()   # leaves whose gradient goes straight into a caller-owned buffer

        def buf(name, shape):
            t = go.get(name) if go else None
            if t is None:
                return torch.empty(shape, dtype=torch.float32, device=dev)
            direct.add(name)
            # a caller-owned gradient buffer that does not fit (e.g. a bucket built before a densification
            # changed N) must not be bypassed silently: the optimizer would then consume stale contents
            if (t.numel() != int(torch.Size(shape).numel()) or not t.is_contiguous() or t.dtype != torch.float32
                    or t.device != dev):
                raise _lib.GGError(
                    f"render_views backward: grad_out[{name!r}] is {tuple(t.shape)} {t.dtype} on {t.device}"
                    f"{'' if t.is_contiguous() else ' (non-contiguous)'}, expected a contiguous fp32 buffer of "
                    f"{tuple(shape)} on {dev}; rebuild the GradientBucket / FactoredExchange after the Gaussian "
                    "count changes")
            return t.view(shape)

        v_means, v_ls, v_q = buf("means", (n, 3)), buf("log_scales", (n, 3)), buf("quats", (n, 4))
        v_op, v_f = buf("opacity_logit", (n,)), buf("features", (n, D))
        # view-sharded training: leave the SH gradient as its per-view factor (distributed.FactoredExchange)
        defer = bool((ctx.holder or {}).get("defer_sh_grad"))
        v_sh = None if defer else buf("sh_coeffs", (n, nb, 3))
        v_rgb = buf("v_rgb_views", (V, n, 3)) if defer else None
        with _lib.device_guard(dev):
            _lib.call("gg_prepare_views_bwd", n, V, D, CP, degree, deg_use, ops.ptr(means), ops.ptr(log_scales),
                      ops.ptr(quats), ops.ptr(opacity_logit), ops.ptr(features), ops.ptr(views.viewmats),
                      ops.ptr(views.fullmats), ops.ptr(views.intrins), ops.ptr(views.positions), H, W, ops.ptr(geo),
                      ops.ptr(chan), ops.ptr(radii), ops.ptr(v_geo), ops.ptr(v_chan), ops.ptr(v_means), ops.ptr(v_ls),
                      ops.ptr(v_q), ops.ptr(v_op), ops.ptr(v_sh), ops.ptr(v_f), ops.ptr(v_rgb), ops.stream_ptr(dev))
        if ctx.holder is not None:
            ctx.holder["v_geo"] = v_geo  # [V*n, 8]: columns 0..1 are d loss / d xys (densification statistic)
            if ctx.holder.get("debug_activations"):
                ctx.holder["v_chan"] = v_chan  # [V*n, CP]: gradient of the channel table (parity tests)
            if defer:
                ctx.holder["v_rgb_views"] = v_rgb
        # A gradient written into a caller-owned buffer is NOT also handed to autograd: AccumulateGrad would either
        # clone it (a wasted pass over the buffer) or adopt the buffer as `.grad` and add the next step's gradient
        # into it in place -- on top of what this kernel has just written there.  Those leaves keep `.grad = None`.
        ret = dict(means=v_means, log_scales=v_ls, quats=v_q, opacity_logit=v_op.reshape(opacity_logit.shape),
                   sh_coeffs=v_sh, features=v_f)
        return tuple(None if k in direct else ret[k] for k in
                     ("means", "log_scales", "quats", "opacity_logit", "sh_coeffs", "features")) + (None,) * 6


def smallest_axis_normals(quats: torch.Tensor, log_scales: torch.Tensor) -> torch.Tensor:
    """gaussian_splatting.py:605-619: column of R(quat) along the smallest scale."""
    R = quat_to_rotmat(quats)
    idx = log_scales.min(dim=-1)[1][..., None, None].expand(-1, 3, -1)
    return R.gather(2, idx).squeeze(dim=2)


def render_views(means, log_scales, quats, opacity_logit, sh_coeffs, features, views: ViewBatch,
                 degrees_to_use: int = 4, depth_background: float = 10.0, clip_thresh: float = 0.01,
                 stats: Optional[torch.Tensor] = None, holder: Optional[dict] = None) -> Dict[str, torch.Tensor]:
    """Render rgb [V,H,W,3], depth [V,H,W,1], normal [V,H,W,3], feature [V,H,W,D] for V views.

    Inputs are the model's raw parameters (gaussian_splatting.py:270-292): log-scales, un-normalised
    wxyz quaternions, opacity logits, SH coefficients [N,(deg+1)^2,3], features [N,D] (D <= 64).
    Backgrounds follow the model: 0 for rgb/normal/feature, `depth_background` (10) for depth.
    `holder` (a dict, optional) receives the per-view projection by-products -- packed geo records
    (pixel centres in columns 0..1), radii, num_tiles_hit, depths, the binning -- and, after
    backward, `v_geo` whose first two columns are d loss / d xys, the statistic the model's
    densification reads from `xys.grad` (gaussian_splatting.py:725,377).  If it carries
    `grad_out` (name -> preallocated fp32 tensor, e.g. `GradientBucket.unpack()`), the backward
    writes the leaf gradients of a single-launch step straight into those buffers (overwriting them) and the
    leaves named there get NO `.grad` from autograd: the buffers are the gradients.  With
    `defer_sh_grad` set, the SH-coefficient gradient is not formed: `sh_coeffs.grad` stays None and
    `holder["v_rgb_views"]` [V,N,3] receives its per-view factor (see distributed.FactoredExchange).
    """
    out, final_T = _RenderViews.apply(means, log_scales, quats, opacity_logit, sh_coeffs, features, views,
                                      int(degrees_to_use), float(depth_background), float(clip_thresh), stats, holder)
    D = features.shape[1]
    return dict(rgb=out[..., 0:3], depth=out[..., 3:4], normal=out[..., 4:7], feature=out[..., 7:7 + D],
                image=out, alpha=1.0 - final_T)
