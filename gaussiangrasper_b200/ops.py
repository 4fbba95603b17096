"""Thin tensor-level wrappers over the C ABI (no autograd here).

Every function takes CUDA tensors, allocates its outputs with torch on the same device and
enqueues the kernels on torch's current stream.  Mirrors, one to one, the gsplat.cuda bindings
that gsplat 0.1.0's Python layer calls (see include/gg_b200.h for the mapping).
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch

from . import _lib
from ._lib import check, f32c, ptr, require_cuda, stream_ptr

TILE = 16
_INFO_SLOTS = 64


def tile_bounds_for(img_height: int, img_width: int) -> Tuple[int, int, int]:
    return ((img_width + TILE - 1) // TILE, (img_height + TILE - 1) // TILE, 1)


# ------------------------------------------------------------------------------------------
# per-device workspace for the binning stage
# ------------------------------------------------------------------------------------------
class _Workspace:
    def __init__(self, device):
        self.device = device
        self.scan = None
        self.sort = None
        self.keys = None
        self.ids = None
        self.keys_sorted = None
        self.order_ws = None
        self.begin = None
        self.finish = None
        self.total = torch.zeros(1, dtype=torch.int32, device=device)
        self.total_host = torch.zeros(1, dtype=torch.int32).pin_memory()
        self.total_event = torch.cuda.Event()
        # tile-first binning: scratch, capacity per problem signature, ring of pinned info records
        self.tiles_scratch = None
        self.capacity = {}
        self.longest = {}        # longest tile list seen per signature (hint for the sort classes to launch)
        self.info_host = torch.zeros((_INFO_SLOTS, 4), dtype=torch.int32).pin_memory()
        self.info_next = 0
        self.info_pending = []   # BinInfo records whose read-back has not been looked at yet
        self.overflow_note = None

    def scan_ws(self, n):
        need = int(_lib.load().gg_cumsum_workspace_bytes(n))
        if self.scan is None or self.scan.numel() < need:
            self.scan = torch.empty(int(need * 1.5) + 256, dtype=torch.uint8, device=self.device)
        return self.scan

    def sort_ws(self, m):
        need = int(_lib.load().gg_sort_workspace_bytes(m))
        if self.sort is None or self.sort.numel() < need:
            # zero-filled once: the library owns the look-back status words afterwards
            self.sort = torch.zeros(int(need * 1.5) + 256, dtype=torch.uint8, device=self.device)
        return self.sort

    def begin_scratch(self, total):
        need = int(_lib.load().gg_bin_begin_scratch_bytes(total))
        if self.begin is None or self.begin.numel() < need:
            self.begin = torch.zeros(int(need * 1.25) + 4096, dtype=torch.uint8, device=self.device)
        return self.begin

    def finish_scratch(self, m):
        need = int(_lib.load().gg_bin_finish_scratch_bytes(m))
        if self.finish is None or self.finish.numel() < need:
            self.finish = torch.zeros(int(need * 1.5) + 4096, dtype=torch.uint8, device=self.device)
        return self.finish

    def note(self, info):
        """Book-keeping of one finished read-back: remember the largest count per signature, raise the
        capacity past an overflow."""
        m, overflow = info.values[0], info.values[1]
        self.longest[info.key] = max(self.longest.get(info.key, 0), info.values[2])
        cap = self.capacity.get(info.key)
        if cap is not None and (overflow or m > 0.85 * cap):
            self.capacity[info.key] = max(cap, _capacity_for(m))
        if overflow and info.capacity > 0:
            self.overflow_note = (m, info.capacity)

    def poll(self):
        """Look at the read-backs that have arrived (never blocks); report an overflow of an earlier call."""
        keep = []
        for info in self.info_pending:
            if info.values is None and info.ready():
                info.wait()
            if info.values is None:
                keep.append(info)
        self.info_pending = keep[-(_INFO_SLOTS - 8):]
        if self.overflow_note is not None:
            m, cap = self.overflow_note
            self.overflow_note = None
            raise _lib.GGError(
                f"tile binning overflow in an earlier call: {m} intersections > capacity {cap}; that call drew the "
                "background only. The capacity has been raised: repeat the step")

    def tiles_scratch_for(self, n_views, tiles_per_view, capacity):
        need = int(_lib.load().gg_bin_tiles_scratch_bytes(n_views, tiles_per_view, capacity))
        if self.tiles_scratch is None or self.tiles_scratch.numel() < need:
            self.tiles_scratch = torch.empty(int(need * 1.25) + 4096, dtype=torch.uint8, device=self.device)
        return self.tiles_scratch

    def key_buffers(self, m):
        if self.keys is None or self.keys.numel() < m:
            cap = int(m * 1.5) + 1024
            self.keys = torch.empty(cap, dtype=torch.int64, device=self.device)
            self.ids = torch.empty(cap, dtype=torch.int32, device=self.device)
            self.keys_sorted = torch.empty(cap, dtype=torch.int64, device=self.device)
        return self.keys, self.ids, self.keys_sorted


_workspaces = {}


def workspace(device) -> _Workspace:
    key = (device.type, device.index if device.index is not None else torch.cuda.current_device())
    ws = _workspaces.get(key)
    if ws is None:
        ws = _workspaces[key] = _Workspace(device)
    return ws


# ------------------------------------------------------------------------------------------
# projection / SH
# ------------------------------------------------------------------------------------------
def project_fwd(means3d, scales, glob_scale, quats, viewmat, fullmat, fx, fy, cx, cy, img_height, img_width,
                tile_bounds, clip_thresh=0.01):
    dev = require_cuda(means3d, scales, quats, viewmat, fullmat)
    means3d, scales, quats = f32c(means3d), f32c(scales), f32c(quats)
    vm = f32c(viewmat).reshape(-1)
    fm = f32c(fullmat).reshape(-1)
    if vm.numel() < 12 or fm.numel() != 16:
        raise ValueError("viewmat must have at least 3x4 and fullmat 4x4 entries")
    n = means3d.shape[0]
    cov3d = torch.empty((n, 6), dtype=torch.float32, device=dev)
    xys = torch.empty((n, 2), dtype=torch.float32, device=dev)
    depths = torch.empty((n,), dtype=torch.float32, device=dev)
    radii = torch.empty((n,), dtype=torch.int32, device=dev)
    conics = torch.empty((n, 3), dtype=torch.float32, device=dev)
    nth = torch.empty((n,), dtype=torch.int32, device=dev)
    with _lib.device_guard(dev):
        _lib.call("gg_project_fwd", 
            n, ptr(means3d), ptr(scales), float(glob_scale), ptr(quats), ptr(vm), ptr(fm), float(fx), float(fy),
            float(cx), float(cy), int(img_height), int(img_width), int(tile_bounds[0]), int(tile_bounds[1]),
            float(clip_thresh), ptr(cov3d), ptr(xys), ptr(depths), ptr(radii), ptr(conics), ptr(nth),
            stream_ptr(dev))
    return xys, depths, radii, conics, nth, cov3d


def project_bwd(means3d, scales, glob_scale, quats, viewmat, fullmat, fx, fy, cx, cy, img_height, img_width, radii,
                conics, v_xys, v_depths, v_conics):
    dev = require_cuda(means3d, scales, quats, viewmat, fullmat, radii, conics, v_xys, v_conics)
    means3d, scales, quats = f32c(means3d), f32c(scales), f32c(quats)
    vm, fm = f32c(viewmat).reshape(-1), f32c(fullmat).reshape(-1)
    conics, v_xys, v_conics = f32c(conics), f32c(v_xys), f32c(v_conics)
    v_depths = f32c(v_depths) if v_depths is not None else None
    n = means3d.shape[0]
    v_means = torch.empty((n, 3), dtype=torch.float32, device=dev)
    v_scales = torch.empty((n, 3), dtype=torch.float32, device=dev)
    v_quats = torch.empty((n, 4), dtype=torch.float32, device=dev)
    with _lib.device_guard(dev):
        _lib.call("gg_project_bwd", 
            n, ptr(means3d), ptr(scales), float(glob_scale), ptr(quats), ptr(vm), ptr(fm), float(fx), float(fy),
            float(cx), float(cy), int(img_height), int(img_width), ptr(radii), ptr(conics), ptr(v_xys),
            ptr(v_depths), ptr(v_conics), ptr(v_means), ptr(v_scales), ptr(v_quats), stream_ptr(dev))
    return v_means, v_scales, v_quats


def sh_degree_from_bases(num_bases: int) -> int:
    table = {1: 0, 4: 1, 9: 2, 16: 3, 25: 4}
    if num_bases not in table:
        raise ValueError(f"coeffs has {num_bases} SH bases; expected 1, 4, 9, 16 or 25")
    return table[num_bases]


def sh_grad_from_views(degree: int, degrees_to_use: int, means, positions, v_rgb_views, out=None):
    """v_sh [N,(degree+1)^2,3] = sum over views of Y(dir) (x) v_rgb: positions [Vt,3], v_rgb_views [Vt,N,3]."""
    dev = require_cuda(means, positions, v_rgb_views)
    means, positions, v_rgb_views = f32c(means), f32c(positions), f32c(v_rgb_views)
    n, vt = means.shape[0], positions.shape[0]
    if v_rgb_views.numel() != vt * n * 3:
        raise ValueError("v_rgb_views must hold [views, N, 3] floats for the views in `positions`")
    nb = (degree + 1) ** 2
    if out is None:
        out = torch.empty((n, nb, 3), dtype=torch.float32, device=dev)
    with _lib.device_guard(dev):
        _lib.call("gg_sh_grad_from_views", int(n), int(vt), int(degree), int(degrees_to_use), ptr(means), ptr(positions),
                  ptr(v_rgb_views), ptr(out), stream_ptr(dev))
    return out


def sh_fwd(degrees_to_use, viewdirs, coeffs):
    dev = require_cuda(viewdirs, coeffs)
    viewdirs, coeffs = f32c(viewdirs), f32c(coeffs)
    n = coeffs.shape[0]
    degree = sh_degree_from_bases(coeffs.shape[-2])
    colors = torch.empty((n, 3), dtype=torch.float32, device=dev)
    with _lib.device_guard(dev):
        _lib.call("gg_sh_fwd", n, degree, int(degrees_to_use), ptr(viewdirs), ptr(coeffs), ptr(colors),
                                    stream_ptr(dev))
    return colors


def sh_bwd(degree, degrees_to_use, viewdirs, v_colors):
    dev = require_cuda(viewdirs, v_colors)
    viewdirs, v_colors = f32c(viewdirs), f32c(v_colors)
    n = viewdirs.shape[0]
    nb = (degree + 1) ** 2
    v_coeffs = torch.empty((n, nb, 3), dtype=torch.float32, device=dev)
    with _lib.device_guard(dev):
        _lib.call("gg_sh_bwd", n, int(degree), int(degrees_to_use), ptr(viewdirs), ptr(v_colors), ptr(v_coeffs),
                                    stream_ptr(dev))
    return v_coeffs


# ------------------------------------------------------------------------------------------
# binning
# ------------------------------------------------------------------------------------------
def cumsum_i32(x: torch.Tensor, total_out: Optional[torch.Tensor] = None) -> torch.Tensor:
    dev = require_cuda(x)
    x = x.contiguous()
    assert x.dtype == torch.int32
    out = torch.empty_like(x)
    ws = workspace(dev).scan_ws(x.numel())
    with _lib.device_guard(dev):
        _lib.call("gg_cumsum", x.numel(), ptr(x), ptr(out), ptr(total_out), ptr(ws), ws.numel(),
                                    stream_ptr(dev))
    return out


def key_bits_for(num_tiles_total: int) -> int:
    return 32 + max(1, (max(num_tiles_total, 1) - 1).bit_length())


def map_to_intersects(n, n_views, xys, depths, radii, cum, tile_bounds, keys, ids, xy_from_geo=False):
    dev = xys.device
    with _lib.device_guard(dev):
        _lib.call("gg_map_to_intersects_geo" if xy_from_geo else "gg_map_to_intersects", int(n), int(n_views), ptr(xys), ptr(depths), ptr(radii), ptr(cum),
                                               int(tile_bounds[0]), int(tile_bounds[1]), ptr(keys), ptr(ids),
                                               stream_ptr(dev))


def sort_pairs(m, key_bits, keys_in, ids_in, keys_out, ids_out):
    dev = keys_in.device
    ws = workspace(dev).sort_ws(m)
    with _lib.device_guard(dev):
        _lib.call("gg_sort_pairs", int(m), int(key_bits), ptr(keys_in), ptr(ids_in), ptr(keys_out), ptr(ids_out),
                                        ptr(ws), ws.numel(), stream_ptr(dev))


def tile_ranges(m, keys_sorted, num_tiles):
    dev = keys_sorted.device
    ranges = torch.empty((num_tiles, 2), dtype=torch.int32, device=dev)
    with _lib.device_guard(dev):
        _lib.call("gg_tile_ranges", int(m), ptr(keys_sorted), int(num_tiles), ptr(ranges), stream_ptr(dev))
    return ranges


def tile_order(ranges: torch.Tensor) -> torch.Tensor:
    dev = ranges.device
    ws = workspace(dev)
    if ws.order_ws is None:
        ws.order_ws = torch.empty(int(_lib.load().gg_tile_order_workspace_bytes()), dtype=torch.uint8, device=dev)
    order = torch.empty((ranges.shape[0],), dtype=torch.int32, device=dev)
    with _lib.device_guard(dev):
        _lib.call("gg_tile_order", int(ranges.shape[0]), ptr(ranges), ptr(order), ptr(ws.order_ws), ws.order_ws.numel(),
                  stream_ptr(dev))
    return order


class BinInfo:
    """{M, overflow, longest tile list} of one gg_bin_tiles call, read back without blocking the stream:
    the record lands in a pinned ring slot; `wait()` blocks the host only when somebody asks."""

    def __init__(self, ws, slot, event, key, capacity):
        self.ws, self.slot, self.event, self.key, self.capacity = ws, slot, event, key, capacity
        self.values = None

    def ready(self) -> bool:
        return self.values is not None or self.event is None or self.event.query()

    def wait(self):
        if self.values is None:
            if self.event is not None:
                self.event.synchronize()
            else:  # recorded inside a stream capture: the caller synchronises after the replay
                torch.cuda.current_stream(self.ws.device).synchronize()
            self.values = tuple(int(x) for x in self.ws.info_host[self.slot])
            self.ws.note(self)
        return self.values


class Binning:
    """Sorted intersection list of one batch of views (what the blend kernels consume).

    ids_buffer holds `capacity` >= M entries; M itself (`num_intersects`) is known on the host only after a
    read-back, which the tile-first path never waits for: `ids_sorted` / `num_intersects` / `check()` block
    on it when they are asked."""

    def __init__(self, n, n_views, ids_buffer, tile_ranges, tile_bounds, tile_order=None, num_intersects=None,
                 info: Optional["BinInfo"] = None):
        self.n, self.n_views = n, n_views
        self.ids_buffer = ids_buffer          # [capacity] int32
        self.tile_ranges = tile_ranges        # [V*T, 2] int32
        self.tile_bounds = tile_bounds
        self.tile_order = tile_order          # [V*T] int32, tiles by descending list length
        self._m = num_intersects
        self.info = info

    @property
    def capacity(self) -> int:
        return int(self.ids_buffer.numel())

    @property
    def num_intersects(self) -> int:
        if self._m is None:
            self.check()
        return self._m

    @property
    def ids_sorted(self) -> torch.Tensor:
        return self.ids_buffer[:self.num_intersects]

    def check(self) -> "Binning":
        """Block until the device's intersection count is on the host; raise if it exceeded the capacity
        (the render that used this binning then drew the background only)."""
        if self._m is None:
            m, overflow, _longest, _ = self.info.wait()
            if overflow:
                self.info.ws.overflow_note = None   # reported here, not again by the next call
                raise _lib.GGError(
                    f"tile binning overflow: {m} intersections > capacity {self.info.capacity}; the images of "
                    "this call are invalid (background only). The capacity has been raised: repeat the call")
            self._m = m
        return self


def _capacity_for(m: int) -> int:
    return int(m * 1.3) + 65536


def bin_views_tiles(n, n_views, xys, depths, radii, tile_bounds, xy_from_geo=False, sync_free=True) -> Binning:
    """Tile-first binning (gg_bin_tiles): count -> scan -> scatter -> per-tile sort, no host read.

    The buffers are sized by a capacity remembered per problem signature (rows, tiles): the first call of a
    signature measures M (one synchronous counting pass), later calls reuse 1.3 x the largest M seen and never
    wait.  An overflow (M jumped past the capacity between two calls) leaves every tile range empty, is
    reported by `Binning.check()` / the next call, and raises the capacity.  sync_free=False waits for the
    count after every call and repeats the binning itself on overflow (always exact, one host wait)."""
    dev = require_cuda(xys, depths, radii)
    ws = workspace(dev)
    xys, depths = f32c(xys), f32c(depths)
    radii = radii.contiguous().reshape(-1)
    tiles_x, tiles_y = int(tile_bounds[0]), int(tile_bounds[1])
    num_tiles = tiles_x * tiles_y * n_views
    key = (int(n) * int(n_views), num_tiles)
    capturing = torch.cuda.is_current_stream_capturing()
    if not capturing:
        ws.poll()
    if _FORCE_SYNC:
        sync_free = False

    def run(capacity):
        scratch = ws.tiles_scratch_for(int(n_views), tiles_x * tiles_y, capacity)
        ids = torch.empty((max(capacity, 1),), dtype=torch.int32, device=dev)
        ranges = torch.empty((num_tiles, 2), dtype=torch.int32, device=dev)
        order_t = torch.empty((num_tiles,), dtype=torch.int32, device=dev)
        slot = ws.info_next
        ws.info_next = (slot + 1) % _INFO_SLOTS
        with _lib.device_guard(dev):
            _lib.call("gg_bin_tiles", int(n), int(n_views), ptr(xys), 8 if xy_from_geo else 2, ptr(depths), ptr(radii),
                      tiles_x, tiles_y, int(capacity), ptr(scratch), scratch.numel(), ptr(ids), ptr(ranges), ptr(order_t),
                      None, ws.info_host[slot].data_ptr(), int(ws.longest.get(key, 0) * 1.25), stream_ptr(dev))
        ev = None
        if not capturing:
            ev = torch.cuda.Event()
            ev.record(torch.cuda.current_stream(dev))
        info = BinInfo(ws, slot, ev, key, capacity)
        if not capturing:
            ws.info_pending.append(info)
        return Binning(n, n_views, ids, ranges, tuple(tile_bounds), order_t, None, info)

    cap = ws.capacity.get(key)
    if cap is None:
        if capturing:
            raise _lib.GGError("bin_views_tiles: the first call of a problem size measures the intersection count "
                               "on the host; run one step before capturing a CUDA graph")
        m, _, _, _ = run(0).info.wait()       # counting pass only
        cap = ws.capacity[key] = _capacity_for(m)
    b = run(cap)
    if not sync_free and not capturing:
        m, overflow, _, _ = b.info.wait()
        if overflow:
            ws.overflow_note = None
            b = run(ws.capacity[key])         # note() has raised it past m
            b.check()
        else:
            b._m = m
    return b


import os as _os
_FORCE_SYNC = _os.environ.get("GG_BIN_SYNC", "0") not in ("", "0")


def bin_views(n, n_views, xys, depths, radii, num_tiles_hit, tile_bounds, xy_from_geo=False,
              while_waiting=None, mode="tiles", sync_free=True) -> Binning:
    """Sorted (tile, depth, id) intersection list of a batch of views.

    mode "tiles" (default): tile-first, no host read -- see bin_views_tiles.
    mode "depth": sort the V*n Gaussians by (view, depth), emit their tile entries in that order, stable-sort
    the M entries by tile id -- 4-5 small passes + 2-3 M-sized passes, one host read of M.
    mode "reference": the reference's formulation: emit 64-bit tile|depth keys in id order and sort all M of
    them (6-7 M-sized passes).  All three give bit-identical ids_sorted / tile_ranges.
    xy_from_geo: `xys` is the packed [V*n, 8] geo table (pixel centres in columns 0..1).
    while_waiting: callable run after the M read-back is enqueued and before the host waits for it."""
    if mode not in ("tiles", "depth", "reference"):
        raise ValueError(f"bin_views: unknown mode {mode!r}")
    if mode == "tiles":
        b = bin_views_tiles(n, n_views, xys, depths, radii, tile_bounds, xy_from_geo, sync_free=sync_free)
        if while_waiting is not None:
            while_waiting()
        return b
    depth_first = mode == "depth"
    dev = require_cuda(xys, depths, radii, num_tiles_hit)
    ws = workspace(dev)
    lib_call = _lib.call
    xys, depths = f32c(xys), f32c(depths)
    radii, num_tiles_hit = radii.contiguous().reshape(-1), num_tiles_hit.contiguous().reshape(-1)
    total = n * n_views
    tiles_x, tiles_y = int(tile_bounds[0]), int(tile_bounds[1])
    num_tiles = tiles_x * tiles_y * n_views
    with _lib.device_guard(dev):
        st = stream_ptr(dev)
        # the one device->host read of these two paths (the reference has five per view): M sizes the sort.
        # Work passed as `while_waiting` is enqueued behind the copy so the GPU stays busy meanwhile.
        if depth_first:
            begin = ws.begin_scratch(total)
            lib_call("gg_bin_begin", int(n), int(n_views), ptr(depths), ptr(num_tiles_hit), ptr(begin), begin.numel(),
                     ws.total_host.data_ptr(), st)
            if while_waiting is not None:
                while_waiting()
            lib_call("gg_bin_wait")
        else:
            cum = cumsum_i32(num_tiles_hit, ws.total)
            ws.total_host.copy_(ws.total, non_blocking=True)
            ws.total_event.record(torch.cuda.current_stream(dev))
            if while_waiting is not None:
                while_waiting()
            ws.total_event.synchronize()
        m = int(ws.total_host[0])
        if m < 0:
            raise _lib.GGError("number of tile intersections overflows int32")
        if m == 0:
            ranges = torch.zeros((num_tiles, 2), dtype=torch.int32, device=dev)
            return Binning(n, n_views, torch.empty((0,), dtype=torch.int32, device=dev), ranges, tuple(tile_bounds), None, 0)
        ids_sorted = torch.empty((m,), dtype=torch.int32, device=dev)
        if depth_first:
            ranges = torch.empty((num_tiles, 2), dtype=torch.int32, device=dev)
            order_t = torch.empty((num_tiles,), dtype=torch.int32, device=dev)
            fin = ws.finish_scratch(m)
            lib_call("gg_bin_finish", int(n), int(n_views), m, ptr(xys), 8 if xy_from_geo else 2, ptr(radii), tiles_x,
                     tiles_y, ptr(begin), ptr(fin), fin.numel(), ptr(ids_sorted), ptr(ranges), ptr(order_t), st)
        else:
            keys, ids, keys_sorted = ws.key_buffers(m)
            map_to_intersects(n, n_views, xys, depths, radii, cum, tile_bounds, keys, ids, xy_from_geo)
            sort_pairs(m, key_bits_for(num_tiles), keys, ids, keys_sorted, ids_sorted)
            ranges = tile_ranges(m, keys_sorted, num_tiles)
            order_t = tile_order(ranges)
    return Binning(n, n_views, ids_sorted, ranges, tuple(tile_bounds), order_t, m)


# ------------------------------------------------------------------------------------------
# blending
# ------------------------------------------------------------------------------------------
def pack_geo(n, n_views, xys, conics, opacity, opac_per_view=False):
    dev = require_cuda(xys, conics, opacity)
    xys, conics, opacity = f32c(xys), f32c(conics), f32c(opacity).reshape(-1)
    geo = torch.empty((n * n_views, 8), dtype=torch.float32, device=dev)
    with _lib.device_guard(dev):
        _lib.call("gg_pack_geo", int(n), int(n_views), ptr(xys), ptr(conics), ptr(opacity),
                                      1 if opac_per_view else 0, ptr(geo), stream_ptr(dev))
    return geo


def _ids_ptr(binning: "Binning"):
    """Pointer of the sorted id list; an empty list (nothing visible) still needs a valid address
    for the C ABI's null checks -- no kernel dereferences it, every tile range is (0,0)."""
    if binning.ids_buffer.numel() > 0:
        return ptr(binning.ids_buffer)
    return ptr(workspace(binning.tile_ranges.device).total)


def max_channels() -> int:
    return int(_lib.load().gg_blend_max_channels())


def _channel_blocks(C: int, step: int):
    """Column blocks of a C-channel blend: one launch for C <= step, else near-equal blocks of a multiple of 4
    channels (16-byte aligned row pieces).  Returns ([(c0, c1), ...], same_batch): same_batch says that every block
    stages the same number of entries per batch (<= 32 channels: 128, more: 64) -- then the contributor masks the
    forward records for the first block are valid for all of them (they depend on the geometry only)."""
    if C <= step:
        return [(0, C)], True
    nblk = (C + step - 1) // step
    per = min(step, ((C + nblk - 1) // nblk + 3) // 4 * 4)
    blocks = [(c0, min(C, c0 + per)) for c0 in range(0, C, per)]
    wide = [(c1 - c0) > 32 for c0, c1 in blocks]
    return blocks, all(wide) or not any(wide)


def blend_fwd(binning: Binning, geo, colors, background, img_height, img_width, colors_per_view=False,
              pair_counter: Optional[torch.Tensor] = None, single_image: bool = False, record_hits: bool = True):
    """colors [rows, C] (C arbitrary; split into launches of <= 64 channels).  Returns
    (out [V,H,W,C], final_T [V,H,W], final_idx [V,H,W], hit_words or None).  hit_words is the forward's
    table of contributing entries per (tile batch, warp); give it to blend_bwd."""
    dev = require_cuda(geo, colors, background)
    colors, background = f32c(colors), f32c(background)
    V, n, C = binning.n_views, binning.n, colors.shape[1]
    # single_image: the drop-in rasterizers return [H,W,C]; it must not be a view of a batched
    # tensor because the model writes into it in place (gaussian_splatting.py:884)
    shape = (img_height, img_width, C) if single_image else (V, img_height, img_width, C)
    out = torch.empty(shape, dtype=torch.float32, device=dev)
    final_T = torch.empty((V, img_height, img_width), dtype=torch.float32, device=dev)
    final_idx = torch.empty((V, img_height, img_width), dtype=torch.int32, device=dev)
    tb = binning.tile_bounds
    blocks, same_batch = _channel_blocks(C, max_channels())
    hit_words = None
    if record_hits and same_batch and binning.capacity > 0:
        nwords = int(_lib.load().gg_blend_hit_words(binning.capacity, binning.tile_ranges.shape[0], blocks[0][1] - blocks[0][0]))
        hit_words = torch.zeros(nwords, dtype=torch.int32, device=dev)
    with _lib.device_guard(dev):
        for c0, c1 in blocks:
            _lib.call("gg_blend_fwd",
                V, n, c1 - c0, C, 1 if colors_per_view else 0, C, int(img_height), int(img_width), tb[0], tb[1],
                _ids_ptr(binning), ptr(binning.tile_ranges), ptr(binning.tile_order), ptr(geo),
                colors.data_ptr() + 4 * c0, background.data_ptr() + 4 * c0, out.data_ptr() + 4 * c0, ptr(final_T),
                ptr(final_idx), ptr(pair_counter) if c0 == 0 else None, ptr(hit_words) if c0 == 0 else None,
                stream_ptr(dev))
    return out, final_T, final_idx, hit_words


def blend_bwd(binning: Binning, geo, colors, background, final_T, final_idx, v_out, img_height, img_width,
              colors_per_view=False, hit_words: Optional[torch.Tensor] = None):
    """Returns (v_geo [V*n, 8], v_colors like colors)."""
    dev = require_cuda(geo, colors, background, v_out)
    colors, background, v_out = f32c(colors), f32c(background), f32c(v_out)
    V, n, C = binning.n_views, binning.n, colors.shape[1]
    # one zero fill for both accumulation targets (the kernels add into them with red.global)
    rows_g, rows_c = V * n * 8, colors.numel()
    acc = torch.zeros(rows_g + rows_c, dtype=torch.float32, device=dev)
    v_geo = acc[:rows_g].view(V * n, 8)
    v_colors = acc[rows_g:].view(colors.shape)
    tb = binning.tile_bounds
    blocks, same_batch = _channel_blocks(C, max_channels())
    if not same_batch:
        hit_words = None
    with _lib.device_guard(dev):
        for c0, c1 in blocks:
            _lib.call("gg_blend_bwd",
                V, n, c1 - c0, C, 1 if colors_per_view else 0, C, int(img_height), int(img_width), tb[0], tb[1],
                _ids_ptr(binning), ptr(binning.tile_ranges), ptr(binning.tile_order), ptr(geo),
                colors.data_ptr() + 4 * c0, background.data_ptr() + 4 * c0, ptr(final_T), ptr(final_idx),
                v_out.data_ptr() + 4 * c0, ptr(hit_words), ptr(v_geo), v_colors.data_ptr() + 4 * c0, stream_ptr(dev))
    return v_geo, v_colors


def blend_pair_stats(binning: Binning, geo, img_height, img_width):
    """(pairs visited, pairs blended) of a binned batch -- the K of SURVEY 8d and its contributing subset."""
    dev = require_cuda(geo)
    stats = torch.zeros(2, dtype=torch.int64, device=dev)
    tb = binning.tile_bounds
    with _lib.device_guard(dev):
        _lib.call("gg_blend_pair_stats", binning.n_views, binning.n, int(img_height), int(img_width), tb[0], tb[1],
                  _ids_ptr(binning), ptr(binning.tile_ranges), ptr(geo), ptr(stats), stream_ptr(dev))
    visited, blended = (int(x) for x in stats.tolist())
    return visited, blended


def unpack_vgeo(n, n_views, v_geo):
    dev = v_geo.device
    v_xys = torch.empty((n_views * n, 2), dtype=torch.float32, device=dev)
    v_conics = torch.empty((n_views * n, 3), dtype=torch.float32, device=dev)
    v_opac = torch.empty((n,), dtype=torch.float32, device=dev)
    with _lib.device_guard(dev):
        _lib.call("gg_unpack_vgeo", int(n), int(n_views), ptr(v_geo), ptr(v_xys), ptr(v_conics), ptr(v_opac), 0,
                                         stream_ptr(dev))
    return v_xys, v_conics, v_opac
