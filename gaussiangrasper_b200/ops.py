"""Thin tensor-level wrappers over the C ABI (no autograd here).

Every function takes CUDA tensors, allocates its outputs with torch on the same device and
enqueues the kernels on torch's current stream.  Mirrors, one to one, the gsplat.cuda bindings
that gsplat 0.1.0's Python layer calls (see include/gg_b200.h for the mapping).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional, Tuple

import torch

from . import _lib
from ._lib import check, f32c, ptr, require_cuda, stream_ptr

TILE = 16


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


@dataclass
class Binning:
    """Sorted intersection list of one batch of views (what the blend kernels consume)."""
    n: int
    n_views: int
    num_intersects: int
    ids_sorted: torch.Tensor    # [M] int32
    tile_ranges: torch.Tensor   # [V*T, 2] int32
    tile_bounds: Tuple[int, int, int]
    tile_order: Optional[torch.Tensor] = None  # [V*T] int32, tiles by descending list length


def bin_views(n, n_views, xys, depths, radii, num_tiles_hit, tile_bounds, xy_from_geo=False,
              while_waiting=None, depth_first=True) -> Binning:
    """Sorted (tile, depth, id) intersection list of a batch of views.

    depth_first (default): sort the V*n Gaussians by (view, depth), emit their tile entries in that
    order, stable-sort the M entries by tile id -- 4-5 small passes + 2-3 M-sized passes.  Otherwise the
    reference's formulation: emit 64-bit tile|depth keys in id order and sort all M of them (6-7
    M-sized passes).  Both give bit-identical ids_sorted / tile_ranges.
    xy_from_geo: `xys` is the packed [V*n, 8] geo table (pixel centres in columns 0..1).
    while_waiting: callable run after the M read-back is enqueued and before the host waits for it."""
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
        # the one device->host read of the path (the reference has five per view): M sizes the sort.
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
            return Binning(n, n_views, 0, torch.empty((0,), dtype=torch.int32, device=dev), ranges, tuple(tile_bounds), None)
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
    return Binning(n, n_views, m, ids_sorted, ranges, tuple(tile_bounds), order_t)


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
    if binning.ids_sorted.numel() > 0:
        return ptr(binning.ids_sorted)
    return ptr(workspace(binning.tile_ranges.device).total)


def max_channels() -> int:
    return int(_lib.load().gg_blend_max_channels())


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
    step = max_channels()
    hit_words = None
    if record_hits and C <= step and binning.num_intersects > 0:
        nwords = int(_lib.load().gg_blend_hit_words(binning.num_intersects, binning.tile_ranges.shape[0], C))
        hit_words = torch.zeros(nwords, dtype=torch.int32, device=dev)
    with _lib.device_guard(dev):
        for c0 in range(0, C, step):
            c1 = min(C, c0 + step)
            _lib.call("gg_blend_fwd",
                V, n, c1 - c0, C, 1 if colors_per_view else 0, C, int(img_height), int(img_width), tb[0], tb[1],
                _ids_ptr(binning), ptr(binning.tile_ranges), ptr(binning.tile_order), ptr(geo),
                colors.data_ptr() + 4 * c0, background.data_ptr() + 4 * c0, out.data_ptr() + 4 * c0, ptr(final_T),
                ptr(final_idx), ptr(pair_counter) if c0 == 0 else None, ptr(hit_words), stream_ptr(dev))
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
    step = max_channels()
    if C > step:
        hit_words = None
    with _lib.device_guard(dev):
        for c0 in range(0, C, step):
            c1 = min(C, c0 + step)
            _lib.call("gg_blend_bwd",
                V, n, c1 - c0, C, 1 if colors_per_view else 0, C, int(img_height), int(img_width), tb[0], tb[1],
                _ids_ptr(binning), ptr(binning.tile_ranges), ptr(binning.tile_order), ptr(geo),
                colors.data_ptr() + 4 * c0, background.data_ptr() + 4 * c0, ptr(final_T), ptr(final_idx),
                v_out.data_ptr() + 4 * c0, ptr(hit_words), ptr(v_geo), v_colors.data_ptr() + 4 * c0, stream_ptr(dev))
    return v_geo, v_colors


def unpack_vgeo(n, n_views, v_geo):
    dev = v_geo.device
    v_xys = torch.empty((n_views * n, 2), dtype=torch.float32, device=dev)
    v_conics = torch.empty((n_views * n, 3), dtype=torch.float32, device=dev)
    v_opac = torch.empty((n,), dtype=torch.float32, device=dev)
    with _lib.device_guard(dev):
        _lib.call("gg_unpack_vgeo", int(n), int(n_views), ptr(v_geo), ptr(v_xys), ptr(v_conics), ptr(v_opac), 0,
                                         stream_ptr(dev))
    return v_xys, v_conics, v_opac
