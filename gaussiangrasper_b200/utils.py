"""compute_cumulative_intersects / bin_and_sort_gaussians -- same surface as gsplat.utils
(run inside every rasterize forward upstream; the integer outputs are what the parity tests
compare bit for bit)."""
from typing import Tuple

import torch

from . import ops


def compute_cumulative_intersects(num_tiles_hit: torch.Tensor) -> Tuple[int, torch.Tensor]:
    """-> (num_intersects, cum_tiles_hit[N] int32).  One device->host read, as upstream."""
    ws = ops.workspace(num_tiles_hit.device)
    cum = ops.cumsum_i32(num_tiles_hit.to(torch.int32).contiguous(), ws.total)
    return int(ws.total.item()), cum


def bin_and_sort_gaussians(num_points: int, num_intersects: int, xys, depths, radii, cum_tiles_hit,
                           tile_bounds: Tuple[int, int, int]):
    """-> (isect_ids_unsorted[M] i64, gaussian_ids_unsorted[M] i32, isect_ids_sorted[M] i64,
           gaussian_ids_sorted[M] i32, tile_bins[T,2] i32)."""
    dev = xys.device
    m = int(num_intersects)
    num_tiles = int(tile_bounds[0]) * int(tile_bounds[1])
    keys = torch.empty((max(m, 1),), dtype=torch.int64, device=dev)
    ids = torch.empty((max(m, 1),), dtype=torch.int32, device=dev)
    keys_sorted = torch.empty_like(keys)
    ids_sorted = torch.empty_like(ids)
    if m > 0:
        ops.map_to_intersects(num_points, 1, ops.f32c(xys), ops.f32c(depths), radii.contiguous(),
                              cum_tiles_hit.contiguous(), tile_bounds, keys, ids)
        ops.sort_pairs(m, ops.key_bits_for(num_tiles), keys, ids, keys_sorted, ids_sorted)
        tile_bins = ops.tile_ranges(m, keys_sorted, num_tiles)
    else:
        tile_bins = torch.zeros((num_tiles, 2), dtype=torch.int32, device=dev)
    return keys[:m], ids[:m], keys_sorted[:m], ids_sorted[:m], tile_bins
