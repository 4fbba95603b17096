"""Shared implementation of RasterizeGaussians / NDRasterizeGaussians.

The reference calls the rasterizers four times per view with the same projection outputs
(nerfstudio/models/gaussian_splatting.py:735,747,759,773) and upstream re-bins and re-sorts in
each call.  Here the binning of a view is computed once and reused while the projection
tensors (xys, depths, radii, num_tiles_hit) are the same storage at the same version.
"""
from __future__ import annotations

import torch

from . import ops

_bin_cache = {}


def _tensor_key(t: torch.Tensor):
    return (t.data_ptr(), t._version, tuple(t.shape))


def binning_for(xys, depths, radii, num_tiles_hit, img_height, img_width) -> ops.Binning:
    dev = xys.device
    key = (_tensor_key(xys), _tensor_key(depths), _tensor_key(radii), _tensor_key(num_tiles_hit),
           int(img_height), int(img_width), torch.cuda.current_stream(dev).cuda_stream)
    slot = (dev.type, dev.index)
    hit = _bin_cache.get(slot)
    if hit is not None and hit[0] == key:
        return hit[1]
    tile_bounds = ops.tile_bounds_for(img_height, img_width)
    # the drop-in signatures hand the image back to code that expects it to be valid at once: exact sizing
    # (one host wait for the intersection count, where the reference has one per rasterize call)
    binning = ops.bin_views(xys.shape[0], 1, xys.detach(), depths.detach(), radii, num_tiles_hit, tile_bounds,
                            sync_free=False)
    # the cache entry keeps the keyed tensors alive so their addresses cannot be recycled under it
    _bin_cache[slot] = (key, binning, (xys.detach(), depths.detach(), radii, num_tiles_hit))
    return binning


def clear_cache():
    _bin_cache.clear()


def rasterize_forward(ctx, xys, depths, radii, conics, num_tiles_hit, colors, opacity, img_height, img_width,
                      background):
    if colors.dtype == torch.uint8:
        colors = colors.float() / 255
    if xys.ndim != 2 or xys.shape[1] != 2:
        raise ValueError("xys must have dimensions (N, 2)")
    n = xys.shape[0]
    if colors.ndim != 2 or colors.shape[0] != n:
        raise ValueError("colors must have dimensions (N, C)")
    channels = colors.shape[1]
    if background is None:
        background = torch.ones(channels, dtype=torch.float32, device=colors.device)
    assert background.shape[0] == channels, "Incorrect shape of background color tensor"
    if opacity.numel() != n:
        raise ValueError("opacity must have N entries")
    binning = binning_for(xys, depths, radii, num_tiles_hit, img_height, img_width)
    geo = ops.pack_geo(n, 1, xys.detach(), conics.detach(), opacity.detach())
    colors_c = ops.f32c(colors.detach())
    background = ops.f32c(background.detach())
    out, final_T, final_idx, hit_words = ops.blend_fwd(binning, geo, colors_c, background, img_height, img_width,
                                                       single_image=True)
    ctx.hit_words = hit_words
    ctx.binning = binning
    ctx.img_size = (int(img_height), int(img_width))
    ctx.opacity_shape = tuple(opacity.shape)
    # only inputs and the transmittance / last-contributor maps are saved (never the image: the
    # model writes into it in place before backward, gaussian_splatting.py:884)
    ctx.save_for_backward(geo, colors_c, background, final_T, final_idx)
    return out


def rasterize_backward(ctx, v_out):
    geo, colors, background, final_T, final_idx = ctx.saved_tensors
    binning = ctx.binning
    h, w = ctx.img_size
    v_geo, v_colors = ops.blend_bwd(binning, geo, colors, background, final_T, final_idx,
                                    v_out.reshape(1, h, w, -1), h, w, hit_words=ctx.hit_words)
    v_xys, v_conics, v_opac = ops.unpack_vgeo(binning.n, 1, v_geo)
    return v_xys, v_conics, v_colors, v_opac.reshape(ctx.opacity_shape)
