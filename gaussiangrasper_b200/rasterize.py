"""RasterizeGaussians -- same call signature as gsplat.rasterize.RasterizeGaussians
(reference call sites: nerfstudio/models/gaussian_splatting.py:735-746 rgb, :759-770 depth,
:773-784 normals)."""
import torch
from torch.autograd import Function

from . import _raster


class RasterizeGaussians(Function):
    """apply(xys[N,2], depths[N], radii[N], conics[N,3], num_tiles_hit[N], colors[N,3],
             opacity[N,1], img_height, img_width, background[3]=ones) -> out_img[H,W,3]"""

    @staticmethod
    def forward(ctx, xys, depths, radii, conics, num_tiles_hit, colors, opacity, img_height, img_width,
                background=None):
        if colors.ndim != 2 or colors.shape[1] != 3:
            raise ValueError("colors must have dimensions (N, 3)")
        return _raster.rasterize_forward(ctx, xys, depths, radii, conics, num_tiles_hit, colors, opacity,
                                         img_height, img_width, background)

    @staticmethod
    def backward(ctx, v_out_img):
        v_xys, v_conics, v_colors, v_opacity = _raster.rasterize_backward(ctx, v_out_img)
        return (v_xys, None, None, v_conics, None, v_colors, v_opacity, None, None, None)
