"""NDRasterizeGaussians -- same call signature as gsplat.nd_rasterize.NDRasterizeGaussians
(reference call site: nerfstudio/models/gaussian_splatting.py:747-758, the 32-channel latent
language-feature map)."""
import torch
from torch.autograd import Function

from . import _raster


class NDRasterizeGaussians(Function):
    """apply(xys[N,2], depths[N], radii[N], conics[N,3], num_tiles_hit[N], colors[N,C],
             opacity[N,1], img_height, img_width, background[C]=ones) -> out_img[H,W,C]"""

    @staticmethod
    def forward(ctx, xys, depths, radii, conics, num_tiles_hit, colors, opacity, img_height, img_width,
                background=None):
        return _raster.rasterize_forward(ctx, xys, depths, radii, conics, num_tiles_hit, colors, opacity,
                                         img_height, img_width, background)

    @staticmethod
    def backward(ctx, v_out_img):
        v_xys, v_conics, v_colors, v_opacity = _raster.rasterize_backward(ctx, v_out_img)
        return (v_xys, None, None, v_conics, None, v_colors, v_opacity, None, None, None)
