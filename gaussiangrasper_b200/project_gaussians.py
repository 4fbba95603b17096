"""ProjectGaussians -- same call signature as gsplat.project_gaussians.ProjectGaussians
(reference call site: nerfstudio/models/gaussian_splatting.py:699-713)."""
from typing import Tuple

import torch
from torch.autograd import Function

from . import ops


class ProjectGaussians(Function):
    """EWA projection of 3D Gaussians to the image plane.

    apply(means3d[N,3], scales[N,3], glob_scale, quats[N,4] wxyz, viewmat[>=3,4], projmat[4,4]
          (= projection @ viewmat), fx, fy, cx, cy, img_height, img_width, tile_bounds,
          clip_thresh=0.01)
      -> xys[N,2], depths[N], radii[N] int32, conics[N,3], num_tiles_hit[N] int32, cov3d[N,6]
    Gradients flow from xys, depths and conics to means3d, scales and quats.
    """

    @staticmethod
    def forward(ctx, means3d, scales, glob_scale, quats, viewmat, projmat, fx, fy, cx, cy, img_height, img_width,
                tile_bounds: Tuple[int, int, int], clip_thresh: float = 0.01):
        if means3d.ndim != 2 or means3d.shape[1] != 3 or means3d.shape[0] < 1:
            raise ValueError(f"means3d must have dimensions (N, 3), N >= 1; got {tuple(means3d.shape)}")
        n = means3d.shape[0]
        if scales.shape != (n, 3) or quats.shape != (n, 4):
            raise ValueError("scales must be (N, 3) and quats (N, 4)")
        xys, depths, radii, conics, num_tiles_hit, cov3d = ops.project_fwd(
            means3d, scales, glob_scale, quats, viewmat, projmat, fx, fy, cx, cy, img_height, img_width,
            tile_bounds, clip_thresh)
        ctx.img_height, ctx.img_width = int(img_height), int(img_width)
        ctx.glob_scale = float(glob_scale)
        ctx.intrinsics = (float(fx), float(fy), float(cx), float(cy))
        ctx.save_for_backward(means3d, scales, quats, viewmat, projmat, radii, conics)
        ctx.mark_non_differentiable(radii, num_tiles_hit, cov3d)
        return xys, depths, radii, conics, num_tiles_hit, cov3d

    @staticmethod
    def backward(ctx, v_xys, v_depths, v_radii, v_conics, v_num_tiles_hit, v_cov3d):
        means3d, scales, quats, viewmat, projmat, radii, conics = ctx.saved_tensors
        fx, fy, cx, cy = ctx.intrinsics
        if v_xys is None:
            v_xys = torch.zeros_like(conics[:, :2])
        if v_conics is None:
            v_conics = torch.zeros_like(conics)
        v_means, v_scales, v_quats = ops.project_bwd(
            means3d, scales, ctx.glob_scale, quats, viewmat, projmat, fx, fy, cx, cy, ctx.img_height, ctx.img_width,
            radii, conics, v_xys, v_depths, v_conics)
        return (v_means, v_scales, None, v_quats, None, None, None, None, None, None, None, None, None, None)
