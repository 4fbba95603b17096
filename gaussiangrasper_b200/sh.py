"""SphericalHarmonics / num_sh_bases -- same surface as gsplat.sh
(reference call site: nerfstudio/models/gaussian_splatting.py:730; num_sh_bases at :265,310)."""
import torch
from torch.autograd import Function

from . import ops


def num_sh_bases(degree: int) -> int:
    if degree == 0:
        return 1
    if degree == 1:
        return 4
    if degree == 2:
        return 9
    if degree == 3:
        return 16
    return 25


class SphericalHarmonics(Function):
    """apply(degrees_to_use, viewdirs[N,3], coeffs[N,K,3]) -> colors[N,3]; gradient to coeffs only."""

    @staticmethod
    def forward(ctx, degrees_to_use: int, viewdirs, coeffs):
        if coeffs.ndim != 3 or coeffs.shape[-1] != 3:
            raise ValueError(f"coeffs must be (N, K, 3); got {tuple(coeffs.shape)}")
        if viewdirs.shape != (coeffs.shape[0], 3):
            raise ValueError("viewdirs must be (N, 3)")
        degree = ops.sh_degree_from_bases(coeffs.shape[-2])
        if not 0 <= int(degrees_to_use) <= degree:
            raise ValueError(f"degrees_to_use={degrees_to_use} outside [0, {degree}]")
        ctx.degrees_to_use = int(degrees_to_use)
        ctx.degree = degree
        ctx.save_for_backward(viewdirs)
        return ops.sh_fwd(int(degrees_to_use), viewdirs, coeffs)

    @staticmethod
    def backward(ctx, v_colors):
        (viewdirs,) = ctx.saved_tensors
        return None, None, ops.sh_bwd(ctx.degree, ctx.degrees_to_use, viewdirs, v_colors)
