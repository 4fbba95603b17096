"""The reference model's per-view render, call for call (nerfstudio/models/gaussian_splatting.py:699-784),
written against the drop-in operator classes: 1 ProjectGaussians + 1 SphericalHarmonics + 4 rasterizations
with the model's torch glue in between.  Used by bench.py (`--path dropin`) and the tests to exercise and
time exactly what an unpatched GaussianGrasper would issue."""
import torch

from ._torch_impl import quat_to_rotmat
from .nd_rasterize import NDRasterizeGaussians
from .project_gaussians import ProjectGaussians
from .rasterize import RasterizeGaussians
from .sh import SphericalHarmonics


def get_outputs(means, log_scales, quats, opacity_logit, sh_coeffs, features, cam, sh_degree_to_use=4):
    dev = means.device
    tb = cam.tile_bounds
    xys, depths, radii, conics, num_tiles_hit, cov3d = ProjectGaussians.apply(
        means, torch.exp(log_scales), 1, quats / quats.norm(dim=-1, keepdim=True), cam.viewmat[:3].to(dev),
        cam.fullmat.to(dev), cam.fx, cam.fy, cam.cx, cam.cy, cam.H, cam.W, tb)                        # :699
    if xys.requires_grad:
        xys.retain_grad()                                                                             # :725
    viewdirs = means.detach() - cam.position.to(dev)                                                  # :727
    viewdirs = viewdirs / viewdirs.norm(dim=-1, keepdim=True)
    rgbs = torch.clamp(SphericalHarmonics.apply(sh_degree_to_use, viewdirs, sh_coeffs) + 0.5, 0.0, 1.0)  # :730
    rgb = RasterizeGaussians.apply(xys, depths, radii, conics, num_tiles_hit, rgbs, torch.sigmoid(opacity_logit),
                                   cam.H, cam.W, torch.zeros(3, device=dev))                           # :735
    feature = NDRasterizeGaussians.apply(xys, depths, radii, conics, num_tiles_hit, features,
                                         torch.sigmoid(opacity_logit), cam.H, cam.W,
                                         torch.zeros(features.shape[1], device=dev))                   # :747
    depth_im = RasterizeGaussians.apply(xys, depths, radii, conics, num_tiles_hit, depths[:, None].repeat(1, 3),
                                        torch.sigmoid(opacity_logit), cam.H, cam.W,
                                        torch.ones(3, device=dev) * 10)[..., 0:1]                      # :759
    R = quat_to_rotmat(quats)                                                                         # :614
    idx = torch.exp(log_scales).min(dim=-1)[1][..., None, None].expand(-1, 3, -1)
    normals = R.gather(2, idx).squeeze(dim=2)
    normal_im = RasterizeGaussians.apply(xys, depths, radii, conics, num_tiles_hit, normals,
                                         torch.sigmoid(opacity_logit), cam.H, cam.W, torch.zeros(3, device=dev))  # :773
    return dict(rgb=rgb, feature=feature, depth=depth_im, normal=normal_im, xys=xys, radii=radii)
