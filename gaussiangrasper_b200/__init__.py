"""B200-native (sm_100a) Gaussian-splatting rasterizer: a drop-in for the gsplat 0.1.0 operators
GaussianGrasper's nerfstudio model calls (ProjectGaussians, SphericalHarmonics,
RasterizeGaussians, NDRasterizeGaussians, quat_to_rotmat), plus a fused multi-view entry point.

Hand-written CUDA behind a C ABI (include/gg_b200.h, libgg_b200.so); no CPU fallback.
"""
from ._torch_impl import quat_to_rotmat
from .nd_rasterize import NDRasterizeGaussians
from .project_gaussians import ProjectGaussians
from .rasterize import RasterizeGaussians
from .sh import SphericalHarmonics, num_sh_bases

__all__ = [
    "ProjectGaussians", "RasterizeGaussians", "NDRasterizeGaussians", "SphericalHarmonics", "num_sh_bases",
    "quat_to_rotmat",
]
__version__ = "0.1.0"
