"""`gsplat` import shim: GaussianGrasper's model imports gsplat.project_gaussians, gsplat.rasterize,
gsplat.nd_rasterize, gsplat.sh and gsplat._torch_impl (nerfstudio/models/gaussian_splatting.py:46-50);
with this directory on sys.path those imports resolve to the B200-native implementation."""
from gaussiangrasper_b200 import (NDRasterizeGaussians, ProjectGaussians, RasterizeGaussians,  # noqa: F401
                                  SphericalHarmonics, num_sh_bases, quat_to_rotmat)

__version__ = "0.1.0"
