"""Drop-in module path `gsplat.nd_rasterize` (imported by nerfstudio/models/gaussian_splatting.py:46-50),
backed by gaussiangrasper_b200."""
from gaussiangrasper_b200.nd_rasterize import *  # noqa: F401,F403
