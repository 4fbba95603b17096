"""Drop-in module path `gsplat.utils` (imported by nerfstudio/models/gaussian_splatting.py:46-50),
backed by gaussiangrasper_b200."""
from gaussiangrasper_b200.utils import *  # noqa: F401,F403
from gaussiangrasper_b200.utils import bin_and_sort_gaussians, compute_cumulative_intersects  # noqa: F401
