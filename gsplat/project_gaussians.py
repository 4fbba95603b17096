"""Drop-in module path `gsplat.project_gaussians` (imported by nerfstudio/models/gaussian_splatting.py:46-50),
backed by gaussiangrasper_b200."""
from gaussiangrasper_b200.project_gaussians import *  # noqa: F401,F403
