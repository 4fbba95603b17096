"""Drop-in module path `gsplat._torch_impl` (imported by nerfstudio/models/gaussian_splatting.py:46-50),
backed by gaussiangrasper_b200."""
from gaussiangrasper_b200._torch_impl import *  # noqa: F401,F403
from gaussiangrasper_b200._torch_impl import quat_to_rotmat  # noqa: F401
