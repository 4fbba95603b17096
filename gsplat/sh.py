"""Drop-in module path `gsplat.sh` (imported by nerfstudio/models/gaussian_splatting.py:46-50),
backed by gaussiangrasper_b200."""
from gaussiangrasper_b200.sh import *  # noqa: F401,F403
