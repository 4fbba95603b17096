"""Pure-PyTorch helpers the reference imports from gsplat._torch_impl
(nerfstudio/models/gaussian_splatting.py:46, used at :516 and :614; scripts/update.py:74).

Only `quat_to_rotmat` is on the reference's import path; it is differentiable torch code
upstream as well (it feeds the normal channel through autograd), so it stays torch here.
"""
import torch
import torch.nn.functional as F


def quat_to_rotmat(quat: torch.Tensor) -> torch.Tensor:
    """(…,4) wxyz quaternion (normalised here) -> (…,3,3) row-major rotation matrix."""
    assert quat.shape[-1] == 4, quat.shape
    w, x, y, z = torch.unbind(F.normalize(quat, dim=-1), dim=-1)
    mat = torch.stack(
        [
            1 - 2 * (y**2 + z**2), 2 * (x * y - w * z), 2 * (x * z + w * y),
            2 * (x * y + w * z), 1 - 2 * (x**2 + z**2), 2 * (y * z - w * x),
            2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x**2 + y**2),
        ],
        dim=-1,
    )
    return mat.reshape(quat.shape[:-1] + (3, 3))
