"""Synthetic workloads of BASELINE.json: seeded random Gaussian scenes and the
reference's camera conventions.  Everything is generated on the CPU with a fixed generator so
that the CPU oracle and the CUDA path see identical inputs.

Camera conventions follow nerfstudio/models/gaussian_splatting.py:87-105 (projection matrix,
w_clip = z_cam) and :658-676 (world->camera view matrix, +z forward, +y down); intrinsics are the
RealSense values of scripts/generate_data.py:88-89 scaled to the requested resolution.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Dict, List

import torch

REALSENSE = dict(W=640, H=480, fx=385.86, fy=385.38, cx=325.68, cy=243.56)

# BASELINE.json configs -> concrete inputs
CONFIGS = {
    0: dict(name="cfg0_50k_640x480_rgbd_fwd_cpu", n=50_000, W=640, H=480, views=1, D=0, normal=False, backward=False),
    1: dict(name="cfg1_500k_640x480_rgb+depth+normal+16ch_fwd+bwd", n=500_000, W=640, H=480, views=1, D=16,
            normal=True, backward=True),
    2: dict(name="cfg2_2M_1280x720_64views_fwd", n=2_000_000, W=1280, H=720, views=64, D=16, normal=True,
            backward=False),
    3: dict(name="cfg3_1M_640x480_8views_per_gpu_fwd+bwd_allreduce", n=1_000_000, W=640, H=480, views=8, D=16,
            normal=True, backward=True),
    4: dict(name="cfg4_1M_1920x1080_Dsweep", n=1_000_000, W=1920, H=1080, views=1, D=16, normal=True, backward=True),
}


def projection_matrix(znear: float, zfar: float, fovx: float, fovy: float) -> torch.Tensor:
    """OpenGL-style perspective matrix of the reference (gaussian_splatting.py:87-105)."""
    t = znear * math.tan(0.5 * fovy)
    b = -t
    r = znear * math.tan(0.5 * fovx)
    l = -r
    n, f = znear, zfar
    return torch.tensor(
        [[2 * n / (r - l), 0.0, (r + l) / (r - l), 0.0],
         [0.0, 2 * n / (t - b), (t + b) / (t - b), 0.0],
         [0.0, 0.0, (f + n) / (f - n), -1.0 * f * n / (f - n)],
         [0.0, 0.0, 1.0, 0.0]], dtype=torch.float32)


@dataclass
class Camera:
    viewmat: torch.Tensor   # [4,4] world->camera
    fullmat: torch.Tensor   # [4,4] projmat @ viewmat
    fx: float
    fy: float
    cx: float
    cy: float
    H: int
    W: int
    position: torch.Tensor  # [3] camera centre in world coordinates

    @property
    def tile_bounds(self):
        return ((self.W + 15) // 16, (self.H + 15) // 16, 1)


def intrinsics_for(W: int, H: int):
    sx, sy = W / REALSENSE["W"], H / REALSENSE["H"]
    # keep pixels square: fx, fy, cx scale with W; cy with H
    return REALSENSE["fx"] * sx, REALSENSE["fy"] * sx, REALSENSE["cx"] * sx, REALSENSE["cy"] * sy


def look_at_camera(position, W: int, H: int, target=(0.0, 0.0, 0.0)) -> Camera:
    pos = torch.tensor(position, dtype=torch.float64)
    tgt = torch.tensor(target, dtype=torch.float64)
    z = tgt - pos
    z = z / z.norm()
    down = torch.tensor([0.0, 1.0, 0.0], dtype=torch.float64)
    x = torch.linalg.cross(down, z)
    x = x / x.norm()
    y = torch.linalg.cross(z, x)
    R = torch.stack([x, y, z], dim=0)  # rows: camera axes in world coordinates
    viewmat = torch.eye(4, dtype=torch.float64)
    viewmat[:3, :3] = R
    viewmat[:3, 3] = -R @ pos
    viewmat = viewmat.float()
    fx, fy, cx, cy = intrinsics_for(W, H)
    fovx = 2 * math.atan(W / (2 * fx))
    fovy = 2 * math.atan(H / (2 * fy))
    projmat = projection_matrix(0.001, 1000.0, fovx, fovy)
    return Camera(viewmat, projmat @ viewmat, fx, fy, cx, cy, H, W, pos.float())


def downscale_factor(step: int, num_downscales: int = 1, resolution_schedule: int = 250) -> int:
    """The resolution warm-up of training (gaussian_splatting.py:599-603): renders are 2^d times smaller until
    step reaches d * resolution_schedule."""
    return 2 ** max(num_downscales - step // resolution_schedule, 0)


def camera_from_c2w(camera_to_world, fx: float, fy: float, cx: float, cy: float, W: int, H: int,
                    downscale: int = 1) -> Camera:
    """The camera set-up of GaussianSplattingModel.get_outputs (gaussian_splatting.py:657-678) for one nerfstudio
    camera: `camera_to_world` [3,4] (or [4,4]) in nerfstudio's convention (+x right, +y up, -z forward).  The y and z
    axes are flipped (the reference multiplies by SO3.from_x_radians(pi): diag(1,-1,-1) up to 1e-16), the pose is
    inverted analytically, the fields of view come from the focal lengths and the projection is
    projection_matrix(0.001, 1000, fovx, fovy).  `downscale` > 1 is the training warm-up (downscale_factor): the
    reference rescales the camera by 1 / downscale first (:656, Cameras.rescale_output_resolution -- fp32 products,
    sizes truncated)."""
    if downscale != 1:
        sc = torch.tensor([1 / downscale])                    # fp32, like the reference's scaling tensor
        fx, fy, cx, cy = ((torch.tensor(float(v), dtype=torch.float32) * sc).item() for v in (fx, fy, cx, cy))
        W, H = int((torch.tensor(int(W)) * sc).to(torch.int64)), int((torch.tensor(int(H)) * sc).to(torch.int64))
    c2w = torch.as_tensor(camera_to_world, dtype=torch.float32).detach().cpu()
    R = c2w[:3, :3] * torch.tensor([1.0, -1.0, -1.0])        # R @ diag(1,-1,-1): columns y, z negated       (:662-663)
    T = c2w[:3, 3:4]
    viewmat = torch.eye(4, dtype=torch.float32)               # analytic inverse                              (:665-669)
    viewmat[:3, :3] = R.T
    viewmat[:3, 3:4] = -R.T @ T
    # nerfstudio keeps the intrinsics as fp32 tensors: `.item()` hands gsplat the rounded values, and the quotient
    # under the atan is an fp32 tensor operation (:671-674)
    f32 = lambda v: torch.tensor(float(v), dtype=torch.float32)
    fx, fy, cx, cy = (f32(v).item() for v in (fx, fy, cx, cy))
    fovx = 2 * math.atan((torch.tensor(int(W)) / (2 * f32(fx))).item())
    fovy = 2 * math.atan((torch.tensor(int(H)) / (2 * f32(fy))).item())
    projmat = projection_matrix(0.001, 1000, fovx, fovy)      #                                               (:677)
    return Camera(viewmat, projmat @ viewmat, float(fx), float(fy), float(cx), float(cy), int(H), int(W),
                  c2w[:3, 3].clone())                         # SH view directions start at the camera centre (:727)


def cameras_from_nerfstudio(cameras, downscale: int = 1) -> List[Camera]:
    """One Camera per entry of a nerfstudio `Cameras` object (anything with `camera_to_worlds` [V,3,4] and per-camera
    `fx, fy, cx, cy, width, height` tensors of shape [V,1]) -- what get_outputs reads from its `camera` argument."""
    c2w = cameras.camera_to_worlds
    if c2w.dim() == 2:
        c2w = c2w[None]
    pick = lambda t, i: float(torch.as_tensor(t).reshape(-1)[i if torch.as_tensor(t).numel() > 1 else 0])
    return [camera_from_c2w(c2w[i], pick(cameras.fx, i), pick(cameras.fy, i), pick(cameras.cx, i), pick(cameras.cy, i),
                            int(pick(cameras.width, i)), int(pick(cameras.height, i)), downscale)
            for i in range(c2w.shape[0])]


def orbit_cameras(n_views: int, W: int, H: int, radius: float = 4.5, first: int = 0, total: int = None) -> List[Camera]:
    """View k of `total` on a circle of radius 4.5 in the xz-plane at height 0.5*sin(2*pi*k/total)."""
    total = total or n_views
    cams = []
    for k in range(first, first + n_views):
        th = 2 * math.pi * k / total
        cams.append(look_at_camera((radius * math.cos(th), 0.5 * math.sin(th), radius * math.sin(th)), W, H))
    return cams


def random_scene(n: int, feature_dim: int = 16, seed: int = 1234, sh_degree: int = 4) -> Dict[str, torch.Tensor]:
    """Model-level parameters (pre-activation where the model stores them that way)."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    nb = (sh_degree + 1) ** 2
    means = (torch.rand((n, 3), generator=g) * 4.0 - 2.0)
    log_scales = torch.rand((n, 3), generator=g) * (math.log(0.05) - math.log(0.005)) + math.log(0.005)
    quats = torch.randn((n, 4), generator=g)
    opac = torch.rand((n, 1), generator=g) * 0.9 + 0.05
    opac_logit = torch.log(opac / (1 - opac))
    sh = torch.randn((n, nb, 3), generator=g) * 0.1
    sh[:, 0, :] = torch.rand((n, 3), generator=g) * 3.54 - 1.77
    feats = torch.rand((n, max(feature_dim, 1)), generator=g) * 2.0 - 1.0
    return dict(means=means, log_scales=log_scales, quats=quats, opacity_logit=opac_logit, sh_coeffs=sh,
                features=feats[:, :feature_dim] if feature_dim > 0 else feats[:, :0])
