#!/usr/bin/env python
"""Generates tests/golden/small_scene.npz with the CPU oracle (oracle/gg_oracle.c).

PARITY UNPINNED: the reference holds no golden vector for this path and gsplat 0.1.0 cannot be run
here, so these vectors pin the ORACLE (and through it the CUDA path) against
regressions; they are not outputs of the reference itself.

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from gaussiangrasper_b200 import scenes  # noqa: E402
from oracle import c_oracle  # noqa: E402

N, W, H, D, SEED = 400, 64, 48, 4, 2024


def build():
    sc = scenes.random_scene(N, feature_dim=D, seed=SEED)
    sc["log_scales"] = sc["log_scales"] + 1.2
    cam = scenes.look_at_camera((4.5, 0.3, 0.2), W, H)
    scales = sc["log_scales"].exp().numpy()
    quats = (sc["quats"] / sc["quats"].norm(dim=-1, keepdim=True)).numpy()
    means = sc["means"].numpy()
    xys, depths, radii, conics, nth, cov3d = c_oracle.project_fwd(
        means, scales, 1.0, quats, cam.viewmat[:3].numpy(), cam.fullmat.numpy(), cam.fx, cam.fy, cam.cx, cam.cy, H, W,
        cam.tile_bounds)
    cum, keys, ids, keys_s, ids_s, ranges = c_oracle.bin_and_sort(xys, depths, radii, nth, cam.tile_bounds)
    dirs = means - cam.position.numpy()
    sh = sc["sh_coeffs"].numpy()
    rgb = np.clip(c_oracle.sh_fwd(4, dirs, sh) + 0.5, 0, 1).astype(np.float32)
    opac = torch.sigmoid(sc["opacity_logit"]).numpy().reshape(-1)
    cols = np.concatenate([rgb, depths[:, None], sc["features"].numpy()], axis=1).astype(np.float32)
    bg = np.array([0, 0, 0, 10] + [0] * D, np.float32)
    out, fT, fidx, frag, pairs = c_oracle.blend_fwd(H, W, cam.tile_bounds, ids_s, ranges, xys, conics, opac, cols, bg,
                                                    eps=2e-5)
    g = torch.Generator().manual_seed(1)
    v_out = torch.randn((H, W, cols.shape[1]), generator=g).numpy()
    v_xy, v_conic, v_cols, v_opac = c_oracle.blend_bwd(H, W, cam.tile_bounds, ids_s, ranges, xys, conics, opac, cols, bg,
                                                       v_out)
    return dict(
        means=means, scales=scales, quats=quats, viewmat=cam.viewmat.numpy(), fullmat=cam.fullmat.numpy(),
        intrinsics=np.array([cam.fx, cam.fy, cam.cx, cam.cy], np.float64), size=np.array([H, W]), dirs=dirs, sh=sh,
        opac=opac, cols=cols, bg=bg, v_out=v_out,
        xys=xys, depths=depths, radii=radii, conics=conics, num_tiles_hit=nth, cov3d=cov3d, cum=cum,
        keys_unsorted=keys, ids_unsorted=ids, keys_sorted=keys_s, ids_sorted=ids_s, tile_ranges=ranges, rgb=rgb,
        image=out, final_T=fT, final_idx=fidx, fragile=frag, pairs=np.array([pairs]),
        v_xy=v_xy, v_conic=v_conic, v_cols=v_cols, v_opac=v_opac)


if __name__ == "__main__":
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "small_scene.npz")
    np.savez_compressed(path, **build())
    print(path, os.path.getsize(path), "bytes")
