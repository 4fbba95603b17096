#!/usr/bin/env python
"""Small end-to-end run (drop-in 4-call path + fused path, fwd+bwd, odd sizes) for compute-sanitizer."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gaussiangrasper_b200 import (NDRasterizeGaussians, ProjectGaussians, RasterizeGaussians, SphericalHarmonics, scenes)
from gaussiangrasper_b200.render import ViewBatch, render_views

dev = torch.device("cuda:0")
for n, W, H, D, V in ((777, 70, 45, 5, 2), (4099, 131, 97, 16, 3)):
    sc = scenes.random_scene(n, feature_dim=D, seed=n)
    sc["log_scales"] = sc["log_scales"] + 0.9
    cams = scenes.orbit_cameras(V, W, H, total=5)
    P = {k: v.to(dev).requires_grad_(True) for k, v in sc.items()}
    out = render_views(P["means"], P["log_scales"], P["quats"], P["opacity_logit"], P["sh_coeffs"], P["features"],
                       ViewBatch.from_cameras(cams, dev))
    out["image"].sum().backward()
    cam = cams[0]
    q = P["quats"] / P["quats"].norm(dim=-1, keepdim=True)
    xys, depths, radii, conics, nth, cov3d = ProjectGaussians.apply(
        P["means"], torch.exp(P["log_scales"]), 1, q, cam.viewmat[:3].to(dev), cam.fullmat.to(dev), cam.fx, cam.fy, cam.cx,
        cam.cy, H, W, cam.tile_bounds)
    rgbs = torch.clamp(SphericalHarmonics.apply(3, P["means"].detach() - cam.position.to(dev), P["sh_coeffs"]) + 0.5, 0, 1)
    op = torch.sigmoid(P["opacity_logit"])
    a = RasterizeGaussians.apply(xys, depths, radii, conics, nth, rgbs, op, H, W, torch.zeros(3, device=dev))
    b = NDRasterizeGaussians.apply(xys, depths, radii, conics, nth, P["features"], op, H, W, torch.zeros(D, device=dev))
    (a.sum() + b.sum() + depths.sum()).backward()
    torch.cuda.synchronize()
    print("ok", n, W, H, D, V, float(a.mean()), float(b.mean()))
