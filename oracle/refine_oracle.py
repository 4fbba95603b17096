"""CPU restatement (torch, fp32) of the reference's densify / cull step and its Adam-state surgery.

TEST INFRASTRUCTURE ONLY -- imported by tests/ (and nothing else); the product path never touches it.
PINNED TO THE REFERENCE: tests/golden/refine_small.npz is the output of the reference's own
GaussianSplattingModel.refinement_after on real torch.optim.Adam objects (tests/golden/make_reference_golden.py, run
where /root/reference is mounted); this file reproduces it bit for bit (tests/test_golden.py).  It follows the
reference source line by line (nerfstudio/models/gaussian_splatting.py):
  refinement_after  :396-464     split_gaussians :485-518     dup_gaussians :520-533
  cull_gaussians    :466-483     dup_in_optim    :352-371     remove_from_optim :333-350
Parameters use this repository's names (log_scales = model.scales, opacity_logit = model.opacities,
sh_coeffs = model.colors_all, features = model.feature); opacity_logit is [N] here, [N,1] there.
"""
from __future__ import annotations

from typing import Dict, Tuple

import torch

PARAMS = ("means", "log_scales", "quats", "opacity_logit", "sh_coeffs", "features")


def quat_to_rotmat(q: torch.Tensor) -> torch.Tensor:
    q = torch.nn.functional.normalize(q, dim=-1)
    w, x, y, z = q.unbind(-1)
    return torch.stack([1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y),
                        2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x),
                        2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)], dim=-1).reshape(-1, 3, 3)


def refine(params: Dict[str, torch.Tensor], moments: Dict[str, Tuple[torch.Tensor, torch.Tensor]],
           xys_grad_norm, vis_counts, max_2dsize, cfg: dict, samples_fn, n_split_samples: int = 2):
    """Returns (new params, new moments, info).  cfg keys = fields of gg_refine_config.
    samples_fn(k) must return the [k, 3] standard-normal draws used for the split children (:491).
    info["fragile"] counts decisions within 1e-5 (relative) of a threshold."""
    P = {k: params[k].clone() for k in PARAMS}
    M = {k: (moments[k][0].clone(), moments[k][1].clone()) for k in PARAMS}
    n = P["means"].shape[0]
    m2d = max_2dsize.clone() if max_2dsize is not None else torch.zeros(n)
    fragile = 0

    def near(v, t):
        return int(((v - t).abs() <= 1e-5 * abs(t)).sum())

    if cfg["do_densify"]:
        avg = (xys_grad_norm / vis_counts) * 0.5 * cfg["max_dim"]                                  # :411-413
        high = avg > cfg["densify_grad_thresh"]                                                     # :414
        smax = P["log_scales"].exp().max(dim=-1).values
        fragile += near(avg, cfg["densify_grad_thresh"]) + near(smax, cfg["densify_size_thresh"])
        splits = smax > cfg["densify_size_thresh"]                                                  # :415
        if cfg["split_by_screen"]:
            splits = splits | (m2d > cfg["split_screen_size"])                                      # :416-417
            fragile += near(m2d, cfg["split_screen_size"])
        splits = splits & high                                                                      # :418
        # split_gaussians (:485-518)
        n_splits = int(splits.sum())
        z = samples_fn(n_split_samples * n_splits)                                                  # :491
        scaled = P["log_scales"][splits].repeat(n_split_samples, 1).exp() * z                       # :492-494
        q = P["quats"][splits] / P["quats"][splits].norm(dim=-1, keepdim=True)                      # :495
        rots = quat_to_rotmat(q.repeat(n_split_samples, 1))                                         # :496
        new_means = torch.bmm(rots, scaled[..., None]).squeeze(-1) + P["means"][splits].repeat(n_split_samples, 1)
        shrunk = torch.log(torch.exp(P["log_scales"][splits]) / 1.6)                                # :512-513
        split_new = dict(means=new_means, sh_coeffs=P["sh_coeffs"][splits].repeat(n_split_samples, 1, 1),
                         features=P["features"][splits].repeat(n_split_samples, 1),
                         opacity_logit=P["opacity_logit"][splits].repeat(n_split_samples),
                         log_scales=shrunk.repeat(n_split_samples, 1), quats=P["quats"][splits].repeat(n_split_samples, 1))
        P["log_scales"][splits] = shrunk                                                            # :514 in place
        # dups are decided on the scales AFTER the in-place shrink (:419-420)
        smax2 = P["log_scales"].exp().max(dim=-1).values
        fragile += near(smax2[splits], cfg["densify_size_thresh"])
        dups = (smax2 <= cfg["densify_size_thresh"]) & high
        dup_new = {k: P[k][dups] for k in PARAMS}                                                   # :520-533
        for k in PARAMS:
            P[k] = torch.cat([P[k], split_new[k], dup_new[k]], dim=0)                               # :422-427
        m2d = torch.cat([m2d, torch.zeros(split_new["means"].shape[0]), torch.zeros(dup_new["means"].shape[0])])
        for k in PARAMS:                                                                             # dup_in_optim
            ea, es = M[k]
            for idx, rep in ((splits, n_split_samples), (dups, 1)):                                 # :435-446
                dims = (rep,) + (1,) * (ea.dim() - 1)
                ea = torch.cat([ea, torch.zeros_like(ea[:n][idx]).repeat(*dims)], dim=0)
                es = torch.cat([es, torch.zeros_like(es[:n][idx]).repeat(*dims)], dim=0)
            M[k] = (ea, es)
    n_cat = P["means"].shape[0]
    culls = torch.zeros(n_cat, dtype=torch.bool)
    if cfg["do_cull"]:                                                                              # cull_gaussians
        op = torch.sigmoid(P["opacity_logit"])
        culls = op < cfg["cull_alpha_thresh"]                                                       # :472
        fragile += near(op, cfg["cull_alpha_thresh"])
        if cfg["cull_by_scale"]:
            smax = torch.exp(P["log_scales"]).max(dim=-1).values
            culls = culls | (smax > cfg["cull_scale_thresh"])                                       # :475-476
            fragile += near(smax, cfg["cull_scale_thresh"])
            if cfg["cull_by_screen"]:
                culls = culls | (m2d > cfg["cull_screen_size"])                                     # :477-480
                fragile += near(m2d, cfg["cull_screen_size"])
        for k in PARAMS:
            P[k] = P[k][~culls]                                                                     # :481-486
            M[k] = (M[k][0][~culls], M[k][1][~culls])                                               # remove_from_optim
    return P, M, dict(n_in=n, n_cat=n_cat, n_out=P["means"].shape[0], n_culled=int(culls.sum()), fragile=fragile)
