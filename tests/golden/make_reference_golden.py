#!/usr/bin/env python
"""Golden vectors produced by the REFERENCE ITSELF (not by this repository's oracle) for the rows of SURVEY 8-f whose
code lives under /root/reference: the unmodified nerfstudio/models/gaussian_splatting.py is imported (third-party
packages this code never calls are stubs, see tests/reference_model_driver.py) and its own methods are run on seeded
inputs.  Needs /root/reference (build container only); the resulting fixtures travel with the repository.

    python tests/golden/make_reference_golden.py [--out DIR]     # writes the three .npz files (default: beside this script)

  refine_small.npz       GaussianSplattingModel.refinement_after (:402-464: split_gaussians, dup_gaussians,
                         cull_gaussians, dup_in_optim, remove_from_optim on real torch.optim.Adam objects) on the
                         seeded inputs of make_refine_golden.inputs(), torch.randn serving the stored split samples.
                         (Until round 2 this file was written by oracle/refine_oracle.py; the reference's own output
                         turned out byte-identical to it, array by array.)
  ref_losses_small.npz   GaussianSplattingModel.get_loss_dict (:841-933) + backward on a synthetic 40x48 output dict;
                         the pixel samples its samplers drew (sampling_pairs_in_mask / sampling_in_mask, :120-148) are
                         recorded; `main_loss` goes through the SSIM restatement (pytorch_msssim is not installable)
                         and is stored as such
  ref_init_small.npz     populate_modules' k-nearest-neighbour scale initialisation (:259-263, k_nearest_sklearn
                         :315-331 with the real scikit-learn), the up-projection MLP (:198-213) forward on seeded
                         weights, projection_matrix (:87-105), SH2RGB (:80-85), the arguments get_outputs (:624-713) hands to
                         ProjectGaussians.apply for a few nerfstudio cameras, after_train's densification statistics (:373-393) over
                         three steps, the rules refinement_after applies at every refinement step of a run (observed on probe Gaussians),
                         nerfstudio's own real-SH basis (utils/math.py:29-92) and quaternion -> rotation
                         matrix (cameras/camera_utils.py:142-161), the optimizer table of the method
                         (configs/method_configs.py:611-664, read from the source's syntax tree) and the trainer's
                         ExponentialDecayScheduler (nerfstudio/engine/schedulers.py:109-140, imported and run) learning
                         rates for it
  ref_trainer_small.npz  Trainer.train_iteration (engine/trainer.py:458-499) itself, 25 iterations with the real Optimizers
                         and the method's gradient-accumulation table: parameters after every iteration
  ref_ply_small.npz      ExportGaussianSplat.main (scripts/exporter.py:482-530) itself on a small model; open3d is absent,
                         the attribute map it hands to o3d.t.geometry.PointCloud is recorded
  ref_targets_small.npz  the ground-truth side of get_loss_dict (:849-875) at half resolution (step 100): the method's own local
                         tensors (gt_img, gt_normal, gt_depth, depth_mask, gt_mask, valid_mask, gt_fea), read from its frame
  ref_outputs_small.npz  GaussianSplattingModel.get_outputs (:624-802) + backward itself, SH degree 4; the operator classes it
                         calls are this repository's with the CPU oracle underneath (gsplat is absent): pins the glue around
                         the operators to the reference's code
"""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
TESTS = os.path.dirname(HERE)
ROOT = os.path.dirname(TESTS)
REF = "/root/reference"
for p in (HERE, TESTS, ROOT, REF):
    if p not in sys.path:
        sys.path.insert(0, p)

import torch.fx  # noqa: F401,E402  (before dataclasses is patched by the stubs)
from reference_model_driver import install_stubs  # noqa: E402

PARAM_OF = dict(means="means", log_scales="scales", quats="quats", opacity_logit="opacities", sh_coeffs="colors_all",
                features="feature")
GROUP_OF = dict(means="xyz", sh_coeffs="color", opacity_logit="opacity", log_scales="scaling", quats="rotation",
                features="feature")


def import_reference():
    install_stubs()
    import nerfstudio.models.gaussian_splatting as gs
    return gs


def small_model(gs, n_seed=4000, num_train_data=4):
    """The reference's constructor, with its hard-coded 500 000 random seed points cut to n_seed."""
    from nerfstudio.data.scene_box import SceneBox
    orig_rand = torch.rand

    def small_rand(*size, **kw):
        if size and size[0] == (500000, 3):
            return orig_rand((n_seed, 3), **kw)
        return orig_rand(*size, **kw)
    torch.rand = small_rand
    try:
        model = gs.GaussianSplattingModel(gs.GaussianSplattingModelConfig(),
                                          scene_box=SceneBox(aabb=torch.tensor([[-1.0, -1, -1], [1, 1, 1]])),
                                          num_train_data=num_train_data)
    finally:
        torch.rand = orig_rand
    return model


# ---------------------------------------------------------------------------------------------------------------
def refine_fixture(gs):
    import make_refine_golden as mrg
    seed = mrg.SEED      # no decision of this seed lies within 1e-5 of a threshold (checked by the oracle's CPU test)
    P, M, st, z = mrg.inputs(seed)
    model = small_model(gs, 64)
    model.train()
    for k, attr in PARAM_OF.items():
        setattr(model, attr, torch.nn.Parameter(P[k].clone()))
    groups = model.get_gaussian_param_groups()
    optimizers = types.SimpleNamespace(optimizers={})
    for k, grp in GROUP_OF.items():
        (param,) = groups[grp]
        opt = torch.optim.Adam([param], lr=1e-3, eps=1e-15)
        opt.state[param] = dict(step=torch.tensor(7.0), exp_avg=M[k][0].clone(), exp_avg_sq=M[k][1].clone())
        optimizers.optimizers[grp] = opt
    # step 3500 of the default schedule: densify (3500 % 3000 = 500 > 4 + 100, < stop_split_at), screen-size split
    # (< stop_screen_size_at = 4000), cull incl. scale (> 3000) and screen size (< 4000): every branch of :402-464
    model.step = 3500
    model.last_size = (480, 640)
    model.xys_grad_norm = st["xys_grad_norm"].clone()
    model.vis_counts = st["vis_counts"].clone()
    model.max_2Dsize = st["max_2dsize"].clone()
    cfg = model.config
    rules = dict(max_dim=640.0, densify_grad_thresh=cfg.densify_grad_thresh, densify_size_thresh=cfg.densify_size_thresh,
                 split_screen_size=cfg.split_screen_size, cull_alpha_thresh=cfg.cull_alpha_thresh,
                 cull_scale_thresh=cfg.cull_scale_thresh, cull_screen_size=cfg.cull_screen_size, do_densify=1,
                 split_by_screen=1, do_cull=1, cull_by_scale=1, cull_by_screen=1)
    assert rules == mrg.RULES and cfg.n_split_samples == 2, "the reference's defaults moved"
    orig_randn = torch.randn
    drawn = []

    def recorded_randn(*size, **kw):
        shape = size[0] if len(size) == 1 and isinstance(size[0], (tuple, list, torch.Size)) else size
        assert len(shape) == 2 and shape[1] == 3, shape           # split_gaussians' one draw (:491)
        drawn.append(int(shape[0]))
        return z[:shape[0]].clone()
    n_cat = []
    orig_cull = model.cull_gaussians

    def counting_cull():
        n_cat.append(model.num_points)     # Gaussians after densification, before the cull
        return orig_cull()
    model.cull_gaussians = counting_cull
    torch.randn = recorded_randn
    try:
        model.refinement_after(optimizers, 3500)
    finally:
        torch.randn = orig_randn
    assert len(drawn) == 1 and len(n_cat) == 1 and model.xys_grad_norm is None
    out = dict(seed=np.array([seed]), z=z.numpy(), step=np.array([3500]), num_train_data=np.array([4]),
               last_size=np.array([480, 640]), n_samples_drawn=np.array(drawn), n_cat=np.array(n_cat),
               generated_by=np.array(["reference: GaussianSplattingModel.refinement_after"]))
    n_out = model.means.shape[0]
    out["n_out"] = np.array([n_out])
    for k, attr in PARAM_OF.items():
        out["in_" + k] = P[k].numpy()
        out["in_m0_" + k], out["in_m1_" + k] = M[k][0].numpy(), M[k][1].numpy()
        param = getattr(model, attr)
        assert param.shape[0] == n_out
        out["out_" + k] = param.detach().numpy().copy()
        opt = optimizers.optimizers[GROUP_OF[k]]
        assert opt.param_groups[0]["params"][0] is param, "the optimizer must hold the new parameter"
        state = opt.state[param]
        out["out_m0_" + k], out["out_m1_" + k] = state["exp_avg"].numpy().copy(), state["exp_avg_sq"].numpy().copy()
        assert out["out_m0_" + k].shape == out["out_" + k].shape
    for k, v in st.items():
        out["st_" + k] = v.numpy()
    return out


def schedule_table(gs):
    """Which rules refinement_after (:402-464) applies at which step, observed on the reference's own method: six probe
    Gaussians, each built so that exactly one rule decides its fate (copies of it afterwards tell whether the rule
    fired).  -1 = not observable at that step (the rule sits inside a branch that was not taken)."""
    import contextlib
    import io
    model = small_model(gs, 16)
    model.train()
    cfg = model.config
    n, K, D = 6, 25, 32
    hi, lo = 1.0, 0.0                                  # avg_grad_norm far above / below densify_grad_thresh

    def probe(step, num_train_data):
        P = dict(means=torch.arange(n * 3, dtype=torch.float32).reshape(n, 3),
                 log_scales=torch.log(torch.tensor([0.005, 0.7, 0.005, 0.005, 0.005, 0.005]))[:, None].repeat(1, 3),
                 quats=torch.tensor([[1.0, 0, 0, 0]]).repeat(n, 1),
                 opacity_logit=torch.tensor([3.0, 3.0, 3.0, -5.0, 3.0, 3.0])[:, None],
                 sh_coeffs=torch.zeros(n, K, 3), features=torch.zeros(n, D))
        P["features"][:, 0] = torch.arange(n)          # the identity each copy carries along
        for k, attr in PARAM_OF.items():
            setattr(model, attr, torch.nn.Parameter(P[k].clone()))
        groups = model.get_gaussian_param_groups()
        optimizers = types.SimpleNamespace(optimizers={})
        for k, grp in GROUP_OF.items():
            (param,) = groups[grp]
            opt = torch.optim.Adam([param], lr=1e-3, eps=1e-15)
            opt.state[param] = dict(step=torch.tensor(1.0), exp_avg=torch.ones_like(param), exp_avg_sq=torch.ones_like(param))
            optimizers.optimizers[grp] = opt
        model.step, model.num_train_data, model.last_size = step, num_train_data, (480, 640)
        grad = torch.tensor([hi, lo, lo, lo, hi, lo]) / (0.5 * 640)
        model.xys_grad_norm, model.vis_counts = grad.clone(), torch.ones(n)
        model.max_2Dsize = torch.tensor([0.1, 0.0, 0.2, 0.0, 0.0, 0.0])
        with contextlib.redirect_stdout(io.StringIO()):
            model.refinement_after(optimizers, step)
        if model.xys_grad_norm is not None:
            return [0, -1, -1, -1, -1, -1, -1]           # before warmup_length: nothing happens
        ids = model.feature.detach()[:, 0].round().long()
        copies = torch.bincount(ids, minlength=n).tolist()
        assert copies[5] == 1
        densify = int(copies[4] == 2)
        # probe 0 is small: split by its screen size it leaves two children AND, being small and high-gradient, a
        # duplicate (:419-421 decide the duplicates after the split) -- 4 copies; without the screen rule 2 (a duplicate)
        assert not densify or copies[0] in (2, 4)
        split_by_screen = int(copies[0] == 4) if densify else -1
        cull = int(copies[3] == 0)
        by_scale = int(copies[1] == 0) if cull else -1
        by_screen = int(copies[2] == 0) if by_scale == 1 else -1
        reset_value = torch.logit(torch.tensor(cfg.cull_alpha_thresh * 0.8)).item()
        reset = int(bool((model.opacities.detach() == reset_value).all()))
        opt = optimizers.optimizers["opacity"]
        moments_zeroed = bool((opt.state[opt.param_groups[0]["params"][0]]["exp_avg"] == 0).all())
        assert moments_zeroed == bool(reset)
        return [1, densify, split_by_screen, cull, by_scale, by_screen, reset]
    rows = []
    for num_train_data, last in ((4, 16000), (150, 7000), (350, 7000)):
        for step in range(0, last + 1, cfg.refine_every):           # the callback fires every refine_every iterations (:566-571)
            rows.append([step, num_train_data] + probe(step, num_train_data))
    return dict(schedule_columns=np.array(["step", "num_train_data", "active", "do_densify", "split_by_screen", "do_cull",
                                           "cull_by_scale", "cull_by_screen", "reset_opacity"]),
                schedule_rows=np.array(rows, dtype=np.int64))


# ---------------------------------------------------------------------------------------------------------------
def losses_fixture(gs):
    H, W, D = 40, 48, 32
    torch.manual_seed(11)
    model = small_model(gs, 300)
    model.train()
    model.step = 600                       # full resolution (:599-603); step % 10 == 0: the regularisers are on (:917)
    g = torch.Generator().manual_seed(12)
    with torch.no_grad():
        model.colors_all[:, 1:, :] = torch.randn(model.colors_all[:, 1:, :].shape, generator=g) * 0.1
        model.colors_all[::9, 1:, :] = 0.0                 # zero SH norm: the norm's subgradient
        model.scales.add_(torch.randn(model.scales.shape, generator=g) * 1.2)    # ratios on both sides of max_gauss_ratio
    leaves = dict(rgb=torch.rand((H, W, 3), generator=g), depth=torch.rand((H, W, 1), generator=g) * 4 + 0.2,
                  normal=torch.randn((H, W, 3), generator=g), feature=torch.randn((H, W, D), generator=g))
    for v in leaves.values():
        v.requires_grad_(True)
    outputs = {k: v * 1.0 for k, v in leaves.items()}      # non-leaves: get_loss_dict writes into outputs["rgb"] (:884)
    seg = torch.randint(0, 4, (H, W), generator=g)
    seg[:, :8] = 0
    depth = torch.rand((H, W, 1), generator=g) * 5 + 0.1
    depth[:5] = 0.01                                         # below the 0.05 validity threshold (:861)
    batch = dict(image=torch.rand((H, W, 3), generator=g), normal=torch.randn((H, W, 3), generator=g), depth=depth,
                 sam_mask=seg, valid_mask=torch.rand((H, W), generator=g) > 0.15,
                 # (multiples of 1/8: the 512-channel target compresses to a small fixture)
                 feature=torch.randint(-24, 25, (H, W, 512), generator=g).float() / 8)
    saved_batch = {k: v.clone() for k, v in batch.items()}
    rec = {}
    orig_pairs, orig_points = gs.sampling_pairs_in_mask, gs.sampling_in_mask

    def rec_pairs(mask, num):
        rec["pairs_mask"], rec["pairs_num"] = mask.clone(), num
        rec["pairs"] = orig_pairs(mask, num)
        return rec["pairs"]

    def rec_points(mask, num):
        rec["points_num"] = num
        rec["points"] = orig_points(mask, num)
        return rec["points"]
    gs.sampling_pairs_in_mask, gs.sampling_in_mask = rec_pairs, rec_points
    try:
        torch.manual_seed(13)
        losses = model.get_loss_dict(outputs, batch)
    finally:
        gs.sampling_pairs_in_mask, gs.sampling_in_mask = orig_pairs, orig_points
    names = ["main_loss", "feature_loss", "up_loss", "depth_loss", "normal_loss", "sh_reg", "scale_reg"]
    assert list(losses) == names
    weights = dict(main_loss=1.0, feature_loss=0.9, up_loss=0.5, depth_loss=0.7, normal_loss=1.3, sh_reg=1.1, scale_reg=0.8)
    sum(weights[k] * losses[k] for k in names).backward()
    out = dict(step=np.array([600]), ssim_lambda=np.array([model.config.ssim_lambda]),
               max_gauss_ratio=np.array([model.config.max_gauss_ratio]),
               loss_names=np.array(names), loss_values=np.array([float(losses[k]) for k in names], dtype=np.float64),
               loss_weights=np.array([weights[k] for k in names]),
               gt_mask=rec["pairs_mask"].numpy(), pairs_num=np.array([rec["pairs_num"]]), points_num=np.array([rec["points_num"]]),
               points=rec["points"].numpy(), n_segments=np.array([len(rec["pairs"])]))
    for i, (a, b) in enumerate(rec["pairs"]):
        out[f"pairs_{i}_a"], out[f"pairs_{i}_b"] = a.numpy(), b.numpy()
    for k, v in leaves.items():
        out["out_" + k] = v.detach().numpy()
        out["grad_" + k] = v.grad.numpy()
    for k, v in saved_batch.items():
        out["batch_" + k] = v.numpy()
    out["batch_feature_x8"] = (out.pop("batch_feature") * 8).astype(np.int8)       # exact: the values are multiples of 1/8
    out["colors_all"], out["grad_colors_all"] = model.colors_all.detach().numpy(), model.colors_all.grad.numpy()
    out["scales"], out["grad_scales"] = model.scales.detach().numpy(), model.scales.grad.numpy()
    for k, p in model.fea_up.named_parameters():
        out["mlp_" + k], out["mlp_grad_" + k] = p.detach().numpy(), p.grad.numpy()
    return out


def targets_fixture(gs):
    """The ground-truth side of get_loss_dict (:849-875) during the resolution warm-up (step 100: half size): the
    method's own local tensors, read from its frame when it returns."""
    H, W, D = 40, 48, 32
    torch.manual_seed(91)
    model = small_model(gs, 100)
    model.train()
    model.step = 100
    d = model._get_downscale_factor()
    assert d == 2
    g = torch.Generator().manual_seed(92)
    outputs = dict(rgb=torch.rand((H // d, W // d, 3), generator=g) * 1.0, depth=torch.rand((H // d, W // d, 1), generator=g) + 0.5,
                   normal=torch.randn((H // d, W // d, 3), generator=g), feature=torch.randn((H // d, W // d, D), generator=g))
    seg = torch.randint(0, 3, (H, W), generator=g)
    depth = torch.rand((H, W, 1), generator=g) * 5 + 0.1
    depth[:7, :] = 0.01
    depth[20:23, 10:30] = 0.04
    batch = dict(image=torch.rand((H, W, 3), generator=g), normal=torch.randn((H, W, 3), generator=g), depth=depth,
                 sam_mask=seg, valid_mask=torch.rand((H, W), generator=g) > 0.2,
                 feature=torch.randint(-1, 2, (H, W, 512), generator=g).float())
    saved = {k: v.clone() for k, v in batch.items()}
    captured = {}
    code = gs.GaussianSplattingModel.get_loss_dict.__code__

    def profiler(frame, event, arg):
        if event == "return" and frame.f_code is code:
            for k in ("gt_img", "gt_normal", "gt_depth", "depth_mask", "gt_mask", "valid_mask", "gt_fea"):
                captured[k] = frame.f_locals[k].detach().clone()
    sys.setprofile(profiler)
    try:
        model.get_loss_dict(outputs, batch)
    finally:
        sys.setprofile(None)
    assert captured["gt_img"].shape == (H // d, W // d, 3)
    out = dict(downscale=np.array([d]), step=np.array([100]))
    for k, v in saved.items():
        out["batch_" + k] = v.numpy()
    out["batch_feature"] = out["batch_feature"].astype(np.int8)              # values -1, 0, 1
    for k, v in captured.items():
        out["local_" + k] = v.numpy() if k != "gt_fea" else v.numpy().astype(np.int8)
    return out


# ---------------------------------------------------------------------------------------------------------------
def init_fixture(gs):
    torch.manual_seed(21)
    model = small_model(gs, 3000)
    out = dict(knn_means=model.means.detach().numpy().copy(), knn_log_scales=model.scales.detach().numpy().copy())
    # the up-projection MLP (:198-213) as the model builds it (:257), full-frame use as base_pipeline.py:408
    g = torch.Generator().manual_seed(22)
    x = torch.randn((9, 13, 32), generator=g)
    with torch.no_grad():
        for p in model.fea_up.parameters():
            p.mul_(2.0)
        out["mlp_x"], out["mlp_y"] = x.numpy(), model.fea_up(x).numpy()
        out["mlp_y64"] = model.fea_up.double()(x.double()).numpy()
        model.fea_up.float()
    for k, p in model.fea_up.named_parameters():
        out["mlp_" + k] = p.detach().numpy().copy()
    # projection_matrix(znear, zfar, fovx, fovy) (:87-105) and SH2RGB (:80-85)
    args = np.array([[0.001, 1000.0, 1.2, 0.9], [0.01, 100.0, 0.5, 0.7], [0.001, 1000.0, 2.0, 1.4]])
    out["proj_args"] = args
    out["proj_mats"] = np.stack([gs.projection_matrix(*[float(a) for a in row]).numpy() for row in args])
    c = torch.rand((50, 3), generator=g)
    out["sh_dc"], out["sh2rgb"] = c.numpy(), gs.SH2RGB(c).numpy()
    # the optimizer table of the method (configs/method_configs.py:611-664), read from the reference's own source by
    # its syntax tree (importing that module needs open3d, tyro and the whole model zoo), and the learning rates the
    # trainer's ExponentialDecayScheduler (engine/schedulers.py:109-140, imported and run) produces from it
    out.update(optimizer_table())
    out.update(camera_table(gs, model))
    out.update(sh_degree_table(gs, model))
    out.update(after_train_table(gs))
    out.update(schedule_table(gs))
    # the reference's own real spherical-harmonics basis (nerfstudio/utils/math.py:29-92), 5 levels = degree 4: the
    # same 25 functions in the same order as gsplat's, without the (-1)^|m| sign gsplat (like the 3DGS authors) carries
    from nerfstudio.utils.math import components_from_spherical_harmonics
    d = torch.nn.functional.normalize(torch.randn((96, 3), generator=torch.Generator().manual_seed(51)).double(), dim=-1)
    out["sh_dirs"] = d.float().numpy()
    out["sh_components"] = components_from_spherical_harmonics(5, d.float()).numpy()
    # the reference's own quaternion (w, x, y, z; any length) -> rotation matrix (nerfstudio/cameras/camera_utils.py:142-161)
    from nerfstudio.cameras.camera_utils import quaternion_matrix
    q = torch.randn((64, 4), generator=torch.Generator().manual_seed(52)) * 1.7
    out["quat_wxyz"] = q.numpy()
    out["quat_rotmat"] = np.stack([quaternion_matrix(row)[:3, :3] for row in q.numpy()])
    return out


def sh_degree_table(gs, model):
    """The number of SH bands get_outputs (:726-731) asks SphericalHarmonics.apply for, per training step: both
    operator classes intercepted (a stand-in projection, the SH call recorded and get_outputs abandoned there)."""
    from nerfstudio.cameras.cameras import Cameras, CameraType

    class Recorded(Exception):
        pass

    class Projection:
        @staticmethod
        def apply(means, *a):
            n = means.shape[0]
            return (means[:, :2] * 1.0, torch.ones(n), torch.ones(n, dtype=torch.int32), torch.ones(n, 3),
                    torch.ones(n, dtype=torch.int32), torch.ones(n, 6))

    class SH:
        @staticmethod
        def apply(n, viewdirs, coeffs):
            raise Recorded(n, tuple(coeffs.shape))
    steps = [0, 499, 500, 999, 1000, 1999, 2000, 2999, 3000, 3999, 4000, 4001, 9000, 29999]
    got = []
    orig = gs.ProjectGaussians, gs.SphericalHarmonics
    gs.ProjectGaussians, gs.SphericalHarmonics = Projection, SH
    model.train()
    try:
        for step in steps:
            model.step = step
            cam = Cameras(camera_to_worlds=torch.eye(4)[None, :3], fx=60.0, fy=60.0, cx=32.0, cy=24.0, width=64, height=48,
                          camera_type=CameraType.PERSPECTIVE)
            try:
                model.get_outputs(cam)
                raise AssertionError("SphericalHarmonics.apply was not reached")
            except Recorded as r:
                assert r.args[1][1:] == (25, 3)
                got.append(int(r.args[0]))
    finally:
        gs.ProjectGaussians, gs.SphericalHarmonics = orig
    return dict(sh_steps=np.array(steps), sh_degrees_to_use=np.array(got))


def after_train_table(gs):
    """GaussianSplattingModel.after_train (:373-393), three training steps: the densification statistics it keeps from
    xys.grad and radii."""
    n, H, W = 701, 480, 640
    model = small_model(gs, n)
    model.train()
    model.last_size = (H, W)
    g = torch.Generator().manual_seed(41)
    out = dict(stats_size=np.array([H, W]))
    for it in range(3):
        radii = torch.randint(-1, 40, (n,), generator=g, dtype=torch.int32).clamp(min=0)
        radii[it::3] = 0
        v_xy = torch.randn((n, 2), generator=g) * 1e-4
        v_xy[radii == 0] = 0                      # invisible Gaussians receive no gradient
        model.radii = radii
        model.xys = torch.zeros((n, 2), requires_grad=True)
        model.xys.grad = v_xy.clone()
        model.after_train(it)
        out[f"stats_radii_{it}"], out[f"stats_vxy_{it}"] = radii.numpy(), v_xy.numpy()
        out[f"stats_norm_{it}"], out[f"stats_count_{it}"] = model.xys_grad_norm.numpy().copy(), model.vis_counts.numpy().copy()
        out[f"stats_max2d_{it}"] = model.max_2Dsize.numpy().copy()
    return out


def camera_table(gs, model):
    """What get_outputs (:624-713) hands to ProjectGaussians.apply for a few nerfstudio cameras: the call is
    intercepted at the gsplat boundary (arguments recorded, get_outputs abandoned there)."""
    from nerfstudio.cameras.cameras import Cameras, CameraType

    class Recorded(Exception):
        pass

    class Recorder:
        @staticmethod
        def apply(means, scales, glob_scale, quats, viewmat, fullmat, fx, fy, cx, cy, H, W, tile_bounds):
            raise Recorded(dict(viewmat=viewmat.detach().clone(), fullmat=fullmat.detach().clone(),
                                intr=[fx, fy, cx, cy], size=[H, W], tile_bounds=list(tile_bounds), glob_scale=glob_scale))
    g = torch.Generator().manual_seed(31)
    model.train()
    model.step = 600
    rows = dict(c2w=[], intr_in=[], size_in=[], viewmat=[], fullmat=[], intr=[], size=[], tile_bounds=[], step=[], downscale=[])
    orig = gs.ProjectGaussians
    gs.ProjectGaussians = Recorder
    try:
        # (step 100 lies in the resolution warm-up: the model renders at half size, :599-603, :655-656)
        for (W, H), step in (((640, 480), 600), ((1280, 720), 600), ((333, 217), 600), ((1920, 1080), 600), ((640, 480), 100),
                             ((333, 217), 249), ((333, 217), 250)):
            model.step = step
            q = torch.nn.functional.normalize(torch.randn(4, generator=g), dim=0)
            R = gs.quat_to_rotmat(q[None])[0]
            t = torch.randn(3, 1, generator=g) * 3
            c2w = torch.cat([R, t], dim=1)[None]
            fx, fy = 0.6 * W * (1 + 0.1 * float(torch.rand(1, generator=g))), 0.6 * W * (1 + 0.1 * float(torch.rand(1, generator=g)))
            cx, cy = W / 2 + float(torch.randn(1, generator=g)) * 5, H / 2 + float(torch.randn(1, generator=g)) * 5
            cam = Cameras(camera_to_worlds=c2w, fx=fx, fy=fy, cx=cx, cy=cy, width=W, height=H, camera_type=CameraType.PERSPECTIVE)
            try:
                model.get_outputs(cam)
                raise AssertionError("ProjectGaussians.apply was not reached")
            except Recorded as r:
                rec = r.args[0]
            d = model._get_downscale_factor()
            assert rec["glob_scale"] == 1 and model.last_size == tuple(rec["size"]) and d == (2 if step < 250 else 1)
            rows["step"].append(step); rows["downscale"].append(d)
            rows["c2w"].append(c2w[0].numpy()); rows["intr_in"].append([fx, fy, cx, cy]); rows["size_in"].append([W, H])
            rows["viewmat"].append(rec["viewmat"].numpy()); rows["fullmat"].append(rec["fullmat"].numpy())
            rows["intr"].append(rec["intr"]); rows["size"].append(rec["size"]); rows["tile_bounds"].append(rec["tile_bounds"])
    finally:
        gs.ProjectGaussians = orig
    return {"cam_" + k: np.array(v) for k, v in rows.items()}


def optimizer_table():
    import ast
    from nerfstudio.engine.schedulers import ExponentialDecayScheduler, ExponentialDecaySchedulerConfig
    tree = ast.parse(open(os.path.join(REF, "nerfstudio/configs/method_configs.py")).read())
    call = None
    for node in ast.walk(tree):
        if isinstance(node, ast.Assign) and isinstance(node.targets[0], ast.Subscript) \
                and getattr(node.targets[0].value, "id", "") == "method_configs" \
                and ast.literal_eval(node.targets[0].slice) == "gaussian-splatting":
            call = node.value
    assert call is not None, "method_configs['gaussian-splatting'] not found"
    kw = {k.arg: k.value for k in call.keywords}
    accumulation = ast.literal_eval(kw["gradient_accumulation_steps"])
    groups, rows = [], []
    for key, val in zip(kw["optimizers"].keys, kw["optimizers"].values):
        entry = {ast.literal_eval(k): v for k, v in zip(val.keys, val.values)}
        opt = {k.arg: ast.literal_eval(k.value) for k in entry["optimizer"].keywords}
        assert entry["optimizer"].func.id == "AdamOptimizerConfig"
        sch = entry["scheduler"]
        if isinstance(sch, ast.Constant) and sch.value is None:
            lr_final, max_steps = float("nan"), -1
        else:
            assert sch.func.id == "ExponentialDecaySchedulerConfig"
            sk = {k.arg: ast.literal_eval(k.value) for k in sch.keywords}
            lr_final, max_steps = sk["lr_final"], sk["max_steps"]
        groups.append(ast.literal_eval(key))
        rows.append([opt["lr"], opt["eps"], lr_final, max_steps])
    probe = np.array([0, 1, 2, 9, 10, 100, 1000, 2999, 15000, 29999, 30000, 35000])
    lrs = np.full((len(groups), len(probe)), np.nan)
    for i, (lr, eps, lr_final, max_steps) in enumerate(rows):
        if max_steps < 0:
            continue
        opt = torch.optim.Adam([torch.nn.Parameter(torch.zeros(1))], lr=lr, eps=eps)
        sch = ExponentialDecayScheduler(ExponentialDecaySchedulerConfig(lr_final=lr_final, max_steps=int(max_steps))).get_scheduler(opt, lr)
        lrs[i] = [lr * float(sch.lr_lambdas[0](int(step))) for step in probe]
        # LambdaLR itself, stepping: lr of the first iterations as the trainer sees them
        seen = []
        for _ in range(3):
            seen.append(opt.param_groups[0]["lr"])
            opt.step()
            sch.step()
        assert np.allclose(seen, lrs[i, :3], rtol=1e-12)
    return dict(opt_groups=np.array(groups), opt_lr_eps_final_maxsteps=np.array(rows, dtype=np.float64),
                opt_probe_steps=probe, opt_probe_lrs=lrs,
                accumulation_groups=np.array(list(accumulation)), accumulation_steps=np.array(list(accumulation.values())))


# ---------------------------------------------------------------------------------------------------------------
def trainer_fixture(gs):
    """Trainer.train_iteration (engine/trainer.py:458-499), the reference's own function, run for 25 iterations on a
    stand-in `self` that holds the REAL Optimizers (built from the method's optimizer table), the method's gradient
    accumulation table and a pipeline whose loss is linear in the parameters -- so that every iteration's gradients
    are the stored arrays.  Stored: initial parameters, gradients per iteration, parameters after every iteration."""
    from collections import defaultdict
    from reference_model_driver import import_with_stubs
    from nerfstudio.engine.optimizers import AdamOptimizerConfig, Optimizers
    from nerfstudio.engine.schedulers import ExponentialDecaySchedulerConfig
    tr, _ = import_with_stubs("nerfstudio.engine.trainer")
    table = optimizer_table()
    cfg = {}
    for group, (lr, eps, lr_final, max_steps) in zip(table["opt_groups"].tolist(), table["opt_lr_eps_final_maxsteps"]):
        cfg[group] = {"optimizer": AdamOptimizerConfig(lr=float(lr), eps=float(eps)),
                      "scheduler": ExponentialDecaySchedulerConfig(lr_final=float(lr_final), max_steps=int(max_steps))
                      if max_steps > 0 else None}
    torch.manual_seed(61)
    n, T = 24, 25
    model = small_model(gs, n)
    g = torch.Generator().manual_seed(62)
    with torch.no_grad():
        model.colors_all[:, 1:, :] = torch.randn(model.colors_all[:, 1:, :].shape, generator=g) * 0.1
    groups = model.get_param_groups()
    accumulation = defaultdict(lambda: 1)
    accumulation.update(dict(zip(table["accumulation_groups"].tolist(), table["accumulation_steps"].tolist())))
    grads = {k: torch.randn((T,) + tuple(getattr(model, attr).shape), generator=g) * 1e-2 for k, attr in PARAM_OF.items()}

    class Pipeline:
        model_ = model

        def get_train_loss_dict(self, step):
            loss = {k: (getattr(model, attr) * grads[k][step]).sum() for k, attr in PARAM_OF.items()}
            return None, loss, {}
    fake = types.SimpleNamespace(optimizers=Optimizers(cfg, groups), gradient_accumulation_steps=accumulation, device="cpu",
                                 mixed_precision=False, pipeline=Pipeline(), config=types.SimpleNamespace(log_gradients=False),
                                 grad_scaler=torch.cuda.amp.GradScaler(enabled=False))
    out = dict(steps=np.array([T]))
    for k, attr in PARAM_OF.items():
        out["init_" + k] = getattr(model, attr).detach().numpy().copy()
        out["grads_" + k] = grads[k].numpy()
    traj = {k: [] for k in PARAM_OF}
    lrs = {k: [] for k in GROUP_OF}
    for step in range(T):
        for k, grp in GROUP_OF.items():
            lrs[k].append(float(fake.optimizers.optimizers[grp].param_groups[0]["lr"]))      # the rate this iteration uses
        tr.Trainer.train_iteration(fake, step)
        for k, attr in PARAM_OF.items():
            traj[k].append(getattr(model, attr).detach().numpy().copy())
    for k in PARAM_OF:
        out["after_" + k] = np.stack(traj[k])
        out["lr_" + k] = np.array(lrs[k], dtype=np.float64)
        opt = fake.optimizers.optimizers[GROUP_OF[k]]
        out["final_step_" + k] = np.array([float(opt.state[opt.param_groups[0]["params"][0]]["step"])])
    return out


def ply_fixture(gs):
    """ExportGaussianSplat.main (nerfstudio/scripts/exporter.py:482-530), the reference's own exporter, run on a small
    model: open3d is absent, so the attribute map the exporter builds and hands to o3d.t.geometry.PointCloud is
    recorded instead of written (names in the order the exporter inserts them, arrays, dtypes).  The order in which
    open3d then writes the properties into the file is open3d's and not visible here."""
    from reference_model_driver import import_with_stubs
    ex, _ = import_with_stubs("nerfstudio.scripts.exporter")
    torch.manual_seed(71)
    model = small_model(gs, 40)
    g = torch.Generator().manual_seed(72)
    with torch.no_grad():
        model.colors_all.copy_(torch.randn(model.colors_all.shape, generator=g) * 3.0)   # colours beyond [0,1] too: the uint8 cast
        model.opacities.copy_(torch.randn(model.opacities.shape, generator=g))
    recorded = {}

    class O3D:
        class core:
            float32 = "float32"

            @staticmethod
            def Tensor(a, dtype):
                return np.asarray(a, dtype=np.float32)

        class t:
            class geometry:
                @staticmethod
                def PointCloud(m):
                    recorded.update(m)
                    return m

            class io:
                @staticmethod
                def write_point_cloud(path, pcd):
                    recorded["__path__"] = path
    orig_o3d, orig_setup = ex.o3d, ex.eval_setup
    ex.o3d = O3D
    ex.eval_setup = lambda cfg: (None, types.SimpleNamespace(model=model), None, None)
    try:
        import pathlib
        import tempfile
        with tempfile.TemporaryDirectory() as d:
            ex.ExportGaussianSplat.main(types.SimpleNamespace(output_dir=pathlib.Path(d), load_config=None))
    finally:
        ex.o3d, ex.eval_setup = orig_o3d, orig_setup
    assert recorded.pop("__path__").endswith("point_cloud.ply")
    out = dict(attribute_names=np.array(list(recorded)))
    for k, v in recorded.items():
        out["attr_" + k] = np.asarray(v)
    for k, attr in PARAM_OF.items():
        out["param_" + k] = getattr(model, attr).detach().numpy().copy()
    return out


def outputs_fixture(gs):
    """GaussianSplattingModel.get_outputs (:624-802) + backward, the reference's own model code, at SH degree 4 on a
    48x64 view.  gsplat itself is absent: the four operator classes the model calls are this repository's autograd
    Functions with the CPU oracle underneath (tests/reference_model_driver.install_oracle_backend), so the fixture pins
    everything AROUND the operators -- activations, view directions, SH clamp, the smallest-axis normal, backgrounds,
    channel order, which gradients flow where -- to the reference's code, and the operators to the oracle.  Also stored:
    the pixels the oracle flags as decided within 2e-5 of a blending threshold."""
    from nerfstudio.cameras.cameras import Cameras, CameraType
    from reference_model_driver import install_oracle_backend
    from oracle import c_oracle
    from gaussiangrasper_b200 import scenes
    install_oracle_backend()
    torch.manual_seed(81)
    n, H, W = 800, 48, 64
    model = small_model(gs, n)
    model.train()
    g = torch.Generator().manual_seed(82)
    with torch.no_grad():
        model.means.mul_(0.3)
        model.scales.add_(-2.3 + 0.4 * torch.randn(model.scales.shape, generator=g))
        model.quats.mul_(1.0 + torch.rand((n, 1), generator=g))                  # un-normalised, as during training
        model.opacities.copy_(torch.randn(model.opacities.shape, generator=g) * 1.5)
        model.colors_all[:, 1:, :] = torch.randn(model.colors_all[:, 1:, :].shape, generator=g) * 0.15
    model.step = 4600                        # full resolution (:599-603), SH degree min(4600 // 1000, 4) = 4 (:729)
    c2w = torch.tensor([[[0.96, 0.0, 0.28, 1.1], [0.0, 1.0, 0.0, 0.1], [-0.28, 0.0, 0.96, 3.8]]])
    fx, fy, cx, cy = 61.5, 60.25, 32.75, 23.5
    cam = Cameras(camera_to_worlds=c2w, fx=fx, fy=fy, cx=cx, cy=cy, width=W, height=H, camera_type=CameraType.PERSPECTIVE)
    out = model.get_outputs(cam)
    v = dict(rgb=torch.randn((H, W, 3), generator=g), feature=torch.randn((H, W, 32), generator=g),
             depth=torch.randn((H, W, 1), generator=g) * 0.1, normal=torch.randn((H, W, 3), generator=g))
    sum((out[k] * v[k]).sum() for k in v).backward()
    res = dict(c2w=c2w[0].numpy(), intrinsics=np.array([fx, fy, cx, cy]), size=np.array([H, W]), step=np.array([4600]),
               radii=model.radii.numpy().copy(), xys=model.xys.detach().numpy().copy(), grad_xys=model.xys.grad.numpy().copy())
    for k, attr in PARAM_OF.items():
        res["param_" + k] = getattr(model, attr).detach().numpy().copy()
        res["grad_" + k] = getattr(model, attr).grad.numpy().copy()
    for k in v:
        res["out_" + k], res["v_" + k] = out[k].detach().numpy().copy(), v[k].numpy()
    # fragile pixels: the oracle's own flag on the same projection
    c = scenes.camera_from_c2w(c2w[0], fx, fy, cx, cy, W, H)
    q = model.quats.detach() / model.quats.detach().norm(dim=-1, keepdim=True)
    xys, depths, radii, conics, nth, _ = c_oracle.project_fwd(model.means.detach().numpy(), model.scales.detach().exp().numpy(), 1.0,
                                                              q.numpy(), c.viewmat[:3].numpy(), c.fullmat.numpy(), c.fx, c.fy,
                                                              c.cx, c.cy, H, W, c.tile_bounds)
    assert np.array_equal(radii, res["radii"]) and np.array_equal(xys, res["xys"])
    _, _, _, _, ids_s, ranges = c_oracle.bin_and_sort(xys, depths, radii, nth, c.tile_bounds)
    ref = c_oracle.blend_fwd(H, W, c.tile_bounds, ids_s, ranges, xys, conics, torch.sigmoid(model.opacities.detach()).reshape(-1).numpy(),
                             np.zeros((n, 3), np.float32), np.zeros(3, np.float32), eps=2e-5)
    res["fragile"] = ref[3]
    res["visible"] = np.array([int((radii > 0).sum())])
    res["intersections"] = np.array([len(ids_s)])
    return res


def main():
    assert os.path.exists(os.path.join(REF, "nerfstudio/models/gaussian_splatting.py")), "needs /root/reference"
    out_dir = sys.argv[sys.argv.index("--out") + 1] if "--out" in sys.argv else HERE
    gs = import_reference()
    for name, fn in (("refine_small", refine_fixture), ("ref_losses_small", losses_fixture), ("ref_init_small", init_fixture),
                     ("ref_trainer_small", trainer_fixture), ("ref_ply_small", ply_fixture), ("ref_targets_small", targets_fixture),
                     ("ref_outputs_small", outputs_fixture)):      # (last: it swaps the oracle in underneath the operators)
        path = os.path.join(out_dir, name + ".npz")
        new = {k: np.asarray(v) for k, v in fn(gs).items()}
        if os.path.exists(path):            # an .npz differs byte-wise from run to run (zip time stamps): keep equal content
            old = np.load(path)
            if set(old.files) == set(new) and all(old[k].dtype == new[k].dtype and old[k].shape == new[k].shape and
                                                  np.array_equal(old[k], new[k], equal_nan=new[k].dtype.kind in "fc")
                                                  for k in new):
                print(path, "unchanged")
                continue
        np.savez_compressed(path, **new)
        print(path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
