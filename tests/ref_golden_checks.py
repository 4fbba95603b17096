"""Checks against the fixtures the REFERENCE ITSELF produced (tests/golden/make_reference_golden.py: the unmodified
nerfstudio/models/gaussian_splatting.py run on seeded inputs): get_loss_dict with its gradients, the k-NN scale
initialisation, the up-projection MLP, after_train's densification statistics, nerfstudio's own SH basis and
quaternion convention, and the parameter trajectory of the reference's own Trainer.train_iteration.  Each check is written once, against the signatures of the product's loss
functions (gaussiangrasper_b200.losses / .training):

  * tests/test_gpu_zz_reference_golden.py passes the product (CUDA kernels through the C ABI);
  * tests/test_reference_golden_cpu.py passes `OracleBackend`, the same signatures served by oracle/loss_oracle.py
    (fp64 autograd) -- that pins the oracle, which every other GPU loss test uses as its checker, to the reference's
    own numbers, and exercises every line of these checks on the CPU.

TEST INFRASTRUCTURE ONLY (imports oracle/)."""
import os

import numpy as np
import torch
import torch.nn.functional as F

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CP = 40                      # channels of the blended image: rgb 0..2, depth 3, normal 4..6, features 7..38, 1 of padding
D = 32


def load(name):
    return dict(np.load(os.path.join(GOLDEN, name + ".npz")))


# ---------------------------------------------------------------------------------------------------------------
# the product's signatures on the CPU oracle
# ---------------------------------------------------------------------------------------------------------------
class OracleBackend:
    """gaussiangrasper_b200.losses / .training signatures served by oracle/loss_oracle.py in fp64 (gradients by
    autograd).  Only what the checks below call."""
    device = torch.device("cpu")

    def __init__(self):
        from oracle import loss_oracle
        self.lo = loss_oracle

    def make_mlp(self, state):
        mlp = self.lo.MLP(in_dim=D, out_dim=512, hidden_list=[128]).double()
        mlp.load_state_dict({k: v.double() for k, v in state.items()})
        return mlp

    def _with_grad(self, image, fn):
        x = image.double().requires_grad_(True)
        loss = fn(x)
        loss.sum().backward()
        return loss.detach().float(), x.grad.float()

    def geom_loss(self, image, gt_depth, gt_normal, depth_mask, w_depth=1.0, w_normal=1.0, grad=None):
        def fn(x):
            d, n = self.lo.geom_losses(x[..., 4:7], x[..., 3:4], gt_normal.permute(2, 0, 1).double(), gt_depth[None].double(),
                                       depth_mask[None])
            return torch.stack([w_depth * d, w_normal * n])
        loss, g = self._with_grad(image, fn)
        return loss, (g if grad is None else grad.add_(g))

    def contrastive_feature_loss(self, image, pairs, grad, feature_channel0=7, feature_dim=None, weight=1.0):
        loss, g = self._with_grad(image, lambda x: weight * self.lo.feature_loss(x[..., 7:7 + feature_dim], pairs).reshape(1))
        grad.add_(g)
        return loss

    def up_loss(self, image, points, gt_features, mlp, grad, feature_channel0=7, feature_dim=None, weight=1.0):
        fn = lambda x: weight * self.lo.up_loss(x[..., 7:7 + feature_dim], points, gt_features.permute(2, 0, 1).double(), mlp).reshape(1)
        loss, g = self._with_grad(image, fn)
        grad.add_(g)
        return loss

    def param_regs(self, sh_coeffs, log_scales, max_gauss_ratio=10.0, w_sh=1.0, w_scale=1.0, v_sh_coeffs=None,
                   v_log_scales=None):
        a, b = sh_coeffs.double().requires_grad_(True), log_scales.double().requires_grad_(True)
        r_sh, r_sc = self.lo.regs(a, b, max_gauss_ratio)
        (w_sh * r_sh + w_scale * r_sc).backward()
        v_sh_coeffs.add_(a.grad.float())
        v_log_scales.add_(b.grad.float())
        return torch.stack([w_sh * r_sh, w_scale * r_sc]).detach().float()

    def pixel_loss(self, pred, target, kind="l1", mask=None, weight=1.0, mean_over="valid"):
        assert kind == "l1" and mean_over == "valid"
        x = pred.double().requires_grad_(True)
        loss = weight * (target.double()[mask] - x[mask]).abs().mean()
        loss.backward()
        return loss.detach().float().reshape(1), x.grad.float()

    def ssim_loss(self, pred, target, channels=3, weight=1.0, grad=None, loss=None):
        x = pred.double().requires_grad_(True)
        val = weight * (1 - self.lo.ssim(target.double().permute(2, 0, 1)[None], x.permute(2, 0, 1)[None]))
        val.backward()
        return val.detach().float().reshape(1), x.grad.float()

    def knn_scale_init(self, means):
        d = torch.cdist(means.double(), means.double())
        d.fill_diagonal_(float("inf"))
        dist = torch.sort(d, dim=1)[0][:, :3]
        return torch.log(dist.mean(dim=-1, keepdim=True).repeat(1, 3)).float(), dist.float()

    def make_stats(self, n):
        """training.DensifyStats' update / attributes: after_train (:373-393) restated."""
        class Stats:
            xys_grad_norm = vis_counts = max_2Dsize = None

            def update(st, v_geo, radii, H, W):
                visible = radii > 0
                grads = v_geo[:, :2].norm(dim=-1)
                if st.xys_grad_norm is None:
                    st.xys_grad_norm, st.vis_counts, st.max_2Dsize = grads.clone(), torch.ones_like(grads), torch.zeros_like(grads)
                else:
                    st.vis_counts[visible] += 1
                    st.xys_grad_norm[visible] += grads[visible]
                st.max_2Dsize[visible] = torch.maximum(st.max_2Dsize[visible], radii[visible].float() / float(max(H, W)))
        return Stats()

    def make_adam(self, params):
        """training.FusedAdam.reference(params): the product's own host logic (FusedAdam.plan, REFERENCE_* tables,
        exponential_decay_lr) around a plain-torch restatement of what gg_adam_step does per mode."""
        from gaussiangrasper_b200 import training

        class CpuAdam:
            def __init__(self):
                self.names = [k for k in params]
                self.bucket = self                      # FusedAdam.plan reads self.bucket.names; pack() below
                self.accumulation = dict(training.REFERENCE_ACCUMULATION)
                self.schedules = dict(training.REFERENCE_SCHEDULES)
                self.lr_init = dict(training.REFERENCE_LRS)
                self.lrs = dict(self.lr_init)
                self.steps = {k: 0 for k in params}
                self.m = {k: torch.zeros_like(v, dtype=torch.float64) for k, v in params.items()}
                self.v = {k: torch.zeros_like(v, dtype=torch.float64) for k, v in params.items()}
                self.acc = {k: torch.zeros_like(v, dtype=torch.float64) for k, v in params.items()}

            def pack(self, grads):
                self.grads = grads

            def train_step(self, it):
                for k, (lr_final, max_steps) in self.schedules.items():
                    self.lrs[k] = training.exponential_decay_lr(self.lr_init[k], lr_final, max_steps, it)
                modes = training.FusedAdam.plan(self, it)
                for k, mode in modes.items():
                    g = self.grads[k].double()
                    if mode == "acc_first":
                        self.acc[k] = g.clone()
                    elif mode in ("acc", "acc_step"):
                        self.acc[k] += g
                    if mode in ("step", "acc_step"):
                        g = self.acc[k] if mode == "acc_step" else g
                        self.steps[k] += 1
                        t = self.steps[k]
                        self.m[k] = 0.9 * self.m[k] + 0.1 * g
                        self.v[k] = 0.999 * self.v[k] + 0.001 * g * g
                        upd = self.lrs[k] * (self.m[k] / (1 - 0.9 ** t)) / ((self.v[k] / (1 - 0.999 ** t)).sqrt() + 1e-15)
                        params[k].copy_((params[k].double() - upd).float())
                return modes
        return CpuAdam()

    def up_project(self, features, mlp):
        with torch.no_grad():
            return mlp(features.double()).float()


# ---------------------------------------------------------------------------------------------------------------
def ground_truth(fix):
    """The ground-truth side of get_loss_dict (:850-875) at full resolution, where its F.interpolate calls (to the
    image's own size) are identities.  Returns CPU tensors."""
    t = torch.from_numpy
    valid = t(fix["batch_valid_mask"])
    gt_depth = t(fix["batch_depth"])[..., 0]
    gt_mask = t(fix["batch_sam_mask"]).float()
    gt_mask[~valid] = -1.0                                                   # :873
    assert torch.equal(gt_mask, t(fix["gt_mask"])), "the samplers of the reference saw another segment mask"
    return dict(image=t(fix["batch_image"]), normal=F.normalize(t(fix["batch_normal"]), dim=-1),      # :857-859
                depth=gt_depth, depth_mask=(gt_depth > 0.05) & valid,                                    # :861, :870-871
                valid=valid, feature=t(fix["batch_feature_x8"]).float() / 8)


def blended_image(fix):
    H, W = fix["out_depth"].shape[:2]
    img = torch.zeros((H, W, CP))
    img[..., 0:3], img[..., 3:4] = torch.from_numpy(fix["out_rgb"]), torch.from_numpy(fix["out_depth"])
    img[..., 4:7], img[..., 7:7 + D] = torch.from_numpy(fix["out_normal"]), torch.from_numpy(fix["out_feature"])
    return img


def _close(got, want, rel, what):
    got, want = got.detach().cpu().double(), torch.as_tensor(want).double()
    assert got.shape == want.shape, (what, tuple(got.shape), tuple(want.shape))
    scale = float(want.abs().max())
    err = float((got - want).abs().max())
    assert err <= rel * scale + 1e-9, f"{what}: max error {err:.3e} against a largest entry of {scale:.3e}"


def check_losses(backend, rel_value=5e-5, rel_grad=1e-4):
    """Every term of the reference's get_loss_dict and the gradient of their weighted sum (the weights of the fixture)
    w.r.t. the blended outputs, the Gaussian parameters and the up-projection's weights."""
    fix = load("ref_losses_small")
    dev = backend.device
    gt = {k: v.to(dev) for k, v in ground_truth(fix).items()}
    want = dict(zip(fix["loss_names"].tolist(), fix["loss_values"].tolist()))
    w = dict(zip(fix["loss_names"].tolist(), fix["loss_weights"].tolist()))
    lam = float(fix["ssim_lambda"][0])
    img = blended_image(fix).to(dev)
    H, W, _ = img.shape
    got = {}

    # depth_loss, normal_loss (:876-880)
    loss, grad = backend.geom_loss(img, gt["depth"], gt["normal"], gt["depth_mask"], w_depth=w["depth_loss"],
                                   w_normal=w["normal_loss"])
    got["depth_loss"], got["normal_loss"] = loss[0].item() / w["depth_loss"], loss[1].item() / w["normal_loss"]
    assert grad.shape == img.shape

    # feature_loss over the pixel pairs the reference's sampler drew (:905-912), up_loss over its points (:913-914)
    pairs = [[torch.from_numpy(fix[f"pairs_{i}_a"]).to(dev), torch.from_numpy(fix[f"pairs_{i}_b"]).to(dev)]
             for i in range(int(fix["n_segments"][0]))]
    points = torch.from_numpy(fix["points"]).to(dev)
    mlp = backend.make_mlp({k[len("mlp_"):]: torch.from_numpy(v) for k, v in fix.items()
                            if k.startswith("mlp_layers")})
    got["feature_loss"] = backend.contrastive_feature_loss(img, pairs, grad, feature_dim=D, weight=w["feature_loss"]).item() \
        / w["feature_loss"]
    got["up_loss"] = backend.up_loss(img, points, gt["feature"], mlp, grad, feature_dim=D, weight=w["up_loss"]).item() / w["up_loss"]

    # main_loss = (1 - lambda) L1 over the valid pixels + lambda (1 - SSIM) of the images with the invalid pixels
    # zeroed on both sides (:882-885, :931); the in-place zeroing passes no gradient to the pixels it overwrites
    rgb = img[..., 0:3].contiguous()
    l1, g_l1 = backend.pixel_loss(rgb, gt["image"], "l1", gt["valid"], weight=(1 - lam) * w["main_loss"])
    rgb_z, gt_z = rgb.clone(), gt["image"].clone()
    rgb_z[~gt["valid"]] = 0.0
    gt_z[~gt["valid"]] = 0.0
    ls, g_ssim = backend.ssim_loss(rgb_z, gt_z, weight=lam * w["main_loss"])
    g_ssim = g_ssim.reshape(H, W, 3).clone()
    g_ssim[~gt["valid"]] = 0.0
    got["main_loss"] = (l1.item() + ls.item()) / w["main_loss"]
    grad[..., 0:3] += g_l1.reshape(H, W, 3) + g_ssim

    # sh_reg, scale_reg (:917-925)
    sh, scales = torch.from_numpy(fix["colors_all"]).to(dev), torch.from_numpy(fix["scales"]).to(dev)
    v_sh, v_sc = torch.zeros_like(sh), torch.zeros_like(scales)
    regs = backend.param_regs(sh, scales, float(fix["max_gauss_ratio"][0]), w_sh=w["sh_reg"], w_scale=w["scale_reg"],
                              v_sh_coeffs=v_sh, v_log_scales=v_sc)
    got["sh_reg"], got["scale_reg"] = regs[0].item() / w["sh_reg"], regs[1].item() / w["scale_reg"]

    for k, v in want.items():
        assert abs(got[k] - v) <= rel_value * abs(v), f"{k}: {got[k]!r} against the reference's {v!r}"
    _close(grad[..., 0:3], fix["grad_rgb"], rel_grad, "d/d rgb")
    _close(grad[..., 3:4], fix["grad_depth"], rel_grad, "d/d depth")
    _close(grad[..., 4:7], fix["grad_normal"], rel_grad, "d/d normal")
    _close(grad[..., 7:7 + D], fix["grad_feature"], rel_grad, "d/d feature")
    assert float(grad[..., 7 + D:].abs().max()) == 0.0
    _close(v_sh, fix["grad_colors_all"], rel_grad, "d/d colors_all")
    _close(v_sc, fix["grad_scales"], rel_grad, "d/d scales")
    for k, p in mlp.named_parameters():
        _close(p.grad, fix["mlp_grad_" + k], rel_grad, "d/d fea_up." + k)
    return got


def check_init(backend):
    """populate_modules' k-NN scale initialisation (sklearn in the reference) and the up-projection MLP's forward."""
    fix = load("ref_init_small")
    dev = backend.device
    ls, dist = backend.knn_scale_init(torch.from_numpy(fix["knn_means"]).to(dev))
    assert ls.shape == fix["knn_log_scales"].shape
    err = float((ls.cpu() - torch.from_numpy(fix["knn_log_scales"])).abs().max())
    assert err <= 2e-5, f"k-NN log scales: {err:.3e}"
    assert torch.allclose(dist.cpu().mean(dim=1).log(), torch.from_numpy(fix["knn_log_scales"])[:, 0], atol=2e-5)
    mlp = backend.make_mlp({k[len("mlp_"):]: torch.from_numpy(v) for k, v in fix.items() if k.startswith("mlp_layers")})
    y = backend.up_project(torch.from_numpy(fix["mlp_x"]).to(dev), mlp)
    assert tuple(y.shape) == fix["mlp_y"].shape
    scale = float(np.abs(fix["mlp_y64"]).max())
    assert float((y.cpu().double() - torch.from_numpy(fix["mlp_y64"])).abs().max()) <= 1e-5 * scale    # the reference's weights in fp64
    assert float((y.cpu() - torch.from_numpy(fix["mlp_y"])).abs().max()) <= 2e-5 * scale                # its own fp32 output


def check_after_train(backend):
    """The densification statistics of the reference's after_train (:373-393) over three steps."""
    fix = load("ref_init_small")
    dev = backend.device
    H, W = (int(v) for v in fix["stats_size"])
    n = fix["stats_radii_0"].shape[0]
    stats = backend.make_stats(n)
    for it in range(3):
        v_geo = torch.zeros((n, 8))
        v_geo[:, :2] = torch.from_numpy(fix[f"stats_vxy_{it}"])
        v_geo[:, 2:] = 7.0                                    # the other columns of the packed gradient are not read
        stats.update(v_geo.to(dev), torch.from_numpy(fix[f"stats_radii_{it}"]).to(dev), H, W)
        assert torch.allclose(stats.xys_grad_norm.cpu(), torch.from_numpy(fix[f"stats_norm_{it}"]), rtol=2e-6, atol=1e-10), it
        assert torch.equal(stats.vis_counts.cpu(), torch.from_numpy(fix[f"stats_count_{it}"])), it
        assert torch.allclose(stats.max_2Dsize.cpu(), torch.from_numpy(fix[f"stats_max2d_{it}"]), rtol=1e-6, atol=0), it
    assert float(stats.vis_counts.max()) == 3.0 and float(stats.vis_counts.min()) == 1.0


def check_sh_basis(sh_forward, device=torch.device("cpu")):
    """The 25 basis functions of the SH evaluation (gsplat's A10 table) against the reference's OWN real-SH basis
    (nerfstudio/utils/math.py:29-92): identical functions in identical order up to the sign (-1)^|m| that gsplat's
    convention carries.  `sh_forward(degrees_to_use, dirs [N,3], coeffs [N,K,3]) -> [N,3]`; the basis is read off with
    one-hot coefficients."""
    fix = load("ref_init_small")
    dirs = torch.from_numpy(fix["sh_dirs"]).to(device)
    want = torch.from_numpy(fix["sh_components"]).double()
    n = dirs.shape[0]
    sign = torch.tensor([(-1.0) ** abs(m) for l in range(5) for m in range(-l, l + 1)], dtype=torch.float64)
    assert sign.numel() == 25
    for b in range(25):
        coeffs = torch.zeros((n, 25, 3), device=device)
        coeffs[:, b, 0], coeffs[:, b, 1], coeffs[:, b, 2] = 1.0, 2.0, -0.5
        got = torch.as_tensor(sh_forward(4, dirs, coeffs)).detach().cpu().double()
        for c, w in enumerate((1.0, 2.0, -0.5)):
            err = float((got[:, c] - w * sign[b] * want[:, b]).abs().max())
            assert err <= 2e-6, f"basis {b}, colour {c}: {err:.3e}"
    # fewer degrees in use: the higher bands are ignored
    coeffs = torch.ones((n, 25, 3), device=device)
    for deg, nb in ((0, 1), (1, 4), (2, 9), (3, 16)):
        got = torch.as_tensor(sh_forward(deg, dirs, coeffs)).detach().cpu().double()
        ref = (sign[:nb] * want[:, :nb]).sum(dim=1)
        assert float((got - ref[:, None]).abs().max()) <= 5e-6, deg


def check_sh_gradient(sh_backward, device=torch.device("cpu")):
    """d colour / d coefficients = the basis itself: `sh_backward(degrees_to_use, dirs [N,3], v_colors [N,3]) ->
    v_coeffs [N,25,3]` must be (-1)^|m| x the reference's basis (nerfstudio/utils/math.py) times v_colors, and zero in
    the bands above degrees_to_use."""
    fix = load("ref_init_small")
    dirs = torch.from_numpy(fix["sh_dirs"]).to(device)
    Y = torch.from_numpy(fix["sh_components"]).double()
    sign = torch.tensor([(-1.0) ** abs(m) for l in range(5) for m in range(-l, l + 1)], dtype=torch.float64)
    v = torch.randn((dirs.shape[0], 3), generator=torch.Generator().manual_seed(7))
    for deg, nb in ((4, 25), (2, 9), (0, 1)):
        got = torch.as_tensor(sh_backward(deg, dirs, v.to(device))).detach().cpu().double()
        want = (sign * Y)[:, :, None] * v.double()[:, None, :]
        want[:, nb:, :] = 0.0
        assert got.shape == want.shape
        assert float((got - want).abs().max()) <= 5e-6, deg


def check_quaternion_convention(quat_to_rotmat=None, project=None, device=torch.device("cpu")):
    """The (w, x, y, z) quaternion convention against the reference's OWN quaternion_matrix
    (nerfstudio/cameras/camera_utils.py:142-161): `quat_to_rotmat(q [N,4]) -> [N,3,3]` directly, and `project(means,
    scales, quats, camera) -> cov3d [N,6]` (the upper triangle of R S S^T R^T the projection emits) through
    the matrices it implies."""
    from gaussiangrasper_b200 import scenes
    fix = load("ref_init_small")
    q = torch.from_numpy(fix["quat_wxyz"])
    R = torch.from_numpy(fix["quat_rotmat"])                       # fp64, [N,3,3]
    if quat_to_rotmat is not None:
        got = torch.as_tensor(quat_to_rotmat(q.to(device))).detach().cpu().double()
        assert float((got - R).abs().max()) <= 2e-6
    if project is not None:
        n = q.shape[0]
        g = torch.Generator().manual_seed(5)
        scales = torch.rand((n, 3), generator=g) * 0.2 + 0.01
        means = (torch.rand((n, 3), generator=g) - 0.5) * 0.5      # all in front of a camera 4.5 away
        cam = scenes.look_at_camera((4.5, 0.3, 0.2), 160, 120)
        qn = q / q.norm(dim=-1, keepdim=True)                      # what the model passes (:703)
        cov3d = torch.as_tensor(project(means.to(device), scales.to(device), qn.to(device), cam)).detach().cpu().double()
        want = R @ torch.diag_embed(scales.double() ** 2) @ R.transpose(1, 2)
        tri = torch.stack([want[:, 0, 0], want[:, 0, 1], want[:, 0, 2], want[:, 1, 1], want[:, 1, 2], want[:, 2, 2]], dim=1)
        assert cov3d.shape == tri.shape
        assert float((cov3d - tri).abs().max()) <= 2e-6 * float(tri.abs().max())


def check_trainer(backend):
    """FusedAdam.reference(...).train_step(it) against the parameters the reference's OWN Trainer.train_iteration
    (engine/trainer.py:458-499) left after each of 25 iterations -- real Optimizers from the method's optimizer table,
    its gradient-accumulation table, its schedulers."""
    fix = load("ref_trainer_small")
    dev = backend.device
    names = ("means", "log_scales", "quats", "opacity_logit", "sh_coeffs", "features")
    params = {k: torch.from_numpy(fix["init_" + k]).clone().to(dev) for k in names}
    params["opacity_logit"] = params["opacity_logit"].reshape(-1)              # [N] here, [N,1] in the reference
    opt = backend.make_adam(params)
    T = int(fix["steps"][0])
    for it in range(T):
        grads = {k: torch.from_numpy(fix["grads_" + k][it]).to(dev).reshape(params[k].shape) for k in names}
        opt.bucket.pack(grads)
        opt.train_step(it)
        for k in names:
            assert opt.lrs[k] == pytest_approx(float(fix["lr_" + k][it])), (it, k, opt.lrs[k])
            want = torch.from_numpy(fix["after_" + k][it]).reshape(params[k].shape)
            assert torch.allclose(params[k].cpu(), want, rtol=3e-6, atol=2e-7), (it, k, float((params[k].cpu() - want).abs().max()))
    for k in names:
        assert opt.steps[k] == int(fix["final_step_" + k][0]), k
    assert opt.steps["means"] == 2 and opt.steps["quats"] == T


def pytest_approx(v, rel=1e-12):
    import pytest
    return pytest.approx(v, rel=rel)


def check_outputs(render, rtol_grad=2e-3):
    """One view through the whole model-side path against what the reference's OWN get_outputs + backward produced
    (ref_outputs_small.npz; the operators underneath were this repository's classes on the CPU oracle): images to
    1e-4 on every pixel the oracle does not flag as threshold-fragile, the six leaf gradients and xys.grad relative to
    each array's largest entry (99.95 % of the entries within rtol_grad, every entry within 50 rtol_grad).
    `render(params, camera, v) -> (outputs, gradients)`: params / v / outputs are dicts of CPU tensors, gradients has
    the six parameter names and "xys"."""
    from gaussiangrasper_b200 import scenes
    fix = load("ref_outputs_small")
    H, W = (int(x) for x in fix["size"])
    fx, fy, cx, cy = (float(x) for x in fix["intrinsics"])
    cam = scenes.camera_from_c2w(fix["c2w"], fx, fy, cx, cy, W, H)
    names = ("means", "log_scales", "quats", "opacity_logit", "sh_coeffs", "features")
    params = {k: torch.from_numpy(fix["param_" + k]) for k in names}
    v = {k: torch.from_numpy(fix["v_" + k]) for k in ("rgb", "feature", "depth", "normal")}
    outs, grads = render(params, cam, v)
    frag = fix["fragile"].astype(bool)
    assert frag.mean() < 0.02
    for k in v:
        got, want = np.asarray(outs[k], dtype=np.float64), fix["out_" + k].astype(np.float64)
        assert got.shape == want.shape, (k, got.shape, want.shape)
        err = np.abs(got - want)
        assert err[~frag].max() <= 1e-4, f"{k}: {err[~frag].max():.3e} on stable pixels"
        assert err.max() <= 5e-2, k
    assert float((fix["out_depth"] > 9.9).mean()) > 0 and int(fix["visible"][0]) > 500       # background and splats both occur
    for k in names + ("xys",):
        got, want = np.asarray(grads[k], dtype=np.float64).reshape(-1), fix["grad_" + k].astype(np.float64).reshape(-1)
        assert got.shape == want.shape, k
        rel = np.abs(got - want) / (np.abs(want).max() + 1e-30)
        assert np.quantile(rel, 0.9995) <= rtol_grad, f"{k}: q99.95 {np.quantile(rel, 0.9995):.3e}"
        assert rel.max() <= 50 * rtol_grad, f"{k}: max {rel.max():.3e}"
