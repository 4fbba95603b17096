"""GPU tests of the "next rows" (SURVEY 8f): fused Adam vs torch.optim.Adam with the reference's
per-group learning rates, and the densification statistics vs a torch restatement of
GaussianSplattingModel.after_train (nerfstudio/models/gaussian_splatting.py:373-393)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _params(n, D, dev, seed=0):
    g = torch.Generator().manual_seed(seed)
    shapes = dict(means=(n, 3), log_scales=(n, 3), quats=(n, 4), opacity_logit=(n, 1), sh_coeffs=(n, 25, 3),
                  features=(n, D))
    return {k: torch.randn(s, generator=g).to(dev) for k, s in shapes.items()}


def test_fused_adam_matches_torch_adam():
    from gaussiangrasper_b200.distributed import GradientBucket
    from gaussiangrasper_b200.training import REFERENCE_LRS, FusedAdam
    dev = torch.device("cuda:0")
    n, D = 5003, 16
    ours = _params(n, D, dev)
    ref = {k: v.clone().requires_grad_(True) for k, v in ours.items()}
    opt_ref = torch.optim.Adam([{"params": [ref[k]], "lr": REFERENCE_LRS[k]} for k in ref], eps=1e-15)
    bucket = GradientBucket(ours)
    opt = FusedAdam(ours, bucket)
    g = torch.Generator().manual_seed(1)
    for step in range(4):
        grads = {k: (torch.randn(v.shape, generator=g) * (10.0 ** (step - 2))).to(dev) for k, v in ours.items()}
        bucket.pack(grads)
        lr_scale = 0.9 ** step  # a scheduler changing the learning rates between steps
        opt.step({k: REFERENCE_LRS[k] * lr_scale for k in ours})
        for grp, k in zip(opt_ref.param_groups, ref):
            grp["lr"] = REFERENCE_LRS[k] * lr_scale
            ref[k].grad = grads[k].clone()
        opt_ref.step()
        for k in ours:
            assert torch.allclose(ours[k], ref[k].detach(), rtol=2e-6, atol=1e-7), (step, k)


def _after_train_torch(state, xys_grad, radii, last_size):
    visible = radii > 0
    grads = xys_grad.norm(dim=-1)
    if state["norm"] is None:
        state["norm"] = grads.clone()
        state["cnt"] = torch.ones_like(grads)
    else:
        state["cnt"][visible] += 1
        state["norm"][visible] += grads[visible]
    if state["max"] is None:
        state["max"] = torch.zeros_like(grads)
    state["max"][visible] = torch.maximum(state["max"][visible], radii[visible].float() / float(max(last_size)))


def test_densify_stats_match_after_train():
    from gaussiangrasper_b200.training import DensifyStats
    dev = torch.device("cuda:0")
    n, H, W = 7001, 480, 640
    stats = DensifyStats(n, dev)
    state = dict(norm=None, cnt=None, max=None)
    g = torch.Generator().manual_seed(3)
    for step in range(3):
        v_geo = torch.randn((n, 8), generator=g).to(dev)
        radii = torch.randint(-1, 40, (n,), generator=g, dtype=torch.int32).clamp(min=0).to(dev)
        radii[::3] = 0
        v_geo[radii == 0, :2] = 0  # invisible Gaussians receive no gradient
        stats.update(v_geo, radii, H, W)
        _after_train_torch(state, v_geo[:, :2], radii, (H, W))
        assert torch.allclose(stats.xys_grad_norm, state["norm"], rtol=1e-6, atol=1e-7)
        assert torch.equal(stats.vis_counts, state["cnt"])
        assert torch.allclose(stats.max_2Dsize, state["max"], rtol=1e-6)


def test_training_step_decreases_loss():
    """render_views -> backward into the bucket -> fused Adam, a few iterations on a tiny scene."""
    from gaussiangrasper_b200 import scenes
    from gaussiangrasper_b200.distributed import GradientBucket
    from gaussiangrasper_b200.render import ViewBatch, render_views
    from gaussiangrasper_b200.training import DensifyStats, FusedAdam
    dev = torch.device("cuda:0")
    n, W, H, D = 2000, 64, 48, 4
    sc = scenes.random_scene(n, feature_dim=D, seed=11)
    sc["log_scales"] = sc["log_scales"] + 1.0
    P = {k: v.to(dev).contiguous().requires_grad_(True) for k, v in sc.items()}
    cams = scenes.orbit_cameras(2, W, H, total=6)
    vb = ViewBatch.from_cameras(cams, dev)
    with torch.no_grad():
        target = render_views(*(P[k] for k in ("means", "log_scales", "quats", "opacity_logit", "sh_coeffs", "features")),
                              vb)["image"].clone()
        P["sh_coeffs"] += 0.3 * torch.randn_like(P["sh_coeffs"])
        P["features"] += 0.3 * torch.randn_like(P["features"])
    bucket = GradientBucket({k: v.detach() for k, v in P.items()})
    opt = FusedAdam({k: v.detach() for k, v in P.items()}, bucket, lrs=dict(means=0.0, log_scales=0.0, quats=0.0,
                                                                           opacity_logit=0.0, sh_coeffs=2e-2, features=2e-2))
    stats = DensifyStats(n, dev)
    losses = []
    for it in range(12):
        holder = {"grad_out": bucket.unpack()}
        out = render_views(*(P[k] for k in ("means", "log_scales", "quats", "opacity_logit", "sh_coeffs", "features")),
                           vb, holder=holder)
        loss = ((out["image"] - target) ** 2).mean()
        for p in P.values():
            p.grad = None
        loss.backward()
        stats.update(holder["v_geo"], holder["radii"].reshape(-1), H, W)
        opt.step()
        losses.append(float(loss))
    assert losses[-1] < 0.5 * losses[0], losses
    assert float(stats.vis_counts.max()) >= 12
