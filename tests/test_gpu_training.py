"""GPU tests of the "next rows" (SURVEY 8f): fused Adam vs torch.optim.Adam with the reference's
per-group learning rates, and the densification statistics vs a torch restatement of
GaussianSplattingModel.after_train (nerfstudio/models/gaussian_splatting.py:373-393)."""
import pytest
import torch

from gaussiangrasper_b200 import scenes

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


def _refine_inputs(n, D, seed):
    g = torch.Generator().manual_seed(seed)
    sc = scenes.random_scene(n, feature_dim=D, seed=seed)
    P = {k: sc[k].clone() for k in ("means", "log_scales", "quats", "opacity_logit", "sh_coeffs", "features")}
    P["opacity_logit"] = P["opacity_logit"].reshape(-1)
    # scales around the 0.01 split threshold, a few huge ones; opacities around the 0.1 cull threshold
    P["log_scales"] = torch.log(torch.rand((n, 3), generator=g) * 0.03 + 0.002)
    P["log_scales"][::97] = torch.log(torch.tensor(0.9))
    P["opacity_logit"] = torch.logit(torch.rand(n, generator=g) * 0.5 + 0.01)
    M = {k: (torch.randn(v.shape, generator=g), torch.rand(v.shape, generator=g)) for k, v in P.items()}
    stats = dict(xys_grad_norm=torch.rand(n, generator=g) * 4e-6 * 3, vis_counts=torch.randint(1, 4, (n,), generator=g).float(),
                 max_2dsize=torch.rand(n, generator=g) * 0.2)
    return P, M, stats


@pytest.mark.parametrize("rules_kw", [dict(), dict(do_densify=0), dict(do_cull=0), dict(split_by_screen=0, cull_by_screen=0),
                                      dict(cull_by_scale=0)])
def test_refine_matches_reference_flow(rules_kw):
    """Densify + cull + Adam-state surgery on the device vs the line-by-line restatement of
    refinement_after / split_gaussians / dup_gaussians / cull_gaussians (oracle/refine_oracle.py): same
    survivors in the same order, copied rows bit-identical, new means / shrunk scales within 1e-6."""
    from gaussiangrasper_b200.training import DensifyStats, refine_gaussians
    from oracle import refine_oracle
    dev = torch.device("cuda:0")
    n, D = 20_003, 5   # not a multiple of 4: the scan arrays of the plan are padded apart
    P, M, st = _refine_inputs(n, D, 12)
    rules = dict(max_dim=640.0, densify_grad_thresh=0.0002, densify_size_thresh=0.01, split_screen_size=0.05,
                 cull_alpha_thresh=0.1, cull_scale_thresh=0.5, cull_screen_size=0.15, do_densify=1, split_by_screen=1,
                 do_cull=1, cull_by_scale=1, cull_by_screen=1)
    rules.update(rules_kw)
    g = torch.Generator().manual_seed(5)
    z_all = torch.randn((2 * n, 3), generator=g)
    want_p, want_m, info = refine_oracle.refine(P, M, st["xys_grad_norm"], st["vis_counts"], st["max_2dsize"], rules,
                                                lambda k: z_all[:k])
    assert info["fragile"] == 0, "pick another seed: a decision sits on a threshold"
    ds = DensifyStats(n, dev)
    ds.xys_grad_norm, ds.vis_counts, ds.max_2Dsize = (st[k].to(dev) for k in ("xys_grad_norm", "vis_counts", "max_2dsize"))
    got_p, got_m, ginfo = refine_gaussians({k: v.to(dev) for k, v in P.items()},
                                           {k: (a.to(dev), b.to(dev)) for k, (a, b) in M.items()}, ds, rules,
                                           samples_fn=lambda k: z_all[:k].to(dev))
    assert ginfo["n_out"] == info["n_out"] and ginfo["n_in"] == n
    if rules["do_densify"] and rules["do_cull"]:
        assert ginfo["n_split"] > 100 and ginfo["n_dup_kept"] > 100 and info["n_culled"] > 100  # the case is not trivial
    for k in refine_oracle.PARAMS:
        a, b = got_p[k].cpu(), want_p[k]
        assert a.shape == b.shape, k
        if k in ("means", "log_scales"):
            assert torch.allclose(a, b, rtol=2e-6, atol=2e-6), k
            same = (a == b).all(dim=-1)
            assert int(same.sum()) >= ginfo["n_kept"] - ginfo["n_split"]  # untouched rows are copies
        else:
            assert torch.equal(a, b), k
        for j in range(2):
            assert torch.equal(got_m[k][j].cpu(), want_m[k][j]), (k, j)


def test_refine_everything_culled_and_no_moments():
    from gaussiangrasper_b200.training import refine_gaussians, refine_schedule
    dev = torch.device("cuda:0")
    P, _, _ = _refine_inputs(300, 3, 2)
    P["opacity_logit"][:] = -10.0
    rules = refine_schedule(3300, 10, 640)  # past warm-up and the post-reset window: cull is on
    assert rules["do_cull"] and rules["do_densify"] and not rules["reset_opacity"]
    rules.update(do_densify=0, split_by_screen=0, cull_by_screen=0)   # no statistics handed in
    with pytest.raises(Exception):
        refine_gaussians({k: v.to(dev) for k, v in P.items()}, None, None, dict(rules, cull_by_screen=1))
    newp, newm, info = refine_gaussians({k: v.to(dev) for k, v in P.items()}, None, None, rules)
    assert info["n_out"] == 0 and newm is None and newp["sh_coeffs"].shape == (0, 25, 3)
    assert refine_schedule(100, 10, 640) is None and refine_schedule(3100, 10, 640)["reset_opacity"]


def test_training_loop_with_refinement():
    """render -> loss -> backward -> fused Adam -> statistics, a refinement (split / duplicate / cull with the
    optimizer state carried over) in the middle, then more steps on the refined set."""
    from gaussiangrasper_b200.render import ViewBatch, render_views
    from gaussiangrasper_b200.training import DensifyStats, FusedAdam, refine_gaussians
    dev = torch.device("cuda:0")
    n, W, H, D = 3000, 96, 64, 4
    sc = scenes.random_scene(n, feature_dim=D, seed=21)
    sc["log_scales"] = sc["log_scales"] + 0.8
    names = ("means", "log_scales", "quats", "opacity_logit", "sh_coeffs", "features")
    P = {k: sc[k].to(dev).clone().contiguous().requires_grad_(True) for k in names}
    cams = scenes.orbit_cameras(2, W, H, total=6)
    vb = ViewBatch.from_cameras(cams, dev)
    target = torch.rand((2, H, W, 7 + D), generator=torch.Generator().manual_seed(0)).to(dev)
    opt = FusedAdam(P)
    stats = DensifyStats(n, dev)

    def run(steps):
        losses = []
        for _ in range(steps):
            holder = {"grad_out": opt.bucket.unpack()}
            out = render_views(*(P[k] for k in names), vb, holder=holder)
            loss = ((out["image"][..., :7 + D] - target) ** 2).mean()
            loss.backward()
            stats.update(holder["v_geo"], holder["radii"], H, W)
            opt.step()
            for p in P.values():
                p.grad = None
            losses.append(float(loss.detach()))
        return losses

    first = run(6)
    rules = dict(max_dim=float(max(W, H)), densify_grad_thresh=1e-7, densify_size_thresh=0.05, split_screen_size=0.05,
                 cull_alpha_thresh=0.1, cull_scale_thresh=0.5, cull_screen_size=0.15, do_densify=1, split_by_screen=1,
                 do_cull=1, cull_by_scale=0, cull_by_screen=0)
    new_p, new_m, info = refine_gaussians({k: P[k].detach() for k in names}, opt.moments(), stats, rules)
    assert info["n_split"] > 0 and info["n_dup_kept"] > 0 and info["n_out"] != n
    n2 = info["n_out"]
    assert new_p["opacity_logit"].shape == (n2, 1) and new_p["sh_coeffs"].shape == (n2, 25, 3)
    P = {k: new_p[k].requires_grad_(True) for k in names}
    opt.rebuild(P, new_m)
    assert opt.bucket.payload == sum(p.numel() for p in P.values()) <= opt.exp_avg.numel() and opt.t == 6 and set(opt.steps.values()) == {6}
    # survivors kept their first moments, children start from zero
    ea = opt.moments()["means"][0]
    assert float(ea[:info["n_kept"]].abs().max()) > 0 and float(ea[info["n_kept"]:].abs().max()) == 0.0
    stats = DensifyStats(n2, dev)
    second = run(6)
    assert all(torch.isfinite(torch.tensor(first + second)))
    assert second[-1] < first[0]
    opt.reset_opacity(0.1)
    assert torch.allclose(P["opacity_logit"].detach(), torch.full_like(P["opacity_logit"], float(torch.logit(torch.tensor(0.08)))))
    assert float(opt.moments()["opacity_logit"][1].abs().max()) == 0.0


@pytest.mark.parametrize("kind", ["l1", "l2"])
@pytest.mark.parametrize("masked", [False, True])
def test_pixel_loss_matches_torch(kind, masked):
    """Fused loss + gradient vs the reference's expressions (gaussian_splatting.py:853-866): masked pixels are
    zeroed in both images and still count in the mean."""
    from gaussiangrasper_b200.training import pixel_loss
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(4)
    pred = torch.randn((2, 37, 53, 7), generator=g).to(dev).requires_grad_(True)
    target = torch.randn((2, 37, 53, 7), generator=g).to(dev)
    mask = (torch.rand((2, 37, 53), generator=g) > 0.3).to(dev) if masked else None
    loss, grad = pixel_loss(pred, target, kind, mask, weight=0.8)
    d = (pred[mask] - target[mask]) if masked else (pred - target)        # :882 averages over the valid pixels
    ref = 0.8 * (d.abs().mean() if kind == "l1" else (d ** 2).mean())
    ref.backward()
    assert torch.allclose(loss[0], ref.detach(), rtol=1e-5, atol=1e-8)
    assert torch.allclose(grad, pred.grad, rtol=1e-6, atol=1e-12)
    if masked:  # mean over all pixels, ignored ones contributing zero
        loss_all, grad_all = pixel_loss(pred, target, kind, mask, weight=0.8, mean_over="all")
        frac = float(mask.float().mean())
        assert torch.allclose(loss_all, loss * frac, rtol=1e-5) and torch.allclose(grad_all, grad * frac, rtol=1e-5, atol=1e-12)
    # a second call reuses the workspace (the block counter resets itself)
    loss2, _ = pixel_loss(pred, target, kind, mask, weight=0.8)
    assert torch.equal(loss, loss2)
    with pytest.raises(ValueError):
        pixel_loss(pred, target[:1], kind)


def test_ssim_loss_and_main_loss_match_the_restatement():
    """SSIM loss + gradient vs oracle/loss_oracle.py (pytorch_msssim's algorithm, autograd gradient), alone and
    accumulated onto the L1 term as the reference's main_loss (gaussian_splatting.py:882-885, :931)."""
    from gaussiangrasper_b200.training import pixel_loss, ssim_loss
    from oracle import loss_oracle
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(8)
    H, W = 45, 70   # partial tiles in both directions
    base = torch.rand((H, W, 3), generator=g)
    gt = (base + 0.1 * torch.randn((H, W, 3), generator=g)).clamp(0, 1)
    pred_full = torch.rand((H, W, 8), generator=g)          # rgb + 5 more channels, like the blended image
    pred_full[..., :3] = (base + 0.2 * torch.randn((H, W, 3), generator=g)).clamp(0, 1)
    x = pred_full[..., :3].clone().double().requires_grad_(True)
    ref = 1 - loss_oracle.ssim(gt.double().permute(2, 0, 1)[None], x.permute(2, 0, 1)[None])
    ref.backward()
    loss, grad = ssim_loss(pred_full.to(dev), gt.to(dev))
    assert abs(float(loss) - float(ref)) < 2e-6
    gmax = float(x.grad.abs().max())
    assert float((grad[..., :3].cpu().double() - x.grad).abs().max()) < 2e-5 * gmax + 1e-10
    assert float(grad[..., 3:].abs().max()) == 0.0
    # main loss: L1 (rgb only) then SSIM accumulated on top
    lam = 0.2
    x2 = pred_full[..., :3].clone().double().requires_grad_(True)
    ref2 = loss_oracle.main_loss(x2, gt.double(), lam)
    ref2.backward()
    p3 = pred_full[..., :3].contiguous().to(dev)
    l1, g1 = pixel_loss(p3, gt.to(dev), "l1", weight=1 - lam)
    tot, gtot = ssim_loss(p3, gt.to(dev), weight=lam, grad=g1, loss=l1)
    assert abs(float(tot) - float(ref2)) < 2e-6
    assert float((gtot.cpu().double() - x2.grad).abs().max()) < 2e-5 * float(x2.grad.abs().max()) + 1e-10
    # a batch of two images equals the mean of the two single-image values
    two = torch.stack([pred_full, pred_full.flip(0)]).to(dev)
    gt2 = torch.stack([gt, gt.flip(0)]).to(dev)
    lb, _ = ssim_loss(two, gt2)
    assert abs(float(lb) - float(loss)) < 2e-6
    with pytest.raises(ValueError):
        ssim_loss(torch.zeros((5, 70, 3), device=dev), torch.zeros((5, 70, 3), device=dev))


def test_fused_adam_accumulation_and_schedules_match_the_trainer_loop():
    """FusedAdam.train_step == the reference trainer's iteration (engine/trainer.py:466-497) driven with
    torch.optim.Adam per group, its accumulation table (method_configs.py:611) and LambdaLR exponential decay
    (engine/schedulers.py:116-141): zero_grad at it % k == 0, step at it % k == k-1, schedulers stepped every
    iteration."""
    import numpy as np
    from gaussiangrasper_b200.distributed import GradientBucket
    from gaussiangrasper_b200.training import (REFERENCE_ACCUMULATION, REFERENCE_LRS, REFERENCE_SCHEDULES, FusedAdam)
    dev = torch.device("cuda:0")
    n, D = 3001, 7
    ours = _params(n, D, dev, seed=4)
    ref = {k: v.clone().requires_grad_(True) for k, v in ours.items()}
    short = {k: (lf, 25) for k, (lf, _) in REFERENCE_SCHEDULES.items()}   # decay visible within the test's 23 steps
    opt = FusedAdam(ours, GradientBucket(ours), accumulation=REFERENCE_ACCUMULATION, schedules=short)
    ref_opt = {k: torch.optim.Adam([ref[k]], lr=REFERENCE_LRS[k], eps=1e-15) for k in ref}
    ref_sched = {}
    for k, (lr_final, max_steps) in short.items():
        lr0 = REFERENCE_LRS[k]

        def func(step, lr0=lr0, lr_final=lr_final, max_steps=max_steps):
            t = np.clip(step / max_steps, 0, 1)
            return np.exp(np.log(lr0) * (1 - t) + np.log(lr_final) * t) / lr0
        ref_sched[k] = torch.optim.lr_scheduler.LambdaLR(ref_opt[k], lr_lambda=func)
    g = torch.Generator().manual_seed(9)
    for it in range(23):
        grads = {k: torch.randn(v.shape, generator=g).to(dev) for k, v in ours.items()}
        # reference loop
        for k in ref:
            a = REFERENCE_ACCUMULATION.get(k, 1)
            if it % a == 0:
                ref_opt[k].zero_grad()
            ref[k].grad = grads[k].clone() if ref[k].grad is None else ref[k].grad + grads[k]
            if it % a == a - 1:
                ref_opt[k].step()
        for sch in ref_sched.values():
            sch.step()
        # ours
        opt.bucket.pack(grads)
        opt.train_step(it)
        for k in ours:
            assert torch.allclose(ours[k], ref[k].detach(), rtol=3e-6, atol=2e-7), (it, k)
    assert opt.steps["means"] == 2 and opt.steps["quats"] == 23
    # optimizer state round trip (what a checkpoint carries)
    state = opt.state_dict()
    opt2 = FusedAdam({k: v.clone() for k, v in ours.items()}, accumulation=REFERENCE_ACCUMULATION, schedules=short)
    opt2.load_state_dict(state)
    assert opt2.steps == opt.steps and torch.equal(opt2.exp_avg, opt.exp_avg) and torch.equal(opt2.exp_avg_sq, opt.exp_avg_sq)


def test_refine_with_counter_based_samples():
    """samples == NULL: the split children are drawn inside the kernel from Philox keyed on (seed, step, parent
    row, sample index).  Same result as handing the kernel those numbers explicitly (gg_philox_normals for the
    split parents in row order), the numbers are standard normal, and two calls agree bit for bit."""
    from gaussiangrasper_b200.training import DensifyStats, philox_normals, refine_gaussians
    from oracle import refine_oracle
    dev = torch.device("cuda:0")
    n, D = 20_003, 5
    P, M, st = _refine_inputs(n, D, 12)
    rules = dict(max_dim=640.0, densify_grad_thresh=0.0002, densify_size_thresh=0.01, split_screen_size=0.05,
                 cull_alpha_thresh=0.1, cull_scale_thresh=0.5, cull_screen_size=0.15, do_densify=1, split_by_screen=1,
                 do_cull=1, cull_by_scale=1, cull_by_screen=1)

    def stats():
        ds = DensifyStats(n, dev)
        ds.xys_grad_norm, ds.vis_counts, ds.max_2Dsize = (st[k].to(dev) for k in ("xys_grad_norm", "vis_counts", "max_2dsize"))
        return ds
    Pd = {k: v.to(dev) for k, v in P.items()}
    Md = {k: (a.to(dev), b.to(dev)) for k, (a, b) in M.items()}
    seed, step = 0x1234_5678_9ABC_DEF0, 700
    a_p, a_m, a_info = refine_gaussians(Pd, Md, stats(), rules, seed=seed, step=step)
    b_p, b_m, b_info = refine_gaussians(Pd, Md, stats(), rules, seed=seed, step=step)
    for k in a_p:
        assert torch.equal(a_p[k], b_p[k]), k
    # which rows split (the oracle's rule, :411-418)
    avg = (st["xys_grad_norm"] / st["vis_counts"]) * 0.5 * rules["max_dim"]
    smax = P["log_scales"].exp().max(dim=-1).values
    splits = ((smax > rules["densify_size_thresh"]) | (st["max_2dsize"] > rules["split_screen_size"])) & \
             (avg > rules["densify_grad_thresh"])
    parents = torch.nonzero(splits).reshape(-1)
    assert parents.numel() == a_info["n_split"] > 100
    z = philox_normals(parents.numel(), 2, seed, step, dev, parents=parents.to(dev))     # [2, n_split, 3]
    c_p, c_m, c_info = refine_gaussians(Pd, Md, stats(), rules, samples_fn=lambda k: z.reshape(-1, 3)[:k])
    for k in a_p:
        assert torch.equal(a_p[k], c_p[k]), k
    # the oracle with the same numbers
    want_p, _, info = refine_oracle.refine(P, M, st["xys_grad_norm"], st["vis_counts"], st["max_2dsize"], rules,
                                           lambda k: z.reshape(-1, 3)[:k].cpu())
    assert torch.allclose(a_p["means"].cpu(), want_p["means"], rtol=2e-6, atol=2e-6)
    # another step or seed gives other children; the draws are standard normal
    d_p, _, _ = refine_gaussians(Pd, Md, stats(), rules, seed=seed, step=step + 100)
    assert not torch.equal(d_p["means"], a_p["means"])
    big = philox_normals(200_000, 2, seed, 3, dev).reshape(-1)
    assert abs(float(big.mean())) < 5e-3 and abs(float(big.std()) - 1.0) < 5e-3
    assert abs(float((big ** 4).mean()) - 3.0) < 0.05 and float(big.abs().max()) < 7.0


def _two_rank_refine_worker(rank, world, port, out_dir):
    import os
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch.distributed as dist
    dist.init_process_group("gloo", rank=rank, world_size=world)   # gloo moves the CUDA tensors through the host
    from gaussiangrasper_b200 import scenes
    from gaussiangrasper_b200.render import ViewBatch, render_views
    from gaussiangrasper_b200.training import DensifyStats, FusedAdam, refine_gaussians
    dev = torch.device("cuda:0")
    n, W, H, D = 6000, 96, 64, 4
    sc = scenes.random_scene(n, feature_dim=D, seed=5)
    sc["log_scales"] = sc["log_scales"] + 0.8
    names = ("means", "log_scales", "quats", "opacity_logit", "sh_coeffs", "features")
    P = {k: sc[k].to(dev).requires_grad_(True) for k in names}
    opt = FusedAdam(P)
    stats = DensifyStats(n, dev)
    cams = scenes.orbit_cameras(2 * world, W, H, total=2 * world)
    for it in range(2):   # two iterations, each rank renders its own view, gradients summed over the ranks
        v = it * world + rank
        holder = {"grad_out": opt.bucket.unpack()}
        out = render_views(*(P[k] for k in names), ViewBatch.from_cameras([cams[v]], dev), holder=holder)
        out["image"].backward(torch.randn((1, H, W, out["image"].shape[-1]), generator=torch.Generator().manual_seed(v)).to(dev) * 1e-3)
        stats.update(holder["v_geo"], holder["radii"], H, W)
        opt.bucket.all_reduce()
        with torch.no_grad():
            opt.step()
    rules = dict(max_dim=float(max(W, H)), densify_grad_thresh=2e-7, densify_size_thresh=0.03, split_screen_size=0.05,
                 cull_alpha_thresh=0.1, cull_scale_thresh=0.5, cull_screen_size=0.15, do_densify=1, split_by_screen=1,
                 do_cull=1, cull_by_scale=0, cull_by_screen=0)
    local_counts = stats.vis_counts.clone()
    new_p, new_m, info = refine_gaussians({k: P[k].detach() for k in names}, opt.moments(), stats, rules, seed=77, step=2)
    torch.save(dict(p={k: v.cpu() for k, v in new_p.items()}, m={k: (a.cpu(), b.cpu()) for k, (a, b) in new_m.items()},
                    info=info, local_counts=local_counts.cpu(), counts=stats.vis_counts.cpu()),
               os.path.join(out_dir, f"rank{rank}.pt"))
    dist.barrier()
    dist.destroy_process_group()


def test_two_ranks_refine_to_byte_identical_parameters(tmp_path):
    """View-sharded refinement: two ranks (two processes on this GPU, gloo) render different
    views, all-reduce gradients, accumulate their own densification statistics -- and after refine_gaussians hold
    byte-identical parameters and Adam moments: statistics all-reduced, split samples counter-based."""
    import torch.multiprocessing as mp
    import socket
    with socket.socket() as s_:
        s_.bind(("127.0.0.1", 0))
        port = s_.getsockname()[1]
    mp.spawn(_two_rank_refine_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    a, b = (torch.load(tmp_path / f"rank{r}.pt") for r in range(2))
    assert a["info"] == b["info"] and a["info"]["n_split"] > 10 and a["info"]["n_out"] != a["info"]["n_in"]
    assert not torch.equal(a["local_counts"], b["local_counts"])     # the ranks really saw different views
    assert torch.equal(a["counts"], b["counts"])
    for k in a["p"]:
        assert a["p"][k].numpy().tobytes() == b["p"][k].numpy().tobytes(), k
        for j in range(2):
            assert a["m"][k][j].numpy().tobytes() == b["m"][k][j].numpy().tobytes(), (k, j)


def test_captured_step_replays_the_same_render_and_gradients():
    """graph.CapturedStep: a whole render + backward captured into a CUDA graph (the binning reads nothing back) gives,
    on replay, the images and gradients of the plain launches -- also after the parameters and the cameras were
    updated in place -- and reports the intersection count without any host read inside the step."""
    from gaussiangrasper_b200.distributed import GradientBucket
    from gaussiangrasper_b200.graph import CapturedStep
    from gaussiangrasper_b200.render import ViewBatch, render_views
    dev = torch.device("cuda:0")
    n, W, H, D = 20_000, 160, 120, 5
    sc = scenes.random_scene(n, feature_dim=D, seed=21)
    sc["log_scales"] = sc["log_scales"] + 0.5
    names = ("means", "log_scales", "quats", "opacity_logit", "sh_coeffs", "features")
    P = {k: sc[k].to(dev).requires_grad_(True) for k in names}
    cams = scenes.orbit_cameras(4, W, H, total=9)
    vb = ViewBatch.from_cameras(cams[:2], dev)
    v = torch.randn((2, H, W, 12), generator=torch.Generator().manual_seed(0)).to(dev)
    bucket = GradientBucket(P)

    def fn():
        out = render_views(*(P[k] for k in names), vb, holder={"grad_out": bucket.unpack()})
        out["image"].backward(v)
        return out["image"].detach()

    def plain(view_cams):
        Q = {k: P[k].detach().clone().requires_grad_(True) for k in names}
        out = render_views(*(Q[k] for k in names), ViewBatch.from_cameras(view_cams, dev))
        out["image"].backward(v)
        return out["image"].detach(), {k: Q[k].grad for k in names}

    cap = CapturedStep(fn, dev, warmup=2)
    assert cap.launches >= 5
    for rep, view_cams in enumerate((cams[:2], cams[2:4])):
        if rep == 1:     # new cameras and perturbed parameters, in place: the graph reads the same storage
            vb.update_(view_cams)
            with torch.no_grad():
                P["means"].add_(0.01 * torch.randn_like(P["means"]))
                P["features"].mul_(0.9)
        bucket.flat.fill_(float("nan"))          # whatever the replay does not write would show
        img = cap.replay()
        m = cap.check()
        want_img, want_grad = plain(view_cams)
        assert m > 10_000
        assert torch.allclose(img, want_img, rtol=0, atol=2e-5), rep
        for k in names:
            ref = want_grad[k].reshape(bucket.view(k).shape)
            assert torch.allclose(bucket.view(k), ref, rtol=2e-4, atol=2e-6 * float(ref.abs().max())), (rep, k)


@pytest.mark.parametrize("n", [1, 2, 3, 5, 3000, 20_011])
def test_knn_scale_init_matches_the_reference_recipe(n):
    """populate_modules' scale initialisation (:259-263): sklearn k-NN there, exact brute force here -- compared with
    sklearn itself when the set is large enough for its k + 1 query, and with torch.cdist otherwise."""
    import numpy as np
    from gaussiangrasper_b200.training import knn_scale_init
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(n)
    means = (torch.rand((n, 3), generator=g) - 0.5) * 10
    if n >= 3000:
        means[7] = means[11]          # an exact duplicate: its nearest "other" point sits at distance zero
    ls, dist = knn_scale_init(means.to(dev))
    d = torch.cdist(means.double(), means.double())
    d.fill_diagonal_(float("inf"))
    k = min(3, n - 1)
    want = torch.sort(d, dim=1)[0][:, :k] if k > 0 else torch.zeros((n, 0), dtype=torch.float64)
    got = dist.cpu().double()
    if k > 0:
        assert torch.allclose(got[:, :k], want, rtol=1e-5, atol=1e-6)
    if k < 3:
        fill = want[:, -1:] if k > 0 else torch.zeros((n, 1), dtype=torch.float64)
        assert torch.allclose(got[:, k:], fill.expand(n, 3 - k), rtol=1e-5, atol=1e-6)
    if n >= 4:
        from sklearn.neighbors import NearestNeighbors
        dist_sk, _ = NearestNeighbors(n_neighbors=4, algorithm="auto", metric="euclidean").fit(means.numpy()).kneighbors(means.numpy())
        avg = torch.from_numpy(dist_sk[:, 1:].astype(np.float32)).mean(dim=-1, keepdim=True)
        ref = torch.log(avg.repeat(1, 3))
        ok = torch.isfinite(ref)
        assert torch.allclose(ls.cpu()[ok], ref[ok], rtol=0, atol=2e-5)
