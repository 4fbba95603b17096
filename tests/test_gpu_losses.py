"""GPU tests of the fused loss kernels (csrc/loss.cu) against the line-by-line torch restatement of
GaussianSplattingModel.get_loss_dict (oracle/loss_oracle.py, fp64 autograd): values and gradients."""
import pytest
import torch
import torch.nn.functional as F

from oracle import loss_oracle

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    assert torch.cuda.is_available()
    return torch.device("cuda:0")


def _image(H, W, CP, seed):
    g = torch.Generator().manual_seed(seed)
    return torch.randn((H, W, CP), generator=g)


@pytest.mark.parametrize("H,W,CP,accumulate", [(48, 64, 24, False), (37, 53, 12, True), (480, 640, 24, False)])
def test_geom_loss_matches_the_restatement(dev, H, W, CP, accumulate):
    from gaussiangrasper_b200.losses import geom_loss
    img = _image(H, W, CP, 1)
    img[..., 3] = img[..., 3].abs() * 2 + 0.1
    g = torch.Generator().manual_seed(2)
    gt_depth = torch.rand((1, H, W), generator=g) * 4
    gt_depth[0, :5] = 0.01                                   # below the 0.05 validity threshold (:861)
    gt_normal = F.normalize(torch.randn((3, H, W), generator=g), dim=0)
    mask = (gt_depth > 0.05) & (torch.rand((1, H, W), generator=g) > 0.2)
    img[3, 4, 4:7] = 0.0                                     # a zero normal inside the mask: F.normalize's eps branch
    x = img.double().requires_grad_(True)
    d_ref, n_ref = loss_oracle.geom_losses(x[..., 4:7], x[..., 3:4], gt_normal.double(), gt_depth.double(), mask)
    (0.7 * d_ref + 1.3 * n_ref).backward()
    base = torch.randn((H, W, CP), generator=g).to(dev) if accumulate else None
    loss, grad = geom_loss(img.to(dev), gt_depth[0].to(dev), gt_normal.permute(1, 2, 0).contiguous().to(dev),
                           mask[0].to(dev), w_depth=0.7, w_normal=1.3, grad=base.clone() if accumulate else None)
    assert loss[0].item() == pytest.approx(0.7 * d_ref.item(), rel=2e-5)
    assert loss[1].item() == pytest.approx(1.3 * n_ref.item(), rel=2e-5)
    want = x.grad.float()
    got = grad.cpu() - (base.cpu() if accumulate else 0)
    # (base + g) - base carries the rounding of the sum: eps_fp32 of |base| (up to ~5) on top of the kernel's own error
    slack = 6e-7 if accumulate else 1e-9
    assert float((got[..., 3:7] - want[..., 3:7]).abs().max()) <= 2e-5 * float(want.abs().max()) + slack
    other = got.clone()
    other[..., 3:7] = 0
    assert float(other.abs().max()) <= (1e-6 if accumulate else 0.0)   # nothing else is touched


def test_contrastive_feature_loss_and_up_loss(dev):
    from gaussiangrasper_b200 import losses
    H, W, D = 60, 80, 32
    CP = 40
    img = _image(H, W, CP, 3)
    g = torch.Generator().manual_seed(4)
    seg = torch.randint(-1, 5, (H, W), generator=g).float()            # segment ids, -1 = invalid (:873)
    pairs = losses.sampling_pairs_in_mask(seg, 200, generator=g)
    points = losses.sampling_in_mask(seg, 300, generator=g)
    assert len(pairs) == 5 and points.shape[0] == 300 and all(p[0].shape == p[1].shape == (200, 2) for p in pairs)
    for p1, p2 in pairs:                                               # pairs stay inside one segment
        assert torch.equal(seg[p1[:, 0], p1[:, 1]], seg[p2[:, 0], p2[:, 1]])
    gt_fea = torch.randn((512, H, W), generator=g)
    mlp_ref = loss_oracle.MLP(in_dim=D, out_dim=512, hidden_list=[128]).double()
    x = img.double().requires_grad_(True)
    f_ref = loss_oracle.feature_loss(x[..., 7:7 + D], pairs)
    u_ref = loss_oracle.up_loss(x[..., 7:7 + D], points, gt_fea.double(), mlp_ref)
    (f_ref + 0.5 * u_ref).backward()

    mlp = losses.UpProjection(D).to(dev)
    mlp.load_state_dict({k: v.float() for k, v in mlp_ref.state_dict().items()})   # same parameter names as the reference's MLP
    grad = torch.zeros((H, W, CP), device=dev)
    d_pairs = [[a.to(dev), b.to(dev)] for a, b in pairs]
    lf = losses.contrastive_feature_loss(img.to(dev), d_pairs, grad, feature_dim=D)
    lu = losses.up_loss(img.to(dev), points.to(dev), gt_fea.permute(1, 2, 0).contiguous().to(dev), mlp, grad, feature_dim=D,
                        weight=0.5)
    assert lf.item() == pytest.approx(f_ref.item(), rel=2e-5)
    assert lu.item() == pytest.approx(0.5 * u_ref.item(), rel=2e-5)
    want = x.grad.float()
    assert float((grad.cpu() - want).abs().max()) <= 3e-5 * float(want.abs().max())
    assert float(grad[..., :7].abs().max()) == 0.0 and float(grad[..., 7 + D:].abs().max()) == 0.0
    for (k, p), (k2, p2) in zip(mlp.named_parameters(), mlp_ref.named_parameters()):
        assert k == k2
        # (the reference side was back-propagated from f + 0.5 u: its parameter gradients carry the 0.5 already)
        assert float((p.grad.cpu() - p2.grad.float()).abs().max()) <= 3e-5 * float(p2.grad.abs().max()) + 1e-9, k


@pytest.mark.parametrize("n", [1, 1000, 100_003])
def test_param_regs(dev, n):
    from gaussiangrasper_b200.losses import param_regs
    g = torch.Generator().manual_seed(n)
    sh = torch.randn((n, 25, 3), generator=g)
    sh[::7, 1:, 0] = 0.0                                               # zero norm: no gradient (torch gives 0 there too)
    ls = torch.randn((n, 3), generator=g) * 1.5
    a, b = sh.double().requires_grad_(True), ls.double().requires_grad_(True)
    r_sh, r_sc = loss_oracle.regs(a, b, 10.0)
    (r_sh + r_sc).backward()
    v_sh = torch.ones((n, 25, 3), device=dev)
    v_ls = torch.ones((n, 3), device=dev)
    loss = param_regs(sh.to(dev), ls.to(dev), 10.0, v_sh_coeffs=v_sh, v_log_scales=v_ls)
    assert loss[0].item() == pytest.approx(r_sh.item(), rel=2e-5)
    assert loss[1].item() == pytest.approx(r_sc.item(), rel=2e-5, abs=1e-9)
    assert float(((v_sh.cpu() - 1) - a.grad.float()).abs().max()) <= 2e-5 * float(a.grad.abs().max()) + 1e-7
    assert float(((v_ls.cpu() - 1) - b.grad.float()).abs().max()) <= 2e-5 * float(b.grad.abs().max()) + 1e-7
    assert float(b.grad.abs().max()) > 0 or n == 1


def test_grad_out_leaves_get_no_autograd_grad(dev):
    """Leaves whose gradient goes into caller-owned buffers (holder['grad_out']) get no `.grad`: the buffers are
    the gradients, and they equal what autograd delivers without them."""
    from gaussiangrasper_b200 import scenes
    from gaussiangrasper_b200.distributed import GradientBucket
    from gaussiangrasper_b200.render import ViewBatch, render_views
    n, W, H, D = 3000, 80, 64, 5
    sc = scenes.random_scene(n, feature_dim=D, seed=8)
    sc["log_scales"] = sc["log_scales"] + 0.8
    names = ("means", "log_scales", "quats", "opacity_logit", "sh_coeffs", "features")
    vb = ViewBatch.from_cameras(scenes.orbit_cameras(2, W, H, total=5), dev)
    v = torch.randn((2, H, W, 12), generator=torch.Generator().manual_seed(0)).to(dev)
    P = {k: sc[k].to(dev).requires_grad_(True) for k in names}
    render_views(*(P[k] for k in names), vb)["image"].backward(v)
    want = {k: P[k].grad.clone() for k in names}
    Q = {k: sc[k].to(dev).requires_grad_(True) for k in names}
    bucket = GradientBucket(Q)
    for rep in range(2):   # the second pass overwrites, it does not accumulate
        render_views(*(Q[k] for k in names), vb, holder={"grad_out": bucket.unpack()})["image"].backward(v)
        assert all(Q[k].grad is None for k in names)
        for k in names:
            ref = want[k].reshape(bucket.view(k).shape)
            assert torch.allclose(bucket.view(k), ref, rtol=1e-4, atol=1e-6 * float(ref.abs().max())), (rep, k)


@pytest.mark.parametrize("shape,strided", [((1, 32), False), ((127, 32), False), ((480, 640, 32), True), ((3, 50, 70, 32), True)])
def test_up_projection_kernel_matches_fp64(dev, shape, strided):
    """The tcgen05 / TMEM up-projection (3xTF32 split) against the same MLP in fp64: fp32-equivalent accuracy --
    the error of a plain fp32 evaluation of the two layers, not TF32's 1e-3."""
    from gaussiangrasper_b200.losses import UpProjection, up_project
    torch.manual_seed(5)
    mlp = UpProjection(32).to(dev)
    with torch.no_grad():
        for p in mlp.parameters():
            p.mul_(3.0)                       # larger activations than the default initialisation gives
    g = torch.Generator().manual_seed(sum(shape))
    if strided:                               # the feature channels of a blended image: rows 40 floats apart
        img = torch.randn(shape[:-1] + (40,), generator=g).to(dev)
        x = img[..., 7:39]
    else:
        x = torch.randn(shape, generator=g).to(dev)
    y = up_project(x, mlp)
    assert y.shape == shape[:-1] + (512,)
    ref64 = mlp.double()(x.double())
    mlp.float()
    scale = float(ref64.abs().max())
    err = float((y.double() - ref64).abs().max()) / scale
    torch.backends.cuda.matmul.allow_tf32 = False
    ref32 = mlp(x.contiguous())
    err32 = float((ref32.double() - ref64).abs().max()) / scale
    assert err <= 2e-6 + 4 * err32, (err, err32)
    assert err < 1e-5
    # repeated calls reuse the packed weights; changed weights are repacked
    with torch.no_grad():
        mlp.layers[2].bias.add_(1.0)
        mlp.layers[0].weight.mul_(0.5)
    y2 = up_project(x, mlp)
    ref2 = mlp.double()(x.double())
    mlp.float()
    assert float((y2.double() - ref2).abs().max()) / float(ref2.abs().max()) < 1e-5
