"""world_size-2 gloo tests of the view-sharding host logic (no GPU): partition of views and the
single flat gradient all-reduce."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from gaussiangrasper_b200 import distributed as ggd


def test_views_partition_is_a_partition():
    for world in (1, 2, 4, 8):
        for n in (1, 7, 8, 64):
            seen = []
            for r in range(world):
                mine = ggd.views_for_rank(n, r, world)
                assert all(ggd.owner_of_view(v, world) == r for v in mine)
                seen += mine
            assert sorted(seen) == list(range(n))
    with pytest.raises(ValueError):
        ggd.views_for_rank(4, 2, 2)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _make_params(n=50, D=4, K=25):
    g = torch.Generator().manual_seed(0)
    shapes = dict(means=(n, 3), log_scales=(n, 3), quats=(n, 4), opacity_logit=(n, 1), sh_coeffs=(n, K, 3),
                  features=(n, D))
    return {k: torch.randn(s, generator=g).requires_grad_(True) for k, s in shapes.items()}


def _per_view_grad(params, view):
    """A stand-in for "gradient of the loss of one view": deterministic function of (param, view)."""
    return {k: torch.sin(p.detach() * (view + 1)) * (0.5 + view) for k, p in params.items()}


def _worker(rank, world, port, n_views, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        params = _make_params()
        # accumulate the local views' gradients into .grad, as a training step would
        for k in params:
            params[k].grad = torch.zeros_like(params[k])
        for v in ggd.views_for_rank(n_views, rank, world):
            for k, g in _per_view_grad(params, v).items():
                params[k].grad += g
        bucket = ggd.all_reduce_gradients(params)
        assert bucket.payload == sum(p.numel() for p in params.values()) <= bucket.numel
        # a second step reuses the persistent flat buffer
        ggd.all_reduce_gradients({k: p for k, p in params.items()}, bucket)
        if rank == 0:
            torch.save({k: p.grad.clone() for k, p in params.items()}, out)
    finally:
        dist.destroy_process_group()


def test_gradient_allreduce_two_ranks_equals_sum_over_views(tmp_path):
    world, n_views = 2, 5
    out = str(tmp_path / "grads.pt")
    mp.spawn(_worker, args=(world, _free_port(), n_views, out), nprocs=world, join=True)
    got = torch.load(out)
    params = _make_params()
    once = {k: sum(_per_view_grad(params, v)[k] for v in range(n_views)) for k in params}
    for k in params:
        # the worker all-reduces twice (the second call sums the already-reduced grads of both ranks)
        assert torch.allclose(got[k], once[k] * world, rtol=1e-5, atol=1e-6), k


def test_bucket_layout_single_process():
    params = _make_params(n=7, D=3, K=4)
    b = ggd.GradientBucket(params)
    assert b.payload == 7 * (3 + 3 + 4 + 1 + 12 + 3) <= b.numel < b.payload + 4 * 6
    assert all(o % 4 == 0 for o in b.offsets.values())
    grads = {k: torch.full_like(p, i + 1.0) for i, (k, p) in enumerate(params.items())}
    grads["quats"] = None
    flat = b.pack(grads)
    assert b.all_reduce() is None  # no process group: no-op
    un = b.unpack()
    assert float(un["means"].mean()) == 1.0 and float(un["quats"].abs().sum()) == 0.0
    assert un["sh_coeffs"].shape == (7, 4, 3) and flat.numel() == b.numel
    with pytest.raises(ValueError):
        ggd.GradientBucket({"bogus": torch.zeros(3)})


# ---------------------------------------------------------------------------------------------
# FactoredExchange: SH gradient exchanged as per-view factors (all-gather) + all-reduce of the rest.
# The CUDA rebuild kernel is replaced by the oracle's SH basis here (no GPU in this suite).
# ---------------------------------------------------------------------------------------------
def _torch_sh_rebuild(degree, degrees_to_use, means, positions, v_rgb_views, out=None):
    from oracle import torch_oracle
    vt, n = positions.shape[0], means.shape[0]
    nb = (degree + 1) ** 2
    acc = torch.zeros((n, nb, 3))
    for v in range(vt):
        dirs = means - positions[v][None]
        Y = torch_oracle.sh_basis(degrees_to_use, dirs / dirs.norm(dim=-1, keepdim=True))   # [n, nuse]
        acc[:, :Y.shape[1]] += Y[:, :, None] * v_rgb_views.view(vt, n, 3)[v][:, None, :]
    out.copy_(acc)
    return out


def _view_factor(n, view):
    g = torch.Generator().manual_seed(1000 + view)
    rgb = torch.randn((n, 3), generator=g)
    rgb[torch.rand(n, generator=g) < 0.3] = 0.0   # Gaussians this view does not see
    pos = torch.randn(3, generator=g) * 3.0
    return rgb, pos


def _factored_worker(rank, world, port, views_per_rank, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        params = _make_params(n=40, D=4, K=25)
        ex = ggd.FactoredExchange(params, views_per_rank, reconstruct=_torch_sh_rebuild)
        holder = ex.holder()
        assert holder["defer_sh_grad"] and "sh_coeffs" not in holder["grad_out"]
        # what the backward of this rank's render_views call would write
        mine = [rank * views_per_rank + j for j in range(views_per_rank)]
        for k, buf in holder["grad_out"].items():
            if k != "v_rgb_views":
                buf.copy_(sum(_per_view_grad(params, v)[k] for v in mine))
        pos = torch.stack([_view_factor(40, v)[1] for v in mine])
        holder["grad_out"]["v_rgb_views"].copy_(torch.stack([_view_factor(40, v)[0] for v in mine]))
        grads = ex.exchange(params["means"], pos, 4, 4, holder)
        if rank == 0:
            torch.save({k: g.clone() for k, g in grads.items()}, out)
    finally:
        dist.destroy_process_group()


def test_factored_exchange_two_ranks(tmp_path):
    world, vpr = 2, 2
    out = str(tmp_path / "fgrads.pt")
    mp.spawn(_factored_worker, args=(world, _free_port(), vpr, out), nprocs=world, join=True)
    got = torch.load(out)
    params = _make_params(n=40, D=4, K=25)
    views = list(range(world * vpr))
    for k in params:
        if k == "sh_coeffs":
            continue
        want = sum(_per_view_grad(params, v)[k] for v in views)
        assert torch.allclose(got[k], want, rtol=1e-5, atol=1e-6), k
    rgb = torch.stack([_view_factor(40, v)[0] for v in views])
    pos = torch.stack([_view_factor(40, v)[1] for v in views])
    want_sh = _torch_sh_rebuild(4, 4, params["means"].detach(), pos, rgb, out=torch.zeros(40, 25, 3))
    assert torch.allclose(got["sh_coeffs"], want_sh, rtol=1e-5, atol=1e-6)


def test_factored_exchange_single_process_is_local():
    params = _make_params(n=9, D=2, K=4)
    ex = ggd.FactoredExchange(params, 3, reconstruct=_torch_sh_rebuild)
    h = ex.holder()
    assert h["grad_out"]["v_rgb_views"].shape == (3, 9, 3) and ex.bucket.payload == 9 * (3 + 3 + 4 + 1 + 2)
    h["grad_out"]["v_rgb_views"].copy_(torch.ones(3, 9, 3))
    g = ex.exchange(params["means"], torch.zeros(3, 3) + torch.arange(3.0)[:, None] + 5.0, 1, 1, h)
    assert g["sh_coeffs"].shape == (9, 4, 3) and torch.isfinite(g["sh_coeffs"]).all()
    with pytest.raises(ValueError):
        ex.exchange(params["means"], torch.zeros(2, 3), 1, 1, h)


def test_factorisation_equals_the_oracle_sh_backward():
    """The identity FactoredExchange relies on: the SH backward of the C oracle (gsplat's compute_sh_backward
    restated) for one view is the outer product Y(dir) (x) v_rgb, and gradients of several views add."""
    import numpy as np
    from oracle import c_oracle
    g = torch.Generator().manual_seed(3)
    n = 64
    means = torch.randn(n, 3, generator=g)
    pos = torch.randn(3, 3, generator=g) * 4
    rgb = torch.randn(3, n, 3, generator=g)
    rgb[1, ::5] = 0.0
    for deg_use in (4, 2, 0):
        want = np.zeros((n, 25, 3), np.float64)
        for v in range(3):
            dirs = (means - pos[v]).numpy()
            want += c_oracle.sh_bwd(4, deg_use, dirs, rgb[v].numpy())
        got = _torch_sh_rebuild(4, deg_use, means, pos, rgb, out=torch.zeros(n, 25, 3))
        np.testing.assert_allclose(got.numpy(), want, rtol=2e-5, atol=2e-6)
        assert not got[:, (deg_use + 1) ** 2:].any()
