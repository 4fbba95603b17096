"""CPU tests: the C oracle, the torch oracle and the host build of the device math must agree.

PARITY UNPINNED: there is no gsplat 0.1.0 binary or golden vector to pin these against
(SURVEY.md 8c); the two independently written restatements plus autograd are cross-checked here.
"""
import ctypes as C
import math

import numpy as np
import pytest
import torch

from gaussiangrasper_b200 import scenes
from oracle import c_oracle, torch_oracle


def small_scene(n=3000, W=96, H=64, seed=7, D=5, big=False):
    sc = scenes.random_scene(n, feature_dim=D, seed=seed)
    if big:
        sc["log_scales"] = sc["log_scales"] + 1.0
    cam = scenes.look_at_camera((4.5, 0.3, 0.2), W, H)
    return sc, cam


def project_c(sc, cam, clip=0.01):
    return c_oracle.project_fwd(sc["means"].numpy(), sc["log_scales"].exp().numpy(), 1.0,
                                (sc["quats"] / sc["quats"].norm(dim=-1, keepdim=True)).numpy(),
                                cam.viewmat[:3].numpy(), cam.fullmat.numpy(), cam.fx, cam.fy, cam.cx, cam.cy,
                                cam.H, cam.W, cam.tile_bounds, clip)


def project_t(sc, cam, dtype=torch.float32, clip=0.01):
    q = sc["quats"] / sc["quats"].norm(dim=-1, keepdim=True)
    return torch_oracle.project_gaussians(sc["means"].to(dtype), sc["log_scales"].exp().to(dtype), 1.0, q.to(dtype),
                                          cam.viewmat, cam.fullmat, cam.fx, cam.fy, cam.cx, cam.cy, cam.H, cam.W,
                                          cam.tile_bounds, clip)


def test_projection_c_vs_torch():
    sc, cam = small_scene(20000, 640, 480)
    xys, depths, radii, conics, nth, cov3d = project_c(sc, cam)
    t = project_t(sc, cam, torch.float64)
    vis = radii > 0
    assert vis.sum() > 1000
    # integer outputs: the fp64 evaluation may round a radius / bbox differently on a few Gaussians
    same = (t[2].numpy() == radii) & (t[4].numpy() == nth)
    assert same.mean() > 0.999
    m = same & vis
    np.testing.assert_allclose(xys[m], t[0].numpy()[m], rtol=1e-4, atol=2e-3)
    np.testing.assert_allclose(depths[m], t[1].numpy()[m], rtol=1e-5)
    np.testing.assert_allclose(conics[m], t[3].numpy()[m], rtol=2e-3, atol=1e-5)
    np.testing.assert_allclose(cov3d[vis], t[5].numpy()[vis], rtol=1e-4, atol=1e-9)
    # culled rows are all-zero
    assert not xys[~vis].any() and not conics[~vis].any() and not depths[~vis].any() and not nth[~vis].any()


def test_hostmath_projection_bit_exact(hostmath):
    """The device math (host build) follows the oracle's operation order exactly."""
    sc, cam = small_scene(20000, 640, 480, seed=11)
    ref = project_c(sc, cam)
    n = sc["means"].shape[0]
    means = sc["means"].numpy()
    scales = sc["log_scales"].exp().numpy()
    quats = (sc["quats"] / sc["quats"].norm(dim=-1, keepdim=True)).numpy()
    vm = np.ascontiguousarray(cam.viewmat[:3].numpy().reshape(-1))
    fm = np.ascontiguousarray(cam.fullmat.numpy().reshape(-1))
    cov3d = np.empty((n, 6), np.float32); xys = np.empty((n, 2), np.float32); depths = np.empty(n, np.float32)
    radii = np.empty(n, np.int32); conics = np.empty((n, 3), np.float32); nth = np.empty(n, np.int32)
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    hostmath.hm_project_fwd(C.c_int(n), p(means), p(scales), C.c_float(1.0), p(quats), p(vm), p(fm),
                            C.c_float(cam.fx), C.c_float(cam.fy), C.c_float(cam.cx), C.c_float(cam.cy),
                            C.c_int(cam.H), C.c_int(cam.W), C.c_int(cam.tile_bounds[0]), C.c_int(cam.tile_bounds[1]),
                            C.c_float(0.01), p(cov3d), p(xys), p(depths), p(radii), p(conics), p(nth))
    for a, b in zip((xys, depths, radii, conics, nth, cov3d), ref):
        assert a.tobytes() == b.tobytes()


def test_hostmath_projection_backward_vs_autograd(hostmath):
    sc, cam = small_scene(4000, 640, 480, seed=3)
    q = sc["quats"]  # un-normalised on purpose: the kernel normalises and must chain through it
    means = sc["means"].double().requires_grad_(True)
    scales = sc["log_scales"].exp().double().requires_grad_(True)
    quats = q.double().requires_grad_(True)
    out = torch_oracle.project_gaussians(means, scales, 1.3, quats, cam.viewmat, cam.fullmat, cam.fx, cam.fy, cam.cx,
                                         cam.cy, cam.H, cam.W, cam.tile_bounds)
    g = torch.Generator().manual_seed(0)
    v_xys = torch.randn(out[0].shape, generator=g, dtype=torch.float64)
    v_dep = torch.randn(out[1].shape, generator=g, dtype=torch.float64)
    v_con = torch.randn(out[3].shape, generator=g, dtype=torch.float64)
    torch.autograd.backward([out[0], out[1], out[3]], [v_xys, v_dep, v_con])

    n = means.shape[0]
    f = lambda t: np.ascontiguousarray(t.detach().float().numpy())
    # forward in fp32 through the same host math to get radii/conics as the kernel would see them
    vm = f(cam.viewmat[:3].reshape(-1)); fm = f(cam.fullmat.reshape(-1))
    cov3d = np.empty((n, 6), np.float32); xys = np.empty((n, 2), np.float32); depths = np.empty(n, np.float32)
    radii = np.empty(n, np.int32); conics = np.empty((n, 3), np.float32); nth = np.empty(n, np.int32)
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    m32, s32, q32 = f(means), f(scales), f(quats)
    args = (C.c_float(cam.fx), C.c_float(cam.fy), C.c_float(cam.cx), C.c_float(cam.cy), C.c_int(cam.H), C.c_int(cam.W))
    hostmath.hm_project_fwd(C.c_int(n), p(m32), p(s32), C.c_float(1.3), p(q32), p(vm), p(fm), *args,
                            C.c_int(cam.tile_bounds[0]), C.c_int(cam.tile_bounds[1]), C.c_float(0.01),
                            p(cov3d), p(xys), p(depths), p(radii), p(conics), p(nth))
    vm_, vs_, vq_ = np.empty((n, 3), np.float32), np.empty((n, 3), np.float32), np.empty((n, 4), np.float32)
    vx32, vd32, vc32 = f(v_xys), f(v_dep), f(v_con)
    hostmath.hm_project_bwd(C.c_int(n), p(m32), p(s32), C.c_float(1.3), p(q32), p(vm), p(fm), *args,
                            p(radii), p(conics), p(vx32), p(vd32), p(vc32), p(vm_), p(vs_), p(vq_))
    same = (out[2].numpy() > 0) == (radii > 0)
    assert same.mean() > 0.999
    for got, ref in ((vm_, means.grad), (vs_, scales.grad), (vq_, quats.grad)):
        ref = ref.numpy()[same]
        got = got[same]
        scale = np.abs(ref).max(axis=1, keepdims=True) + 1e-12
        # fp32 evaluation against fp64 autograd: 1e-3 relative to the row's largest component
        bad = np.abs(got - ref) > 2e-3 * scale + 1e-6 * np.abs(ref).max()
        assert bad.mean() < 1e-3, bad.mean()


def test_sh_c_vs_torch_and_hostmath(hostmath):
    g = torch.Generator().manual_seed(5)
    n = 500
    dirs = torch.randn((n, 3), generator=g)
    coeffs = torch.randn((n, 25, 3), generator=g)
    for deg in range(5):
        ref = torch_oracle.spherical_harmonics(deg, dirs.double(), coeffs.double()).numpy()
        got = c_oracle.sh_fwd(deg, dirs.numpy(), coeffs.numpy())
        np.testing.assert_allclose(got, ref, rtol=1e-4, atol=1e-5)
        v = torch.randn((n, 3), generator=g)
        vb = c_oracle.sh_bwd(4, deg, dirs.numpy(), v.numpy())
        c = coeffs.double().clone().requires_grad_(True)
        torch_oracle.spherical_harmonics(deg, dirs.double(), c).backward(v.double())
        np.testing.assert_allclose(vb, c.grad.numpy(), rtol=1e-4, atol=1e-6)
    Y = np.zeros((n, 25), np.float32)
    d = np.ascontiguousarray(dirs.numpy())
    hostmath.hm_sh_basis(C.c_int(n), C.c_int(4), d.ctypes.data_as(C.c_void_p), Y.ctypes.data_as(C.c_void_p))
    np.testing.assert_allclose(Y, torch_oracle.sh_basis(4, dirs.double()).numpy(), rtol=1e-4, atol=1e-6)
    assert [c_oracle.num_sh_bases(k) for k in range(6)] == [1, 4, 9, 16, 25, 25]


def test_binning_c_vs_torch():
    sc, cam = small_scene(20000, 640, 480, seed=9, big=True)
    xys, depths, radii, conics, nth, _ = project_c(sc, cam)
    cum, keys, ids, keys_s, ids_s, ranges = c_oracle.bin_and_sort(xys, depths, radii, nth, cam.tile_bounds)
    t = torch_oracle.bin_and_sort(torch.from_numpy(xys), torch.from_numpy(depths), torch.from_numpy(radii),
                                  torch.from_numpy(nth), cam.tile_bounds)
    assert (t[0].numpy() == cum).all()
    assert (t[1].numpy() == keys_s).all()
    assert (t[2].numpy() == ids_s).all()
    assert (t[3].numpy() == ranges).all()
    m = int(cum[-1])
    assert m == len(keys_s) and m > 20000
    # sortedness and range consistency
    assert (np.diff(keys_s) >= 0).all()
    tiles = keys_s >> 32
    for tid in np.unique(tiles)[:50]:
        s, e = ranges[tid]
        assert (tiles[s:e] == tid).all() and (s == 0 or tiles[s - 1] != tid) and (e == m or tiles[e] != tid)


def test_binning_empty_and_single():
    # nothing visible
    n = 10
    xys = np.zeros((n, 2), np.float32); depths = np.zeros(n, np.float32)
    radii = np.zeros(n, np.int32); nth = np.zeros(n, np.int32)
    cum, keys, ids, keys_s, ids_s, ranges = c_oracle.bin_and_sort(xys, depths, radii, nth, (4, 3, 1))
    assert len(keys_s) == 0 and not ranges.any()
    # one Gaussian alone in the last tile (SURVEY App. B-5 edge case)
    xys[3] = (60.0, 40.0); depths[3] = 2.0; radii[3] = 1; nth[3] = 1
    cum, keys, ids, keys_s, ids_s, ranges = c_oracle.bin_and_sort(xys, depths, radii, nth, (4, 3, 1))
    assert len(keys_s) == 1 and ids_s[0] == 3
    tile = 2 * 4 + 3
    assert tuple(ranges[tile]) == (0, 1) and ranges.sum() == 1


@pytest.mark.parametrize("channels", [1, 3, 7])
def test_blend_c_loop_vs_torch_vectorised(channels):
    sc, cam = small_scene(4000, 96, 64, seed=21, big=True)
    xys, depths, radii, conics, nth, _ = project_c(sc, cam)
    _, _, _, _, ids_s, ranges = c_oracle.bin_and_sort(xys, depths, radii, nth, cam.tile_bounds)
    g = torch.Generator().manual_seed(1)
    colors = torch.rand((xys.shape[0], channels), generator=g)
    opac = torch.sigmoid(sc["opacity_logit"]).reshape(-1)
    bg = torch.rand(channels, generator=g)
    out, fT, fi, frag, pairs = c_oracle.blend_fwd(cam.H, cam.W, cam.tile_bounds, ids_s, ranges, xys, conics,
                                                  opac.numpy(), colors.numpy(), bg.numpy(), eps=2e-5)
    to, tT, ti = torch_oracle.rasterize(torch.from_numpy(xys), torch.from_numpy(conics), opac, colors,
                                        torch.from_numpy(ids_s), torch.from_numpy(ranges), cam.H, cam.W, bg)
    ok = ~frag
    assert ok.mean() > 0.97
    assert pairs > 0
    np.testing.assert_allclose(out[ok], to.numpy()[ok], atol=2e-5, rtol=0)
    np.testing.assert_allclose(fT[ok], tT.numpy()[ok], atol=1e-5, rtol=0)
    assert (fi[ok] == ti.numpy()[ok]).mean() > 0.999
    # a covered image: some pixels must be far from background
    assert (fT < 0.5).mean() > 0.05


def test_blend_backward_c_vs_autograd():
    sc, cam = small_scene(1500, 64, 48, seed=33, big=True)
    xys, depths, radii, conics, nth, _ = project_c(sc, cam)
    _, _, _, _, ids_s, ranges = c_oracle.bin_and_sort(xys, depths, radii, nth, cam.tile_bounds)
    g = torch.Generator().manual_seed(2)
    ch = 4
    colors = torch.rand((xys.shape[0], ch), generator=g)
    opac = torch.sigmoid(sc["opacity_logit"]).reshape(-1)
    bg = torch.rand(ch, generator=g)
    v_out = torch.randn((cam.H, cam.W, ch), generator=g)
    v_xy, v_conic, v_colors, v_opac = c_oracle.blend_bwd(cam.H, cam.W, cam.tile_bounds, ids_s, ranges, xys, conics,
                                                         opac.numpy(), colors.numpy(), bg.numpy(), v_out.numpy())
    _, gx, gc, go, gcol = torch_oracle.rasterize_grads(
        torch.from_numpy(xys).double(), torch.from_numpy(conics).double(), opac.double(), colors.double(),
        torch.from_numpy(ids_s), torch.from_numpy(ranges), cam.H, cam.W, bg.double(), v_out.double())
    for got, ref in ((v_xy, gx), (v_conic, gc), (v_opac, go), (v_colors, gcol)):
        ref = ref.numpy()
        denom = np.abs(ref).max() + 1e-30
        # both are fp64 on fp32 inputs; the contributing sets can differ on threshold pairs only
        err = np.abs(got - ref) / denom
        assert np.quantile(err, 0.999) < 1e-6, np.quantile(err, 0.999)
        assert err.max() < 5e-2


def test_quat_to_rotmat_matches_reference_convention():
    from gaussiangrasper_b200 import quat_to_rotmat
    q = torch.tensor([[1.0, 0, 0, 0], [math.cos(0.3), math.sin(0.3), 0, 0], [2.0, 0.0, 0.0, 2.0]])
    R = quat_to_rotmat(q)
    assert torch.allclose(R[0], torch.eye(3), atol=1e-7)
    # rotation by 0.6 rad about x
    c, s = math.cos(0.6), math.sin(0.6)
    assert torch.allclose(R[1], torch.tensor([[1, 0, 0], [0, c, -s], [0, s, c]], dtype=torch.float32), atol=1e-6)
    # un-normalised input is normalised: 90 degrees about z
    assert torch.allclose(R[2], torch.tensor([[0.0, -1, 0], [1, 0, 0], [0, 0, 1]]), atol=1e-6)
    assert torch.allclose(R, torch_oracle.quat_to_rotmat(q), atol=1e-6)


def test_literal_loop_equals_the_vectorised_blend_and_the_c_oracle():
    """the literal per-pixel loop on a small crop validates the vectorised torch blend (and the
    C oracle's loop) -- three restatements of A9, one result."""
    n, W, H = 1500, 64, 48
    sc = scenes.random_scene(n, feature_dim=2, seed=3)
    sc["log_scales"] = sc["log_scales"] + 0.8
    cam = scenes.look_at_camera((4.5, 0.3, 0.2), W, H)
    q = sc["quats"] / sc["quats"].norm(dim=-1, keepdim=True)
    xys, depths, radii, conics, nth, _ = torch_oracle.project_gaussians(
        sc["means"], sc["log_scales"].exp(), 1.0, q, cam.viewmat, cam.fullmat, cam.fx, cam.fy, cam.cx, cam.cy, H, W,
        cam.tile_bounds)
    _, _, ids_s, ranges = torch_oracle.bin_and_sort(xys, depths, radii, nth, cam.tile_bounds)
    g = torch.Generator().manual_seed(0)
    cols = torch.rand((n, 4), generator=g)
    bg = torch.tensor([0.0, 0.0, 0.0, 10.0])
    op = torch.sigmoid(sc["opacity_logit"]).reshape(-1)
    crop = (16, 8, 32, 24)
    lit, pairs = torch_oracle.rasterize_literal(xys, conics, op, cols, ids_s, ranges, H, W, bg, crop)
    vec, _, _ = torch_oracle.rasterize(xys.double(), conics.double(), op.double(), cols.double(), ids_s, ranges, H, W, bg)
    assert pairs > 10_000
    assert torch.allclose(lit, vec[8:32, 16:48], rtol=0, atol=1e-9)
    cout, _, _, frag, _ = c_oracle.blend_fwd(H, W, cam.tile_bounds, ids_s.numpy(), ranges.numpy(), xys.numpy(), conics.numpy(),
                                              op.numpy(), cols.numpy(), bg.numpy(), eps=2e-5)
    ok = ~torch.from_numpy(frag[8:32, 16:48])
    assert float((lit.float() - torch.from_numpy(cout[8:32, 16:48]))[ok].abs().max()) < 2e-5
