"""GPU edge cases of the path: nothing visible, a Gaussian covering the whole screen, one Gaussian,
no feature channels, far-off-screen centres -- each against the CPU oracle where there is something to
compare, and always for finite outputs / gradients."""
import numpy as np
import pytest
import torch

from gaussiangrasper_b200 import scenes
from oracle import c_oracle

pytestmark = pytest.mark.gpu
NAMES = ("means", "log_scales", "quats", "opacity_logit", "sh_coeffs", "features")


def _render(sc, cams, dev, **kw):
    from gaussiangrasper_b200.render import ViewBatch, render_views
    P = {k: sc[k].to(dev).contiguous().requires_grad_(True) for k in NAMES}
    holder = {"debug_activations": True}
    out = render_views(*(P[k] for k in NAMES), ViewBatch.from_cameras(cams, dev), holder=holder, **kw)
    return P, out, holder


def test_nothing_visible_gives_background_and_zero_gradients():
    dev = torch.device("cuda:0")
    sc = scenes.random_scene(500, feature_dim=4, seed=1)
    cam = scenes.look_at_camera((4.5, 0.3, 0.2), 64, 48)
    sc["means"] = sc["means"] * 0.01 + cam.position + (cam.position / cam.position.norm()) * 3.0  # behind the camera
    P, out, holder = _render(sc, [cam], dev)
    assert int(holder["radii"].sum()) == 0 and holder["binning"].num_intersects == 0
    img = out["image"][0]
    assert torch.all(img[..., 3] == 10.0) and torch.all(img[..., :3] == 0) and torch.all(img[..., 4:] == 0)
    assert torch.all(out["alpha"] == 0)
    out["image"].sum().backward()
    for k in NAMES:
        assert float(P[k].grad.abs().max()) == 0.0, k


def test_screen_filling_gaussian_and_single_gaussian():
    dev = torch.device("cuda:0")
    W, H = 200, 120
    cam = scenes.look_at_camera((4.5, 0.0, 0.0), W, H)
    for n in (1, 3):
        sc = scenes.random_scene(n, feature_dim=2, seed=2)
        sc["means"][:] = 0.0
        sc["means"][:, 0] = torch.linspace(0.0, 1.0, n)          # in front of the camera, different depths
        sc["log_scales"][:] = np.log(3.0)                        # metres wide: covers every tile
        sc["opacity_logit"][:] = 1.0
        P, out, holder = _render(sc, [cam], dev)
        tb = cam.tile_bounds
        assert int(holder["num_tiles_hit"].max()) == tb[0] * tb[1]
        ref = c_oracle.project_fwd(sc["means"].numpy(), holder["scales"].cpu().numpy(), 1.0, holder["quats"].cpu().numpy(),
                                   cam.viewmat[:3].numpy(), cam.fullmat.numpy(), cam.fx, cam.fy, cam.cx, cam.cy, H, W, tb)
        assert np.array_equal(holder["radii"][0].cpu().numpy(), ref[2])
        _, _, _, _, ids_s, ranges = c_oracle.bin_and_sort(ref[0], ref[1], ref[2], ref[4], tb)
        assert np.array_equal(holder["binning"].ids_sorted.cpu().numpy(), ids_s)
        assert np.array_equal(holder["binning"].tile_ranges.cpu().numpy(), ranges)
        assert float(out["alpha"].min()) > 0.1          # every pixel is covered
        out["image"].sum().backward()
        for k in NAMES:
            assert torch.isfinite(P[k].grad).all(), k
        assert float(P["opacity_logit"].grad.abs().max()) > 0


def test_no_feature_channels_and_far_offscreen_centres():
    dev = torch.device("cuda:0")
    sc = scenes.random_scene(3000, feature_dim=0, seed=3)
    sc["log_scales"] = sc["log_scales"] + 1.0
    cam = scenes.look_at_camera((4.5, 0.3, 0.2), 96, 64)
    sc["means"][:200, 1] += 3.0e4        # projected millions of pixels away: culled, no overflow trouble
    sc["means"][200:400] = cam.position + torch.randn(200, 3) * 0.02   # around the camera centre / near plane
    P, out, holder = _render(sc, [cam], dev)
    assert out["feature"].shape[-1] == 0 and out["image"].shape[-1] == 8
    assert torch.isfinite(out["image"]).all()
    out["image"].sum().backward()
    for k in NAMES:
        assert P[k].grad is not None and torch.isfinite(P[k].grad).all(), k
    # integer outputs still bit-exact against the oracle
    ref = c_oracle.project_fwd(sc["means"].numpy(), holder["scales"].cpu().numpy(), 1.0, holder["quats"].cpu().numpy(),
                               cam.viewmat[:3].numpy(), cam.fullmat.numpy(), cam.fx, cam.fy, cam.cx, cam.cy, 64, 96,
                               cam.tile_bounds)
    assert np.array_equal(holder["radii"][0].cpu().numpy(), ref[2])
    assert np.array_equal(holder["num_tiles_hit"][0].cpu().numpy(), ref[4])


def test_lower_sh_degrees_in_the_fused_path():
    """degrees_to_use < 4 (the model ramps it up every 1000 steps, gaussian_splatting.py:729)."""
    dev = torch.device("cuda:0")
    sc = scenes.random_scene(2000, feature_dim=3, seed=4)
    sc["log_scales"] = sc["log_scales"] + 1.0
    cam = scenes.look_at_camera((4.5, 0.3, 0.2), 80, 64)
    for deg in (0, 2):
        P, out, holder = _render(sc, [cam], dev, degrees_to_use=deg)
        dirs = sc["means"].numpy() - cam.position.numpy()
        rgb = np.clip(c_oracle.sh_fwd(deg, dirs, sc["sh_coeffs"].numpy()) + 0.5, 0, 1)
        out["rgb"].sum().backward()
        g = P["sh_coeffs"].grad
        nb = (deg + 1) ** 2
        assert float(g[:, nb:].abs().max()) == 0.0 and float(g[:, :nb].abs().max()) > 0.0
        # colour of an isolated check: re-blend with the oracle
        ref = c_oracle.project_fwd(sc["means"].numpy(), holder["scales"].cpu().numpy(), 1.0, holder["quats"].cpu().numpy(),
                                   cam.viewmat[:3].numpy(), cam.fullmat.numpy(), cam.fx, cam.fy, cam.cx, cam.cy, 64, 80,
                                   cam.tile_bounds)
        _, _, _, _, ids_s, ranges = c_oracle.bin_and_sort(ref[0], ref[1], ref[2], ref[4], cam.tile_bounds)
        opac = torch.sigmoid(sc["opacity_logit"]).numpy()
        img, _, _, frag, _ = c_oracle.blend_fwd(64, 80, cam.tile_bounds, ids_s, ranges, ref[0], ref[3], opac,
                                                rgb.astype(np.float32), np.zeros(3, np.float32), eps=2e-5)
        err = np.abs(out["rgb"][0].detach().cpu().numpy() - img)
        assert err[~frag].max() <= 1e-4
