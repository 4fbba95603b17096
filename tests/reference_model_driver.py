"""Drive the reference's OWN model file through this repository's gsplat shim (run as a script by
tests/test_reference_model_cpu.py; needs /root/reference, so it only runs where the reference is mounted).

    nerfstudio/models/gaussian_splatting.py  --imports-->  gsplat.{project_gaussians,rasterize,nd_rasterize,sh,_torch_impl}
                                                             = this repository's autograd Functions (gsplat/ shim)

`GaussianSplattingModel.get_outputs` (:624-802), `get_loss_dict` (:841-933), `loss.backward()` and `after_train`
(:373-393) run unmodified.  There is no GPU here and the product has no CPU fallback, so -- in THIS TEST ONLY -- the
C-ABI calls underneath the autograd Functions (gaussiangrasper_b200.ops.*) are replaced by the CPU oracle: what is
exercised is the boundary itself -- import paths, positional argument orders, shapes / dtypes, error behaviour and
the autograd obligations of SURVEY 8b (xys.retain_grad(), depths carrying gradient, the in-place write into the
returned rgb before backward).

The reference targets Python 3.8-3.10; two workarounds keep its files importable on 3.12 without touching them:
its dataclasses use dataclass instances as field defaults (accepted again once the classes are hashable), and the
packages it imports but this path never calls (viser, pytorch_msssim, torchmetrics, comet_ml, nerfacc) are stubs.
"""
import dataclasses
import json
import math
import os
import sys
import types

import numpy as np
import torch

import torch.fx  # noqa: F401  (imported before dataclasses is patched: torch's own dataclasses must not be touched)

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"


def install_stubs():
    _orig = dataclasses.dataclass

    def _patched(cls=None, **kw):
        def wrap(c):
            k2 = dict(kw)
            if c.__module__.startswith("nerfstudio") and not k2.get("frozen") and k2.get("eq", True):
                k2.setdefault("unsafe_hash", True)
            return _orig(c, **k2)
        return wrap if cls is None else wrap(cls)
    dataclasses.dataclass = _patched

    class _Any:
        def __init__(self, *a, **k):
            pass

        def __call__(self, *a, **k):
            return _Any()

        def __getattr__(self, k):
            if k.startswith("__"):
                raise AttributeError(k)
            return _Any()

        def __getitem__(self, k):
            return _Any()

        def __mro_entries__(self, bases):
            return (object,)

    def stub(name, **attrs):
        m = types.ModuleType(name)
        m.__path__ = []
        m.__file__ = "/nonexistent/" + name + ".py"

        def ga(k):
            if k.startswith("__"):
                raise AttributeError(k)
            return _Any()
        m.__getattr__ = ga
        for k, v in attrs.items():
            setattr(m, k, v)
        sys.modules[name] = m
        return m

    # viser.transforms.SO3.from_x_radians(pi).as_matrix() is the one thing the path needs from viser (:664)
    class SO3:
        def __init__(self, m):
            self.m = m

        @staticmethod
        def from_x_radians(t):
            c, s = math.cos(t), math.sin(t)
            return SO3(np.array([[1.0, 0.0, 0.0], [0.0, c, -s], [0.0, s, c]]))

        def as_matrix(self):
            return self.m
    stub("viser").transforms = stub("viser.transforms", SO3=SO3)

    # pytorch_msssim.SSIM(data_range=1.0, size_average=True, channel=3): the restatement of oracle/loss_oracle.py
    from oracle import loss_oracle

    class SSIM(torch.nn.Module):
        def __init__(self, data_range=1.0, size_average=True, channel=3):
            super().__init__()
            self.data_range = data_range

        def forward(self, x, y):
            return loss_oracle.ssim(x, y, self.data_range)
    stub("pytorch_msssim", SSIM=SSIM)

    class _Metric(torch.nn.Module):
        def __init__(self, *a, **k):
            super().__init__()

        def forward(self, *a, **k):
            return torch.zeros(())
    tm = stub("torchmetrics")
    tm.image = stub("torchmetrics.image", PeakSignalNoiseRatio=_Metric)
    tm.image.lpip = stub("torchmetrics.image.lpip", LearnedPerceptualImagePatchSimilarity=_Metric)
    stub("comet_ml")
    stub("nerfacc")


def import_with_stubs(module_name: str, max_stubs: int = 40):
    """Import a reference module whose import chain pulls in third-party packages that are not installed (open3d,
    imageio, rawpy, matplotlib, ...): every package a ModuleNotFoundError names is replaced by an empty stand-in and
    the import retried.  Only code paths that never touch those packages are usable afterwards."""
    import importlib
    made = []
    for _ in range(max_stubs):
        try:
            return importlib.import_module(module_name), made
        except ModuleNotFoundError as e:
            if e.name is None or e.name.startswith("nerfstudio"):
                raise
            m = types.ModuleType(e.name)
            m.__path__ = []
            m.__file__ = "/nonexistent/" + e.name

            class _Missing:
                def __init__(self, *a, **k):
                    pass

                def __call__(self, *a, **k):
                    return _Missing()

                def __getattr__(self, k):
                    if k.startswith("__"):
                        raise AttributeError(k)
                    return _Missing()

                def __getitem__(self, k):
                    return _Missing()

                def __mro_entries__(self, bases):
                    return (object,)

            def ga(k, _M=_Missing):
                if k.startswith("__"):
                    raise AttributeError(k)
                return _M()
            m.__getattr__ = ga
            sys.modules[e.name] = m
            made.append(e.name)
            for k in [k for k in sys.modules if k.startswith("nerfstudio.")]:      # half-imported modules
                spec = getattr(sys.modules[k], "__spec__", None)
                if spec is not None and getattr(spec, "_initializing", False):
                    del sys.modules[k]
    raise ImportError(f"{module_name}: more than {max_stubs} missing packages ({made})")


def install_oracle_backend():
    """Replace the C-ABI calls under the autograd Functions by the CPU oracle (this test only)."""
    from gaussiangrasper_b200 import _raster, ops
    from oracle import c_oracle, torch_oracle
    calls = []
    t = torch.from_numpy

    def project_fwd(means3d, scales, glob_scale, quats, viewmat, fullmat, fx, fy, cx, cy, H, W, tile_bounds, clip_thresh=0.01):
        calls.append("project_fwd")
        assert means3d.dtype == torch.float32 and viewmat.shape[-2:] == (3, 4) and fullmat.shape == (4, 4)
        assert isinstance(fx, float) and isinstance(H, int) and len(tile_bounds) == 3
        o = c_oracle.project_fwd(means3d.detach().numpy(), scales.detach().numpy(), glob_scale, quats.detach().numpy(),
                                 viewmat.detach().numpy(), fullmat.detach().numpy(), fx, fy, cx, cy, H, W, tile_bounds, clip_thresh)
        return tuple(t(a) for a in o)

    def project_bwd(means3d, scales, glob_scale, quats, viewmat, fullmat, fx, fy, cx, cy, H, W, radii, conics, v_xys,
                    v_depths, v_conics):
        calls.append("project_bwd")
        m, s, q = (x.detach().double().requires_grad_(True) for x in (means3d, scales, quats))
        vm = torch.eye(4, dtype=torch.float64)
        vm[:3] = viewmat.double()
        with torch.enable_grad():
            xys, depths, _, con, _, _ = torch_oracle.project_gaussians(m, s, glob_scale, q, vm, fullmat.double(), fx, fy, cx,
                                                                       cy, H, W, ((W + 15) // 16, (H + 15) // 16, 1))
            outs, cots = [xys, con], [v_xys.double(), v_conics.double()]
            if v_depths is not None:
                outs.append(depths)
                cots.append(v_depths.double())
            g = torch.autograd.grad(outs, [m, s, q], cots, allow_unused=True)
        return tuple((torch.zeros_like(x) if gi is None else gi).float() for gi, x in zip(g, (m, s, q)))

    def sh_fwd(degrees_to_use, viewdirs, coeffs):
        calls.append("sh_fwd")
        return t(c_oracle.sh_fwd(degrees_to_use, viewdirs.detach().numpy(), coeffs.detach().numpy()))

    def sh_bwd(degree, degrees_to_use, viewdirs, v_colors):
        calls.append("sh_bwd")
        return t(c_oracle.sh_bwd(degree, degrees_to_use, viewdirs.detach().numpy(), v_colors.detach().numpy()))

    class CpuBinning:
        def __init__(self, n, ids, ranges, tb):
            self.n, self.n_views, self.ids, self.ranges, self.tile_bounds = n, 1, ids, ranges, tb

    cache = {}

    def binning_for(xys, depths, radii, num_tiles_hit, H, W):
        key = (xys.data_ptr(), xys._version, depths.data_ptr(), int(H), int(W))
        if cache.get("key") != key:
            calls.append("bin")
            tb = ((W + 15) // 16, (H + 15) // 16, 1)
            _, _, _, _, ids, ranges = c_oracle.bin_and_sort(xys.detach().numpy(), depths.detach().numpy(), radii.numpy(),
                                                            num_tiles_hit.numpy(), tb)
            cache["key"], cache["b"] = key, CpuBinning(xys.shape[0], ids, ranges, tb)
        return cache["b"]

    def pack_geo(n, n_views, xys, conics, opacity, opac_per_view=False):
        return torch.cat([xys, conics, opacity.reshape(-1, 1)], dim=1)     # [n, 6]: x y A B C o

    def blend_fwd(b, geo, colors, bg, H, W, single_image=False, **kw):
        calls.append(f"blend_fwd[{colors.shape[1]}]")
        g = geo.numpy()
        out, fT, fi, _, _ = c_oracle.blend_fwd(H, W, b.tile_bounds, b.ids, b.ranges, g[:, 0:2], g[:, 2:5], g[:, 5], colors.numpy(),
                                               bg.numpy())
        return t(out), t(fT), t(fi), None

    def blend_bwd(b, geo, colors, bg, final_T, final_idx, v_out, H, W, hit_words=None, **kw):
        calls.append(f"blend_bwd[{colors.shape[1]}]")
        g = geo.numpy()
        v_xy, v_con, v_col, v_op = c_oracle.blend_bwd(H, W, b.tile_bounds, b.ids, b.ranges, g[:, 0:2], g[:, 2:5], g[:, 5],
                                                      colors.numpy(), bg.numpy(), v_out.reshape(H, W, -1).contiguous().numpy())
        return (t(v_xy).float(), t(v_con).float(), t(v_op).float()), t(v_col).float()

    def unpack_vgeo(n, n_views, v_geo):
        return v_geo

    ops.project_fwd, ops.project_bwd, ops.sh_fwd, ops.sh_bwd = project_fwd, project_bwd, sh_fwd, sh_bwd
    ops.pack_geo, ops.blend_fwd, ops.blend_bwd, ops.unpack_vgeo = pack_geo, blend_fwd, blend_bwd, unpack_vgeo
    _raster.binning_for = binning_for
    return calls


def main():
    sys.path.insert(0, ROOT)
    sys.path.insert(0, REF)
    install_stubs()
    import nerfstudio.models.gaussian_splatting as gs
    from nerfstudio.cameras.cameras import Cameras, CameraType
    from nerfstudio.data.scene_box import SceneBox
    import gaussiangrasper_b200 as gg
    # the model file is bound to this repository's classes through the exact import paths it names (:46-50)
    assert gs.ProjectGaussians is gg.ProjectGaussians and gs.RasterizeGaussians is gg.RasterizeGaussians
    assert gs.NDRasterizeGaussians is gg.NDRasterizeGaussians and gs.SphericalHarmonics is gg.SphericalHarmonics
    assert gs.num_sh_bases(4) == 25 and gs.quat_to_rotmat.__module__.startswith("gaussiangrasper_b200")
    calls = install_oracle_backend()

    torch.manual_seed(0)
    # 4000 seed points (the default random initialisation hard-codes 500 000; the seeded branch calls .cuda())
    orig_rand = torch.rand

    def small_rand(*size, **kw):
        if size and size[0] == (500000, 3):
            return orig_rand((4000, 3), **kw)
        return orig_rand(*size, **kw)
    torch.rand = small_rand
    try:
        model = gs.GaussianSplattingModel(gs.GaussianSplattingModelConfig(),
                                          scene_box=SceneBox(aabb=torch.tensor([[-1.0, -1, -1], [1, 1, 1]])), num_train_data=4)
    finally:
        torch.rand = orig_rand
    model.train()
    with torch.no_grad():
        model.means.mul_(0.3)
        model.scales.add_(1.0)            # bigger splats: the 64x48 image sees hundreds of them
        model.opacities.fill_(1.0)
    model.step = 600                      # past the half-resolution warm-up (:599-603) and at SH degree 0
    H, W = 48, 64
    c2w = torch.tensor([[[1.0, 0, 0, 0.0], [0, 1, 0, 0.0], [0, 0, 1, 4.0]]])     # camera at z = +4 looking down -z
    cam = Cameras(camera_to_worlds=c2w, fx=60.0, fy=60.0, cx=W / 2, cy=H / 2, width=W, height=H, camera_type=CameraType.PERSPECTIVE)
    out = model.get_outputs(cam)
    assert set(out) >= {"rgb", "feature", "depth", "normal"}
    assert out["rgb"].shape == (H, W, 3) and out["feature"].shape == (H, W, 32) and out["depth"].shape == (H, W, 1)
    assert out["normal"].shape == (H, W, 3)
    n_fwd = [c for c in calls if c.startswith("blend_fwd")]
    assert n_fwd == ["blend_fwd[3]", "blend_fwd[32]", "blend_fwd[3]", "blend_fwd[3]"], n_fwd    # rgb, feature, depth, normal
    assert calls.count("bin") == 1, "the four rasterizations of one projection share one binning"
    visible = int((model.radii > 0).sum())
    assert visible > 500 and float(out["rgb"].detach().max()) > 0.05
    assert model.xys.requires_grad and model.radii.dtype == torch.int32

    g = torch.Generator().manual_seed(1)
    batch = dict(image=torch.rand((H, W, 3), generator=g), normal=torch.randn((H, W, 3), generator=g),
                 depth=torch.rand((H, W, 1), generator=g) * 5 + 0.1,
                 sam_mask=torch.randint(0, 3, (H, W), generator=g), valid_mask=torch.rand((H, W), generator=g) > 0.1,
                 feature=torch.randn((H, W, 512), generator=g))
    losses = model.get_loss_dict(out, batch)     # writes zeros into outputs["rgb"] in place (:884) before backward
    assert set(losses) == {"main_loss", "feature_loss", "up_loss", "depth_loss", "normal_loss", "sh_reg", "scale_reg"}
    total = sum(losses.values())
    total.backward()
    names = dict(means=model.means, scales=model.scales, quats=model.quats, opacities=model.opacities,
                 colors_all=model.colors_all, feature=model.feature)
    grads = {}
    for k, p in names.items():
        assert p.grad is not None and torch.isfinite(p.grad).all(), k
        grads[k] = float(p.grad.abs().max())
        assert grads[k] > 0, k
    assert model.xys.grad is not None and float(model.xys.grad.abs().max()) > 0      # densification statistic (:377)
    model.after_train(600)
    assert model.xys_grad_norm is not None and model.max_2Dsize is not None
    assert float(model.max_2Dsize.max()) > 0
    back = [c for c in calls if c.startswith("blend_bwd")]
    assert sorted(back) == sorted(["blend_bwd[3]", "blend_bwd[32]", "blend_bwd[3]", "blend_bwd[3]"]), back
    assert "project_bwd" in calls and "sh_bwd" in calls
    # error behaviour of the boundary as gsplat's: wrong shapes raise ValueError
    try:
        gs.RasterizeGaussians.apply(model.xys.detach(), torch.zeros(4000), model.radii, torch.zeros(4000, 3), model.radii,
                                    torch.zeros(4000, 4), torch.ones(4000, 1), H, W)
        raise AssertionError("colors [N,4] must be refused by the 3-channel rasterizer")
    except ValueError:
        pass
    print(json.dumps(dict(ok=True, visible=visible, losses={k: float(v) for k, v in losses.items()}, grads=grads,
                          calls=len(calls))))


if __name__ == "__main__":
    main()
