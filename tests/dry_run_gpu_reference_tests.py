"""Dry run of tests/test_gpu_zz_reference_golden.py WITHOUT a GPU (run as a script by
tests/test_reference_golden_cpu.py): the file's test functions are executed with "cuda:0" read as "cpu" and every CUDA
entry point they reach replaced by a CPU stand-in built on the oracle (ref_golden_checks.OracleBackend, the oracle
backend of reference_model_driver, a torch_oracle version of render_views).  What this checks is the GPU tests' own
code -- argument orders, shapes, holder keys, tolerances that the oracle must meet too -- so that a mistake in a test
cannot first show up on the GPU box.  TEST INFRASTRUCTURE ONLY; nothing here is a product path."""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path[:0] = [HERE, os.path.dirname(HERE)]
import numpy as np, torch
import ref_golden_checks as checks
import reference_model_driver as drv
from gaussiangrasper_b200 import losses, training, render, scenes
import gaussiangrasper_b200 as gg
from gaussiangrasper_b200 import _lib
ob = checks.OracleBackend()
drv.install_oracle_backend()          # ProjectGaussians / SphericalHarmonics run on the CPU oracle
for k in ('geom_loss', 'contrastive_feature_loss', 'up_loss', 'param_regs', 'up_project'):
    setattr(losses, k, getattr(ob, k))
for k in ('pixel_loss', 'ssim_loss', 'knn_scale_init'):
    setattr(training, k, getattr(ob, k))
class DS:
    def __new__(cls, n, dev): return ob.make_stats(n)
training.DensifyStats = DS
training.FusedAdam.reference = classmethod(lambda cls, params: ob.make_adam(params))
class UP(torch.nn.Module):
    def __init__(self, D):
        super().__init__()
        from oracle import loss_oracle
        self.m = loss_oracle.MLP(in_dim=D, out_dim=512, hidden_list=[128]).double()
        self.layers = self.m.layers
    def forward(self, x): return self.m(x)
    def load_state_dict(self, sd): return self.m.load_state_dict({k: v.double() for k, v in sd.items()})
    def named_parameters(self): return self.m.named_parameters()
    def to(self, d): return self
losses.UpProjection = UP
_lib.launch_count = iter(range(1000)).__next__
import test_gpu_parity as tgp
def fake_render_views(means, log_scales, quats, opacity_logit, sh_coeffs, features, views, degrees_to_use=4, holder=None):
    from oracle import c_oracle, torch_oracle
    cam = views
    P = dict(means=means, log_scales=log_scales, quats=quats, opacity_logit=opacity_logit, sh_coeffs=sh_coeffs, features=features)
    Pd = {k: v.double() for k, v in P.items()}
    q = Pd["quats"] / Pd["quats"].norm(dim=-1, keepdim=True)
    xys, depths, radii, conics, nth, _ = torch_oracle.project_gaussians(Pd["means"], torch.exp(Pd["log_scales"]), 1.0, q, cam.viewmat, cam.fullmat, cam.fx, cam.fy, cam.cx, cam.cy, cam.H, cam.W, cam.tile_bounds)
    sc = {k: v.detach() for k, v in P.items()}
    f32 = tgp.oracle_project(sc, cam, sc["log_scales"].exp(), sc["quats"] / sc["quats"].norm(dim=-1, keepdim=True))
    _, _, _, _, ids_s, ranges = c_oracle.bin_and_sort(f32[0], f32[1], f32[2], f32[4], cam.tile_bounds)
    n = means.shape[0]
    def hook(g):
        vg = torch.zeros((n, 8)); vg[:, :2] = g.float(); holder["v_geo"] = vg
    xys.register_hook(hook)
    viewdirs = Pd["means"].detach() - cam.position.double()
    rgbs = torch.clamp(torch_oracle.spherical_harmonics(4, viewdirs, Pd["sh_coeffs"]) + 0.5, 0.0, 1.0)
    op = torch.sigmoid(Pd["opacity_logit"]).reshape(-1)
    R = torch_oracle.quat_to_rotmat(Pd["quats"])
    idx = torch.exp(Pd["log_scales"]).min(dim=-1)[1][..., None, None].expand(-1, 3, -1)
    normals = R.gather(2, idx).squeeze(dim=2)
    D = features.shape[1]
    cols = torch.cat([rgbs, Pd["features"], depths[:, None], normals], dim=1)
    bg = torch.cat([torch.zeros(3 + D), torch.ones(1) * 10, torch.zeros(3)]).double()
    out, _, _ = torch_oracle.rasterize(xys, conics, op, cols, torch.from_numpy(ids_s), torch.from_numpy(ranges), cam.H, cam.W, bg)
    holder["radii"] = torch.from_numpy(f32[2])[None]
    out = out.float()[None]
    return dict(rgb=out[..., :3], feature=out[..., 3:3 + D], depth=out[..., 3 + D:4 + D], normal=out[..., 4 + D:])
render.render_views = fake_render_views
render.ViewBatch.from_cameras = staticmethod(lambda cams, dev: cams[0])
TARGET = os.path.join(HERE, 'test_gpu_zz_reference_golden.py')
src = open(TARGET).read().replace('"cuda:0"', '"cpu"').replace('assert torch.cuda.is_available()', 'pass')
ns = {'__name__': 'dry', '__file__': TARGET}
exec(compile(src, 'dry', 'exec'), ns)
ran = 0
for k, f in list(ns.items()):
    if k.startswith('test_'):
        f()
        ran += 1
        print("ok", k)
print("DRY RUN PASSED", ran)
