"""The fixtures the REFERENCE ITSELF produced (tests/golden/make_reference_golden.py; the unmodified
nerfstudio/models/gaussian_splatting.py imported and run in the build container) against

  * the CPU oracle -- oracle/loss_oracle.py and oracle/refine_oracle.py are the checkers of the GPU loss / refinement
    tests; here they are pinned to the reference's own numbers;
  * the host-side constants and formulas of the product (projection matrix, SH constant, optimizer table, learning
    rate schedules);
  * the generator itself: where the reference tree is mounted, a fresh run must reproduce the committed files.

The GPU side of the same checks is tests/test_gpu_zz_reference_golden.py."""
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_golden_checks as checks  # noqa: E402

REFERENCE = "/root/reference/nerfstudio/models/gaussian_splatting.py"


def test_loss_oracle_reproduces_the_references_get_loss_dict():
    got = checks.check_losses(checks.OracleBackend(), rel_value=2e-6, rel_grad=2e-5)
    assert set(got) == {"main_loss", "feature_loss", "up_loss", "depth_loss", "normal_loss", "sh_reg", "scale_reg"}
    assert all(v > 0.1 for v in got.values())              # every term is active in the fixture


def test_references_initialisation_and_up_projection():
    checks.check_init(checks.OracleBackend())


def test_oracle_sh_basis_equals_the_references_own_real_sh_basis():
    """Row a2: both restatements of gsplat's SH table against nerfstudio/utils/math.py's basis."""
    from oracle import c_oracle, torch_oracle
    checks.check_sh_basis(lambda deg, d, c: c_oracle.sh_fwd(deg, d.numpy(), c.numpy()))
    checks.check_sh_basis(lambda deg, d, c: torch_oracle.spherical_harmonics(deg, d.double(), c.double()))
    checks.check_sh_gradient(lambda deg, d, v: c_oracle.sh_bwd(4, deg, d.numpy(), v.numpy()))


def test_quaternion_convention_equals_the_references_own_quaternion_matrix():
    """Rows a1 / a6: the product's quat_to_rotmat, the oracles' and the covariance the projection oracle emits."""
    from gaussiangrasper_b200 import _torch_impl
    from oracle import c_oracle, refine_oracle, torch_oracle
    checks.check_quaternion_convention(_torch_impl.quat_to_rotmat)
    checks.check_quaternion_convention(refine_oracle.quat_to_rotmat)
    checks.check_quaternion_convention(torch_oracle.quat_to_rotmat)

    def project(means, scales, quats, cam):
        out = c_oracle.project_fwd(means.numpy(), scales.numpy(), 1.0, quats.numpy(), cam.viewmat[:3].numpy(),
                                   cam.fullmat.numpy(), cam.fx, cam.fy, cam.cx, cam.cy, cam.H, cam.W, cam.tile_bounds)
        return out[5]
    checks.check_quaternion_convention(None, project)


def test_samplers_draw_what_the_references_samplers_drew():
    """losses.sampling_pairs_in_mask / sampling_in_mask (:120-148) on the segment mask of the fixture, from the seed the
    reference's get_loss_dict ran under: the same pixel lists, in the same order."""
    from gaussiangrasper_b200 import losses
    fix = checks.load("ref_losses_small")
    mask = torch.from_numpy(fix["gt_mask"])
    torch.manual_seed(13)                                   # make_reference_golden.losses_fixture
    pairs = losses.sampling_pairs_in_mask(mask, int(fix["pairs_num"][0]))
    points = losses.sampling_in_mask(mask, int(fix["points_num"][0]))
    assert len(pairs) == int(fix["n_segments"][0])
    for i, (a, b) in enumerate(pairs):
        assert torch.equal(a, torch.from_numpy(fix[f"pairs_{i}_a"])) and torch.equal(b, torch.from_numpy(fix[f"pairs_{i}_b"])), i
    assert torch.equal(points, torch.from_numpy(fix["points"]))


def test_host_side_of_the_fused_adam_follows_the_references_own_train_iteration():
    """FusedAdam.plan, the REFERENCE_* tables and exponential_decay_lr (the host logic around gg_adam_step) against
    the trajectory of the reference's own Trainer.train_iteration."""
    checks.check_trainer(checks.OracleBackend())


def test_ply_writer_carries_what_the_references_exporter_hands_to_open3d(tmp_path):
    """ply.write_ply against the attribute map the reference's own ExportGaussianSplat.main (scripts/exporter.py:482-530)
    built for the same parameters (open3d's PointCloud call recorded): every attribute, value for value, with the
    uint8 colours' wrap-around cast, the single f_rest_0 column, and the exporter's insertion order (positions,
    normals, colors, f_dc_*, f_rest_0, opacity, scale_*, rot_*)."""
    from gaussiangrasper_b200 import ply
    fix = checks.load("ref_ply_small")
    names = ("means", "log_scales", "quats", "opacity_logit", "sh_coeffs", "features")
    params = {k: torch.from_numpy(fix["param_" + k]) for k in names}
    path = str(tmp_path / "point_cloud.ply")
    assert ply.write_ply(path, params) == params["means"].shape[0]
    got = ply.read_ply(path)
    expand = dict(positions=("x", "y", "z"), normals=("nx", "ny", "nz"), colors=("red", "green", "blue"))
    order = []
    for attr in fix["attribute_names"].tolist():
        want = fix["attr_" + attr]
        cols = expand.get(attr, (attr,))
        assert want.shape[1] == len(cols), attr
        for j, c in enumerate(cols):
            assert got[c].dtype == want.dtype, (c, got[c].dtype, want.dtype)
            assert np.array_equal(got[c], want[:, j]), c
            order.append(c)
    assert list(got) == order
    assert (fix["attr_colors"] == 0).sum() + (fix["attr_colors"] == 255).sum() < fix["attr_colors"].size   # a real test of the cast
    colors = fix["param_sh_coeffs"][:, 0, :] * 0.28209479177387814 + 0.5
    assert (colors < 0).any() and (colors > 1).any()          # values outside [0, 1] went through the uint8 cast


def test_model_oracle_reproduces_the_references_get_outputs_and_backward():
    """Row a7: `oracle_model` (tests/test_gpu_parity.py) -- the fp64 restatement of the model-side path that the GPU
    tests of the fused render_views use as their checker -- against the reference's own get_outputs + backward."""
    import test_gpu_parity as tgp

    def render(params, cam, v):
        outs, grads = tgp.oracle_model(params, cam, v)
        return {k: o.detach().numpy() for k, o in outs.items()}, {k: g.numpy() for k, g in grads.items()}
    checks.check_outputs(render, rtol_grad=2e-5)


def test_refine_schedule_equals_the_rules_the_references_refinement_after_applied():
    """training.refine_schedule against what the reference's own refinement_after (:402-464) was OBSERVED to do at every
    refinement step of a run (probe Gaussians whose fate depends on one rule each; -1 = not observable at that step),
    for three training-set sizes."""
    from gaussiangrasper_b200 import training
    fix = checks.load("ref_init_small")
    cols = fix["schedule_columns"].tolist()
    assert cols[:3] == ["step", "num_train_data", "active"]
    seen = {k: set() for k in cols[3:]}
    for step, num_train_data, active, *flags in fix["schedule_rows"].tolist():
        got = training.refine_schedule(step, num_train_data, 640)
        if not active:
            assert got is None, step
            continue
        assert got is not None, step
        for k, v in zip(cols[3:], flags):
            if v >= 0:
                assert int(got[k]) == v, (step, num_train_data, k)
                seen[k].add(v)
    assert all(v == {0, 1} for v in seen.values()), seen          # every rule was seen both on and off


def test_prepare_targets_equals_the_ground_truth_side_of_the_references_get_loss_dict():
    """losses.prepare_targets against the local tensors of the reference's own get_loss_dict (:849-875), read from its
    frame: at half resolution (training step 100: image / normal / depth resized bilinearly, masks and features by
    nearest neighbour, depth validity decided before the resize) bit for bit, and at full resolution."""
    from gaussiangrasper_b200 import losses, scenes
    fix = checks.load("ref_targets_small")
    d = int(fix["downscale"][0])
    assert d == 2 == scenes.downscale_factor(int(fix["step"][0]))
    batch = {k[len("batch_"):]: torch.from_numpy(v) for k, v in fix.items() if k.startswith("batch_")}
    t = losses.prepare_targets(batch, d)
    loc = {k[len("local_"):]: torch.from_numpy(v) for k, v in fix.items() if k.startswith("local_")}
    valid = loc["valid_mask"]
    assert torch.equal(t["valid"], valid) and 0.05 < float((~valid).float().mean()) < 0.5
    assert torch.equal(t["image"][valid], loc["gt_img"][valid])        # (the reference zeroes its copy's invalid pixels later, :883)
    assert torch.equal(t["normal"].permute(2, 0, 1), loc["gt_normal"])
    assert torch.equal(t["depth"][None], loc["gt_depth"]) and torch.equal(t["depth_mask"][None], loc["depth_mask"])
    assert torch.equal(t["segments"], loc["gt_mask"]) and torch.equal(t["feature"].permute(2, 0, 1), loc["gt_fea"].float())
    assert bool((t["depth_mask"] != ((t["depth"] > 0.05) & valid)).any())   # validity was decided before the resize: it matters here
    # full resolution: the fixture of the loss values
    fix1 = checks.load("ref_losses_small")
    batch1 = {k[len("batch_"):]: torch.from_numpy(v) for k, v in fix1.items() if k.startswith("batch_") and k != "batch_feature_x8"}
    batch1["feature"] = torch.from_numpy(fix1["batch_feature_x8"]).float() / 8
    t1, want = losses.prepare_targets(batch1, 1), checks.ground_truth(fix1)
    assert torch.equal(t1["segments"], torch.from_numpy(fix1["gt_mask"]))
    for k in ("image", "depth", "depth_mask", "valid", "feature"):
        assert torch.equal(t1[k], want[k]), k
    assert torch.allclose(t1["normal"], want["normal"], rtol=0, atol=2e-7)     # normalised in another memory layout there


def test_references_after_train_statistics():
    checks.check_after_train(checks.OracleBackend())


def test_refinement_fixture_is_the_references_and_the_oracle_reproduces_it():
    """refine_small.npz is written by GaussianSplattingModel.refinement_after itself; oracle/refine_oracle.py must give
    the same Gaussians and Adam moments (the detailed comparison is tests/test_golden.py)."""
    g = checks.load("refine_small")
    assert str(g["generated_by"][0]).startswith("reference:")
    n_in = g["in_means"].shape[0]
    assert int(g["n_samples_drawn"][0]) % 2 == 0 and int(g["n_cat"][0]) > n_in + int(g["n_samples_drawn"][0]) > n_in
    assert int(g["n_out"][0]) < int(g["n_cat"][0])          # splits, duplicates and culls all occur


def test_host_constants_match_the_reference():
    from gaussiangrasper_b200 import ply, scenes, training
    from gaussiangrasper_b200.checkpoint import GROUP_MAP
    fix = checks.load("ref_init_small")
    for args, want in zip(fix["proj_args"], fix["proj_mats"]):
        got = scenes.projection_matrix(*[float(a) for a in args]).numpy()
        assert got.tobytes() == want.tobytes(), args
    assert np.array_equal(ply.sh2rgb(fix["sh_dc"]).astype(np.float32), fix["sh2rgb"])
    # the optimizer table of configs/method_configs.py:611-664
    table = {g: row for g, row in zip(fix["opt_groups"].tolist(), fix["opt_lr_eps_final_maxsteps"])}
    for group, ours in GROUP_MAP.items():
        lr, eps, lr_final, max_steps = table[group]
        assert training.REFERENCE_LRS[ours] == lr and eps == 1e-15, group
        if max_steps > 0:
            assert training.REFERENCE_SCHEDULES[ours] == (lr_final, int(max_steps)), group
        else:
            assert ours not in training.REFERENCE_SCHEDULES, group
    acc = dict(zip(fix["accumulation_groups"].tolist(), fix["accumulation_steps"].tolist()))
    for group, ours in GROUP_MAP.items():
        assert training.REFERENCE_ACCUMULATION.get(ours, 1) == acc.get(group, 1), group
    # the SH bands get_outputs asked SphericalHarmonics.apply for, per step (:729)
    for step, want in zip(fix["sh_steps"].tolist(), fix["sh_degrees_to_use"].tolist()):
        assert training.sh_degrees_to_use(step) == want, step
    assert sorted(set(fix["sh_degrees_to_use"].tolist())) == [0, 1, 2, 3, 4]
    # ExponentialDecayScheduler (engine/schedulers.py:109-140) at the probed steps
    for group, row, lrs in zip(fix["opt_groups"].tolist(), fix["opt_lr_eps_final_maxsteps"], fix["opt_probe_lrs"]):
        if row[3] < 0:
            continue
        for step, want in zip(fix["opt_probe_steps"].tolist(), lrs.tolist()):
            got = training.exponential_decay_lr(row[0], row[2], int(row[3]), step)
            assert got == pytest.approx(want, rel=1e-12), (group, step)


def test_camera_setup_equals_what_get_outputs_hands_to_the_projection():
    """scenes.camera_from_c2w / cameras_from_nerfstudio against the arguments the reference's get_outputs (:624-713)
    passed to ProjectGaussians.apply for the same nerfstudio cameras: bit for bit."""
    import types
    from gaussiangrasper_b200 import scenes
    from gaussiangrasper_b200.render import ViewBatch
    fix = checks.load("ref_init_small")
    V = len(fix["cam_c2w"])
    assert V >= 4
    cams = []
    for i in range(V):
        fx, fy, cx, cy = fix["cam_intr_in"][i]
        W, H = (int(v) for v in fix["cam_size_in"][i])
        d = scenes.downscale_factor(int(fix["cam_step"][i]))
        assert d == int(fix["cam_downscale"][i])
        c = scenes.camera_from_c2w(fix["cam_c2w"][i], fx, fy, cx, cy, W, H, downscale=d)
        assert c.viewmat[:3].numpy().tobytes() == fix["cam_viewmat"][i].tobytes(), i
        assert c.fullmat.numpy().tobytes() == fix["cam_fullmat"][i].tobytes(), i
        assert [c.fx, c.fy, c.cx, c.cy] == fix["cam_intr"][i].tolist() and [c.H, c.W] == fix["cam_size"][i].tolist()
        assert list(c.tile_bounds) == fix["cam_tile_bounds"][i].tolist()
        assert torch.equal(c.position, torch.from_numpy(fix["cam_c2w"][i][:, 3]))
        cams.append(c)
    # a nerfstudio-shaped batch ([V,3,4] poses, [V,1] intrinsics) of equal-sized views -> the same cameras, and the
    # packed ViewBatch the fused path consumes
    assert sorted(set(fix["cam_downscale"].tolist())) == [1, 2]
    same = [i for i in range(V) if fix["cam_size_in"][i].tolist() == fix["cam_size_in"][0].tolist()
            and int(fix["cam_downscale"][i]) == 1] + [0]
    col = lambda j: torch.tensor([[fix["cam_intr_in"][i][j]] for i in same], dtype=torch.float32)
    W, H = (int(v) for v in fix["cam_size_in"][0])
    batch = types.SimpleNamespace(camera_to_worlds=torch.from_numpy(fix["cam_c2w"][same]), fx=col(0), fy=col(1), cx=col(2),
                                  cy=col(3), width=torch.full((len(same), 1), W), height=torch.full((len(same), 1), H))
    got = scenes.cameras_from_nerfstudio(batch)
    assert len(got) == len(same)
    for c, i in zip(got, same):
        assert torch.equal(c.viewmat, cams[i].viewmat) and torch.equal(c.fullmat, cams[i].fullmat) and c.fx == cams[i].fx
    vb = ViewBatch.from_cameras(got, torch.device("cpu"))
    assert vb.n_views == len(same) and (vb.H, vb.W) == (H, W)
    assert vb.viewmats[0].numpy().tobytes() == fix["cam_viewmat"][same[0]].tobytes()
    assert vb.fullmats[-1].numpy().tobytes() == fix["cam_fullmat"][0].tobytes()
    assert torch.equal(vb.positions[0], cams[same[0]].position)


def test_gpu_reference_tests_dry_run_on_the_cpu(tmp_path):
    """tests/dry_run_gpu_reference_tests.py: every test function of tests/test_gpu_zz_reference_golden.py executed with
    the CUDA entry points replaced by oracle-backed stand-ins (a check of the GPU tests' own code)."""
    r = subprocess.run([sys.executable, os.path.join(HERE, "dry_run_gpu_reference_tests.py")], capture_output=True, text=True,
                       timeout=900, cwd=str(tmp_path))
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    assert "DRY RUN PASSED 7" in r.stdout, r.stdout[-2000:]


@pytest.mark.skipif(not os.path.exists(REFERENCE), reason="the reference tree is not mounted on this machine")
def test_the_reference_regenerates_the_committed_fixtures(tmp_path):
    r = subprocess.run([sys.executable, os.path.join(HERE, "golden", "make_reference_golden.py"), "--out", str(tmp_path)],
                       capture_output=True, text=True, timeout=900, cwd=str(tmp_path))
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    for name in ("refine_small", "ref_losses_small", "ref_init_small", "ref_trainer_small", "ref_ply_small", "ref_outputs_small", "ref_targets_small"):
        want, got = checks.load(name), dict(np.load(os.path.join(str(tmp_path), name + ".npz")))
        assert set(want) == set(got), name
        for k in want:
            if want[k].dtype.kind in "fc" and name != "refine_small":
                np.testing.assert_allclose(got[k], want[k], rtol=1e-5, atol=1e-7 * float(np.nanmax(np.abs(want[k])) + 1e-30),
                                           err_msg=f"{name}:{k}", equal_nan=True)
            else:
                assert np.array_equal(got[k], want[k]), f"{name}:{k}"     # the refinement fixture: bit for bit
