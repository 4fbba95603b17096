"""CPU tests of the host logic behind the refinement step (no GPU): the schedule of
GaussianSplattingModel.refinement_after (nerfstudio/models/gaussian_splatting.py:396-410, 456, 471-475, 458-464)
and the layout of the config struct shared with the C ABI."""
import ctypes
import os
import subprocess
import sys

import pytest
import torch

from gaussiangrasper_b200 import training

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _reference_conditions(step, num_train_data, c):
    """The reference's own boolean expressions, spelled out."""
    if step < c["warmup_length"]:
        return None
    reset_interval = c["reset_alpha_every"] * c["refine_every"]
    densify = step < c["stop_split_at"] and step % reset_interval > num_train_data + c["refine_every"]
    cull = step % reset_interval > num_train_data + c["refine_every"]
    return dict(do_densify=int(densify), do_cull=int(cull), split_by_screen=int(step < c["stop_screen_size_at"]),
                cull_by_scale=int(step > c["refine_every"] * c["reset_alpha_every"]),
                cull_by_screen=int(step < c["stop_screen_size_at"]),
                reset_opacity=step % reset_interval == c["refine_every"])


def test_refine_schedule_follows_the_reference_conditions():
    c = training.REFERENCE_REFINE
    for num_train_data in (10, 150):
        for step in list(range(0, 20000, 100)) + [499, 500, 3001, 3100, 3111, 14999, 15000]:
            want = _reference_conditions(step, num_train_data, c)
            got = training.refine_schedule(step, num_train_data, 640)
            if want is None:
                assert got is None, step
                continue
            for k, v in want.items():
                assert got[k] == v, (step, k)
            assert got["max_dim"] == 640.0 and got["cull_alpha_thresh"] == c["cull_alpha_thresh"]
    # overrides
    assert training.refine_schedule(600, 10, 640, dict(warmup_length=1000)) is None


def test_refine_config_struct_matches_the_header(tmp_path):
    src = tmp_path / "sz.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "gg_b200.h"\n'
                   'int main(void){printf("%zu %zu %zu", sizeof(gg_refine_config), offsetof(gg_refine_config, do_densify),'
                   ' offsetof(gg_refine_config, cull_by_screen)); return 0;}\n')
    exe = tmp_path / "sz"
    subprocess.check_call(["/usr/bin/gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    size, off_a, off_b = (int(x) for x in subprocess.check_output([str(exe)]).split())
    assert ctypes.sizeof(training.RefineConfig) == size
    assert training.RefineConfig.do_densify.offset == off_a and training.RefineConfig.cull_by_screen.offset == off_b
    assert [f for f, _ in training.RefineConfig._fields_][:7] == ["max_dim", "densify_grad_thresh", "densify_size_thresh",
                                                                 "split_screen_size", "cull_alpha_thresh",
                                                                 "cull_scale_thresh", "cull_screen_size"]


def test_refine_oracle_counts_are_consistent():
    """The restatement itself: sizes add up and the Adam moments follow the parameters."""
    from oracle import refine_oracle
    g = torch.Generator().manual_seed(0)
    n = 500
    P = dict(means=torch.randn(n, 3, generator=g), log_scales=torch.log(torch.rand(n, 3, generator=g) * 0.03 + 0.002),
             quats=torch.randn(n, 4, generator=g), opacity_logit=torch.randn(n, generator=g),
             sh_coeffs=torch.randn(n, 4, 3, generator=g), features=torch.randn(n, 2, generator=g))
    M = {k: (torch.ones_like(v), torch.ones_like(v) * 2) for k, v in P.items()}
    rules = dict(max_dim=640.0, densify_grad_thresh=0.0002, densify_size_thresh=0.01, split_screen_size=0.05,
                 cull_alpha_thresh=0.1, cull_scale_thresh=0.5, cull_screen_size=0.15, do_densify=1, split_by_screen=1,
                 do_cull=0, cull_by_scale=0, cull_by_screen=0)
    p, m, info = refine_oracle.refine(P, M, torch.rand(n, generator=g) * 1e-5, torch.ones(n), torch.rand(n, generator=g) * 0.2,
                                      rules, lambda k: torch.zeros(k, 3))
    assert info["n_out"] == info["n_cat"] > n and p["means"].shape[0] == info["n_out"]
    assert torch.equal(p["quats"][:n], P["quats"])                      # originals first, untouched
    assert float(m["means"][0][:n].min()) == 1.0 and float(m["means"][0][n:].abs().max()) == 0.0   # children: zero moments
    # zero samples: split children sit on their parent's mean
    new_means = p["means"][n:]
    assert all(bool((P["means"] == r).all(dim=-1).any()) for r in new_means[:20])


def test_reference_checkpoint_round_trip(tmp_path):
    """Parameter names / shapes of a reference checkpoint (trainer.py:427-456, gaussian_splatting.py:300-312)."""
    from gaussiangrasper_b200 import checkpoint
    g = torch.Generator().manual_seed(0)
    n = 17
    model = {"_model.means": torch.randn(n, 3, generator=g), "_model.scales": torch.randn(n, 3, generator=g),
             "_model.quats": torch.randn(n, 4, generator=g), "_model.opacities": torch.randn(n, 1, generator=g),
             "_model.colors_all": torch.randn(n, 25, 3, generator=g), "_model.feature": torch.randn(n, 32, generator=g),
             "_model.camera_optimizer.pose_adjustment": torch.zeros(5, 6), "_model.back_color": torch.zeros(35)}
    path = str(tmp_path / "step-000000100.ckpt")
    torch.save({"step": 100, "pipeline": model, "optimizers": {}, "schedulers": {}, "scalers": {}}, path)
    P = checkpoint.load_reference_checkpoint(path)
    assert set(P) == {"means", "log_scales", "quats", "opacity_logit", "sh_coeffs", "features"}
    assert torch.equal(P["log_scales"], model["_model.scales"]) and P["opacity_logit"].shape == (n, 1)
    assert P["sh_coeffs"].shape == (n, 25, 3) and P["features"].shape == (n, 32)
    back = checkpoint.reference_state_dict(P)
    for k in ("means", "scales", "quats", "opacities", "colors_all", "feature"):
        assert torch.equal(back["_model." + k], model["_model." + k]), k
    # the bare model state dict works too; a broken one is refused
    assert torch.equal(checkpoint.params_from_reference({k[7:]: v for k, v in model.items()})["means"], model["_model.means"])
    bad = dict(model); bad["_model.quats"] = torch.zeros(n, 3)
    import pytest
    with pytest.raises(ValueError):
        checkpoint.params_from_reference(bad)
    del bad["_model.quats"]
    with pytest.raises(KeyError):
        checkpoint.params_from_reference(bad)


def test_ssim_restatement_properties():
    """oracle/loss_oracle.ssim: 1 for identical images, symmetric, below 1 otherwise, window sums to 1."""
    from oracle import loss_oracle
    g = torch.Generator().manual_seed(0)
    a = torch.rand((1, 3, 40, 33), generator=g, dtype=torch.float64)
    b = (a + 0.1 * torch.randn(a.shape, generator=g, dtype=torch.float64)).clamp(0, 1)
    assert abs(float(loss_oracle.ssim(a, a)) - 1.0) < 1e-9
    assert abs(float(loss_oracle.ssim(a, b)) - float(loss_oracle.ssim(b, a))) < 1e-12
    assert 0.0 < float(loss_oracle.ssim(a, b)) < 1.0
    assert abs(float(loss_oracle._gauss_1d().sum()) - 1.0) < 1e-6 and loss_oracle._gauss_1d().shape == (11,)
    # constant images: mu terms only -> (2ab + c1) / (a^2 + b^2 + c1)
    ca, cb = torch.full((1, 1, 20, 20), 0.3, dtype=torch.float64), torch.full((1, 1, 20, 20), 0.6, dtype=torch.float64)
    want = (2 * 0.3 * 0.6 + 1e-4) / (0.09 + 0.36 + 1e-4)
    assert abs(float(loss_oracle.ssim(ca, cb)) - want) < 1e-5   # the fp32 window sums to 1 only to ~1e-7


def _philox_np(counter, key):
    """Philox4x32-10 as published (Salmon, Moraes, Dror, Shaw: "Parallel random numbers: as easy as 1, 2, 3", SC'11;
    Random123's reference implementation): plain integer arithmetic, written independently of csrc/train.cu."""
    c = [int(x) & 0xFFFFFFFF for x in counter]
    k = [int(x) & 0xFFFFFFFF for x in key]
    for _ in range(10):
        p0, p1 = 0xD2511F53 * c[0], 0xCD9E8D57 * c[2]
        c = [((p1 >> 32) ^ c[1] ^ k[0]) & 0xFFFFFFFF, p1 & 0xFFFFFFFF, ((p0 >> 32) ^ c[3] ^ k[1]) & 0xFFFFFFFF,
             p0 & 0xFFFFFFFF]
        k = [(k[0] + 0x9E3779B9) & 0xFFFFFFFF, (k[1] + 0xBB67AE85) & 0xFFFFFFFF]
    return c


def test_philox_block_function_known_answers():
    """The generator behind the rank-consistent split samples (gg_refine_apply with samples == NULL) against the
    known-answer vectors of Random123 (kat_vectors, philox4x32 10 rounds) and the independent Python
    restatement above -- host build of the same __host__ __device__ function the kernel calls."""
    import ctypes as C
    from gaussiangrasper_b200 import _lib
    lib = _lib.load()
    kats = [((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
            ((0xffffffff,) * 4, (0xffffffff,) * 2, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
            ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
             (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1))]
    for ctr, key, want in kats:
        out = (C.c_uint32 * 4)()
        lib.gg_philox4x32_10_host((C.c_uint32 * 4)(*ctr), (C.c_uint32 * 2)(*key), out)
        assert tuple(out) == want, (ctr, [hex(x) for x in out])
        assert tuple(_philox_np(ctr, key)) == want
    import random
    rnd = random.Random(3)
    for _ in range(200):
        ctr = [rnd.getrandbits(32) for _ in range(4)]
        key = [rnd.getrandbits(32) for _ in range(2)]
        out = (C.c_uint32 * 4)()
        lib.gg_philox4x32_10_host((C.c_uint32 * 4)(*ctr), (C.c_uint32 * 2)(*key), out)
        assert list(out) == _philox_np(ctr, key)


def test_adam_plan_follows_the_trainer_loop():
    """FusedAdam.plan == engine/trainer.py:466-481 for the reference's accumulation table (method_configs.py:611),
    and the learning-rate schedule == ExponentialDecayScheduler's lambda (engine/schedulers.py:122-138)."""
    import math
    from gaussiangrasper_b200.training import (REFERENCE_ACCUMULATION, REFERENCE_LRS, REFERENCE_SCHEDULES,
                                               exponential_decay_lr)
    acc = dict(REFERENCE_ACCUMULATION)
    assert acc == dict(means=10, sh_coeffs=10, features=10)
    # the trainer: zero at step % k == 0, optimizer step at step % k == k - 1
    from gaussiangrasper_b200.training import FusedAdam

    class _B:   # plan() only looks at the names
        names = ["means", "log_scales", "quats", "opacity_logit", "sh_coeffs", "features"]
    opt = FusedAdam.__new__(FusedAdam)
    opt.bucket, opt.accumulation = _B(), acc
    for it in range(40):
        plan = opt.plan(it)
        for name in _B.names:
            k = acc.get(name, 1)
            zero, step = it % k == 0, it % k == k - 1
            want = "step" if k == 1 else ("acc_step" if step else ("acc_first" if zero else "acc"))
            assert plan[name] == want, (it, name)
    for name, (lr_final, max_steps) in REFERENCE_SCHEDULES.items():
        lr0 = REFERENCE_LRS[name]
        assert exponential_decay_lr(lr0, lr_final, max_steps, 0) == pytest.approx(lr0, rel=1e-12)
        assert exponential_decay_lr(lr0, lr_final, max_steps, max_steps) == pytest.approx(lr_final, rel=1e-12)
        assert exponential_decay_lr(lr0, lr_final, max_steps, 10 * max_steps) == pytest.approx(lr_final, rel=1e-12)
        mid = exponential_decay_lr(lr0, lr_final, max_steps, max_steps // 2)
        assert mid == pytest.approx(math.sqrt(lr0 * lr_final), rel=1e-9)


def test_ply_export_has_the_reference_exporters_properties(tmp_path):
    """scripts/exporter.py:482-530: x y z, zero normals, uint8 colours = SH2RGB(dc) * 255, f_dc_*, f_rest_0 (the
    reference's loop over the last axis of a (N, -1, 1) reshape writes exactly one), opacity, scale_*, rot_*."""
    import numpy as np
    from gaussiangrasper_b200 import ply, scenes
    n = 321
    sc = scenes.random_scene(n, feature_dim=4, seed=2)
    path = str(tmp_path / "point_cloud.ply")
    assert ply.write_ply(path, sc) == n
    got = ply.read_ply(path)
    want = ["x", "y", "z", "nx", "ny", "nz", "red", "green", "blue", "f_dc_0", "f_dc_1", "f_dc_2", "f_rest_0", "opacity",
            "scale_0", "scale_1", "scale_2", "rot_0", "rot_1", "rot_2", "rot_3"]
    assert list(got) == want
    assert np.array_equal(got["x"], sc["means"][:, 0].numpy()) and not got["nx"].any()
    colors = sc["sh_coeffs"][:, 0, :].numpy() * 0.28209479177387814 + 0.5
    assert np.allclose(got["f_dc_1"], colors[:, 1]) and got["red"].dtype == np.uint8
    assert np.array_equal(got["red"], (colors.astype(np.float32) * 255).astype(np.uint8)[:, 0])
    assert np.array_equal(got["f_rest_0"], sc["sh_coeffs"][:, 1, 0].numpy())
    assert np.array_equal(got["opacity"], sc["opacity_logit"].reshape(-1).numpy())
    assert np.array_equal(got["scale_2"], sc["log_scales"][:, 2].numpy()) and np.array_equal(got["rot_0"], sc["quats"][:, 0].numpy())
    ply.write_ply(path, sc, full_sh=True)
    full = ply.read_ply(path)
    assert sum(k.startswith("f_rest_") for k in full) == 72
    assert np.array_equal(full["f_rest_24"], sc["sh_coeffs"][:, 1, 1].numpy())     # channel-major: 24 coefficients per colour
    header = open(path, "rb").read(200).decode("ascii", "replace")
    assert header.startswith("ply\nformat binary_little_endian 1.0\nelement vertex 321\nproperty float x\n")


def test_trainer_checkpoint_loads_into_torch_adam(tmp_path):
    """engine/trainer.py:437-449: the saved dict has step / pipeline / optimizers / schedulers / scalers, and every
    optimizer entry is a torch.optim.Adam state dict under the reference's group name."""
    from gaussiangrasper_b200 import checkpoint, scenes
    from gaussiangrasper_b200.losses import UpProjection
    from gaussiangrasper_b200.training import REFERENCE_LRS, REFERENCE_SCHEDULES
    n = 50
    sc = scenes.random_scene(n, feature_dim=32, seed=3)
    g = torch.Generator().manual_seed(1)

    class FakeAdam:      # the parts of training.FusedAdam the writer reads (FusedAdam itself needs CUDA tensors)
        betas, eps, schedules = (0.9, 0.999), 1e-15, REFERENCE_SCHEDULES

        def state_dict(self):
            return {k: dict(step=7 if k != "means" else 0, lr=REFERENCE_LRS[k] * 0.9, lr_init=REFERENCE_LRS[k],
                            exp_avg=torch.randn(sc[k].shape, generator=g), exp_avg_sq=torch.rand(sc[k].shape, generator=g))
                    for k in sc}
    up = UpProjection(32)
    ck = checkpoint.trainer_checkpoint(1234, sc, FakeAdam(), up_projection=up)
    assert set(ck) == {"step", "pipeline", "optimizers", "schedulers", "scalers"} and ck["step"] == 1234
    assert set(ck["optimizers"]) == {"xyz", "scaling", "rotation", "opacity", "color", "feature"}
    assert set(ck["schedulers"]) == {"xyz", "scaling", "color", "feature"}          # opacity / rotation have none
    assert {"_model.fea_up.layers.0.weight", "_model.fea_up.layers.2.bias"} <= set(ck["pipeline"])
    ref_names = dict(xyz="means", scaling="scales", rotation="quats", opacity="opacities", color="colors_all", feature="feature")
    for group, leaf in ref_names.items():
        p = torch.nn.Parameter(ck["pipeline"]["_model." + leaf].clone())
        opt = torch.optim.Adam([p], lr=1.0, eps=1e-15)
        import copy
        opt.load_state_dict(copy.deepcopy(ck["optimizers"][group]))   # the reference's Optimizers.load_optimizers does this
        if group == "xyz":
            assert len(opt.state) == 0                        # never stepped: torch keeps no state either
        else:
            st = opt.state[p]
            assert st["exp_avg"].shape == p.shape and float(st["step"]) == 7.0
        assert opt.param_groups[0]["lr"] == pytest.approx(REFERENCE_LRS[checkpoint.GROUP_MAP[group]] * 0.9)
        p.grad = torch.ones_like(p)
        opt.step()                                            # and it keeps training
    path = str(tmp_path / "step-000001234.ckpt")
    torch.save(ck, path)
    loaded = torch.load(path, weights_only=False)
    back = checkpoint.optimizer_state_from_reference(loaded)
    assert set(back) == set(sc) - {"means"} and back["quats"]["step"] == 7
    assert torch.equal(back["opacity_logit"]["exp_avg"], ck["optimizers"]["opacity"]["state"][0]["exp_avg"])
    P = checkpoint.params_from_reference(loaded)
    assert torch.equal(P["features"], sc["features"])
    gauss, rest = checkpoint.split_reference_state(loaded["pipeline"])
    assert len(gauss) == 6 and all("fea_up" in k for k in rest)
