"""Checkpoint resume in both directions against the reference's OWN classes (run as a script by
tests/test_reference_model_cpu.py; needs /root/reference): GaussianSplattingModel.load_state_dict (:296-313, strict),
nerfstudio.engine.optimizers.Optimizers (load_optimizers / load_schedulers, :194-210) built from the method's optimizer
table, and the dict layout of Trainer.save_checkpoint (engine/trainer.py:437-449).

  reference -> here : a checkpoint dict assembled the way the trainer does, after three real optimizer / scheduler
                      steps, is read by checkpoint.params_from_reference / optimizer_state_from_reference
  here -> reference : checkpoint.trainer_checkpoint written from that state loads STRICTLY into a fresh reference
                      model and fresh Optimizers; one more identical step on the original and on the resumed objects
                      ends in bit-identical parameters, moments and learning rates
"""
import io
import json
import os
import sys

import numpy as np
import torch

import torch.fx  # noqa: F401

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
for p in (os.path.join(HERE, "golden"), HERE, ROOT, "/root/reference"):
    sys.path.insert(0, p)

from reference_model_driver import install_stubs  # noqa: E402


def main():
    install_stubs()
    import make_reference_golden as mk
    import nerfstudio.models.gaussian_splatting as gs
    from nerfstudio.engine.optimizers import AdamOptimizerConfig, Optimizers
    from nerfstudio.engine.schedulers import ExponentialDecaySchedulerConfig
    from gaussiangrasper_b200 import checkpoint
    from gaussiangrasper_b200.losses import UpProjection
    from gaussiangrasper_b200.training import REFERENCE_SCHEDULES

    table = mk.optimizer_table()          # configs/method_configs.py:611-664, from the reference's source

    def optimizer_config():
        cfg = {}
        for group, (lr, eps, lr_final, max_steps) in zip(table["opt_groups"].tolist(), table["opt_lr_eps_final_maxsteps"]):
            cfg[group] = {"optimizer": AdamOptimizerConfig(lr=float(lr), eps=float(eps)),
                          "scheduler": ExponentialDecaySchedulerConfig(lr_final=float(lr_final), max_steps=int(max_steps))
                          if max_steps > 0 else None}
        return cfg

    def fresh(seed):
        torch.manual_seed(seed)
        model = mk.small_model(gs, 200)
        return model, Optimizers(optimizer_config(), model.get_param_groups())

    def step(model, optimizers, seed):
        g = torch.Generator().manual_seed(seed)
        for group, params in model.get_param_groups().items():
            for p in params:
                p.grad = torch.randn(p.shape, generator=g) * 1e-2
            optimizers.optimizer_step(group)
            if optimizers.config[group]["scheduler"] is not None:
                optimizers.scheduler_step(group)

    model, optimizers = fresh(1)
    for it in range(3):
        step(model, optimizers, 100 + it)
    # engine/trainer.py:437-449 (the pipeline holds the model as `_model`)
    ck = {"step": 3, "pipeline": {"_model." + k: v.clone() for k, v in model.state_dict().items()},
          "optimizers": {k: v.state_dict() for k, v in optimizers.optimizers.items()},
          "schedulers": {k: v.state_dict() for k, v in optimizers.schedulers.items()}, "scalers": {}}
    buf = io.BytesIO()
    torch.save(ck, buf)                   # through the file format, like the trainer (state_dict() hands out live tensors)
    buf.seek(0)
    ck = torch.load(buf, weights_only=False)

    # reference -> here
    P = checkpoint.params_from_reference(ck)
    names = dict(means="means", log_scales="scales", quats="quats", opacity_logit="opacities", sh_coeffs="colors_all", features="feature")
    for ours, attr in names.items():
        assert torch.equal(P[ours], getattr(model, attr).detach()), ours
    state = checkpoint.optimizer_state_from_reference(ck)
    assert set(state) == set(names)
    for group, ours in checkpoint.GROUP_MAP.items():
        opt = optimizers.optimizers[group]
        st = opt.state[opt.param_groups[0]["params"][0]]
        assert state[ours]["step"] == 3 and torch.equal(state[ours]["exp_avg"], st["exp_avg"])
        assert torch.equal(state[ours]["exp_avg_sq"], st["exp_avg_sq"])
        assert state[ours]["lr"] == opt.param_groups[0]["lr"], group
        assert state[ours]["lr_init"] == optimizers.config[group]["optimizer"].lr, group

    # here -> reference
    class ResumedAdam:       # the parts of training.FusedAdam the writer reads (FusedAdam itself holds CUDA tensors)
        betas, eps, schedules = (0.9, 0.999), 1e-15, REFERENCE_SCHEDULES

        def state_dict(self):
            return state
    up = UpProjection(32)
    up.load_state_dict(model.fea_up.state_dict())            # same parameter names
    _, rest = checkpoint.split_reference_state(ck["pipeline"])
    others = [g for g in ck["optimizers"] if g not in checkpoint.GROUP_MAP]
    assert others == ["up_net"]                              # the one group the fused Adam does not own here
    ck2 = checkpoint.trainer_checkpoint(3, P, ResumedAdam(), up_projection=up, extra_pipeline=rest,
                                        extra_optimizers={g: ck["optimizers"][g] for g in others},
                                        extra_schedulers={g: ck["schedulers"][g] for g in others if g in ck["schedulers"]})
    assert set(ck2["optimizers"]) == set(ck["optimizers"]) and set(ck2["schedulers"]) == set(ck["schedulers"])
    buf = io.BytesIO()
    torch.save(ck2, buf)
    buf.seek(0)
    ck2 = torch.load(buf, weights_only=False)
    assert set(ck2) == set(ck) and set(ck2["pipeline"]) == set(ck["pipeline"])
    model2, optimizers2 = fresh(2)                            # other initial values, other point count on purpose
    model2.means = torch.nn.Parameter(model2.means[:150].detach().clone())
    model2.load_state_dict({k[len("_model."):]: v for k, v in ck2["pipeline"].items()}, strict=True)
    optimizers2 = Optimizers(optimizer_config(), model2.get_param_groups())    # the trainer builds them after the load
    optimizers2.load_optimizers({k: v for k, v in ck2["optimizers"].items()})
    optimizers2.load_schedulers(ck2["schedulers"])
    model2.step = model.step
    step(model, optimizers, 999)
    step(model2, optimizers2, 999)
    for attr in names.values():
        assert torch.equal(getattr(model, attr), getattr(model2, attr)), attr
    for group in checkpoint.GROUP_MAP:
        a, b = optimizers.optimizers[group], optimizers2.optimizers[group]
        sa, sb = a.state[a.param_groups[0]["params"][0]], b.state[b.param_groups[0]["params"][0]]
        assert torch.equal(sa["exp_avg"], sb["exp_avg"]) and torch.equal(sa["exp_avg_sq"], sb["exp_avg_sq"]), group
        assert float(sa["step"]) == float(sb["step"]) == 4.0
        assert a.param_groups[0]["lr"] == b.param_groups[0]["lr"], (group, a.param_groups[0]["lr"], b.param_groups[0]["lr"])
    for (k, p), (_, q) in zip(model.fea_up.named_parameters(), model2.fea_up.named_parameters()):
        assert torch.equal(p, q), k
    print(json.dumps(dict(ok=True, groups=sorted(ck2["optimizers"]), schedulers=sorted(ck2["schedulers"]),
                          pipeline_keys=len(ck2["pipeline"]))))


if __name__ == "__main__":
    main()
