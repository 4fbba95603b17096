"""Reading / writing the Gaussian parameters of a reference checkpoint (formats row).

The reference trainer saves `{"step", "pipeline": pipeline.state_dict(), "optimizers", ...}` with torch.save
(nerfstudio/engine/trainer.py:427-456); the Gaussian model's parameters sit in the pipeline state dict as
`_model.means`, `_model.scales`, `_model.quats`, `_model.opacities`, `_model.colors_all`, `_model.feature`
(nerfstudio/models/gaussian_splatting.py:255-281, load_state_dict :300-312).  This module maps them to the
names `render_views` takes and back, so a scene trained with the reference can be rendered here and vice
versa.  Pure host code (the reference's is Python too); no GPU involved.
"""
from __future__ import annotations

from typing import Dict, Mapping, Optional

import torch

# reference parameter name -> ours
NAME_MAP = dict(means="means", scales="log_scales", quats="quats", opacities="opacity_logit", colors_all="sh_coeffs",
                feature="features")
_SHAPES = dict(means=(3,), scales=(3,), quats=(4,), opacities=(1,))


def _find(state: Mapping[str, torch.Tensor], leaf: str) -> str:
    hits = [k for k in state if k == leaf or k.endswith("." + leaf)]
    if len(hits) != 1:
        raise KeyError(f"expected exactly one '{leaf}' entry in the state dict, found {hits}")
    return hits[0]


def params_from_reference(ckpt: Mapping) -> Dict[str, torch.Tensor]:
    """ckpt: a loaded reference checkpoint, its "pipeline" state dict, or the model's own state dict."""
    state = ckpt["pipeline"] if "pipeline" in ckpt and isinstance(ckpt["pipeline"], Mapping) else ckpt
    out = {}
    n = None
    for ref_name, ours in NAME_MAP.items():
        t = state[_find(state, ref_name)].detach().to(torch.float32).contiguous()
        if n is None:
            n = t.shape[0]
        if t.shape[0] != n:
            raise ValueError(f"{ref_name} has {t.shape[0]} rows, means has {n}")
        if ref_name in _SHAPES and tuple(t.shape[1:]) != _SHAPES[ref_name]:
            raise ValueError(f"{ref_name} has shape {tuple(t.shape)}, expected [N, {_SHAPES[ref_name][0]}]")
        if ref_name == "colors_all" and (t.dim() != 3 or t.shape[2] != 3 or t.shape[1] not in (1, 4, 9, 16, 25)):
            raise ValueError(f"colors_all has shape {tuple(t.shape)}, expected [N, (deg+1)^2, 3]")
        out[ours] = t
    return out


def reference_state_dict(params: Mapping[str, torch.Tensor], prefix: str = "_model.",
                         up_projection: Optional[torch.nn.Module] = None,
                         extra: Optional[Mapping[str, torch.Tensor]] = None) -> Dict[str, torch.Tensor]:
    """The inverse: the Gaussian entries of a reference pipeline state dict.

    `GaussianSplattingModel.load_state_dict` (:300-312) resizes its six Gaussian parameters to the checkpoint and
    then forwards to `Module.load_state_dict(dict, **kwargs)`: a STRICT load also wants the model's other entries
    -- the CLIP up-projection `fea_up.layers.{0,2}.{weight,bias}` and whatever the camera optimizer holds.  Pass
    `up_projection` (losses.UpProjection: same parameter names) and / or `extra` (entries taken over verbatim, e.g.
    the non-Gaussian entries of the checkpoint the scene came from, see split_reference_state) to produce them;
    without them the result loads with `strict=False` only."""
    out = {}
    for ref_name, ours in NAME_MAP.items():
        t = params[ours].detach().to(torch.float32)
        if ref_name == "opacities":
            t = t.reshape(-1, 1)
        out[prefix + ref_name] = t.contiguous()
    if up_projection is not None:
        for k, v in up_projection.state_dict().items():
            out[f"{prefix}fea_up.{k}"] = v.detach().to(torch.float32).cpu().contiguous()
    for k, v in (extra or {}).items():
        out.setdefault(k, v)
    return out


def split_reference_state(state: Mapping[str, torch.Tensor]):
    """(Gaussian entries, everything else) of a reference pipeline state dict."""
    leaves = set(NAME_MAP)
    gauss = {k: v for k, v in state.items() if k.rsplit(".", 1)[-1] in leaves}
    return gauss, {k: v for k, v in state.items() if k not in gauss}


# optimizer group names of the reference (get_gaussian_param_groups :577-586, method_configs.py:618-650) -> ours
GROUP_MAP = dict(xyz="means", scaling="log_scales", rotation="quats", opacity="opacity_logit", color="sh_coeffs",
                 feature="features")


def trainer_checkpoint(step: int, params: Mapping[str, torch.Tensor], optimizer=None, up_projection=None,
                       extra_pipeline: Optional[Mapping[str, torch.Tensor]] = None,
                       extra_optimizers: Optional[Mapping[str, dict]] = None,
                       extra_schedulers: Optional[Mapping[str, dict]] = None) -> dict:
    """The dict the reference trainer saves (engine/trainer.py:437-449): {"step", "pipeline", "optimizers",
    "schedulers", "scalers"}.  `optimizer` is a training.FusedAdam: every group becomes a torch.optim.Adam state dict
    ({"state": {0: {step, exp_avg, exp_avg_sq}}, "param_groups": [...]}) and a LambdaLR-style scheduler entry, under
    the reference's group names, so `Optimizers.load_optimizers` / `load_schedulers` accept them.
    `extra_optimizers` / `extra_schedulers`: state dicts of the groups FusedAdam does not own -- "up_net" (the torch
    optimizer of the up-projection MLP, whose parameters get their `.grad` from losses.up_loss) and "camera_opt" --
    taken over verbatim; without them a resumed reference run restarts those groups' moments."""
    ck = {"step": int(step), "pipeline": reference_state_dict(params, up_projection=up_projection, extra=extra_pipeline),
          "optimizers": {}, "schedulers": {}, "scalers": {}}
    if optimizer is not None:
        st = optimizer.state_dict()
        for group, ours in GROUP_MAP.items():
            if ours not in st:
                continue
            e = st[ours]
            shape = (params[ours].shape[0], 1) if ours == "opacity_logit" else tuple(params[ours].shape)
            ck["optimizers"][group] = {
                "state": {0: {"step": torch.tensor(float(e["step"])), "exp_avg": e["exp_avg"].reshape(shape).cpu(),
                              "exp_avg_sq": e["exp_avg_sq"].reshape(shape).cpu()}} if e["step"] > 0 else {},
                "param_groups": [{"lr": e["lr"], "betas": tuple(optimizer.betas), "eps": optimizer.eps, "weight_decay": 0,
                                  "amsgrad": False, "maximize": False, "foreach": None, "capturable": False,
                                  "differentiable": False, "fused": None, "initial_lr": e["lr_init"], "params": [0]}]}
            if ours in optimizer.schedules:
                ck["schedulers"][group] = {"base_lrs": [e["lr_init"]], "last_epoch": int(step), "_step_count": int(step) + 1,
                                           "_get_lr_called_within_step": False, "_last_lr": [e["lr"]], "lr_lambdas": [None]}
    for k, v in (extra_optimizers or {}).items():
        ck["optimizers"].setdefault(k, v)
    for k, v in (extra_schedulers or {}).items():
        ck["schedulers"].setdefault(k, v)
    return ck


def optimizer_state_from_reference(ckpt: Mapping) -> Dict[str, dict]:
    """The "optimizers" entry of a reference checkpoint as training.FusedAdam.load_state_dict takes it."""
    out = {}
    for group, ours in GROUP_MAP.items():
        sd = ckpt.get("optimizers", {}).get(group)
        if sd is None:
            continue
        pg = sd["param_groups"][0]
        st = sd["state"].get(0) or next(iter(sd["state"].values()), None)
        if st is None:
            continue
        out[ours] = dict(step=int(float(st["step"])), lr=float(pg["lr"]), lr_init=float(pg.get("initial_lr", pg["lr"])),
                         exp_avg=st["exp_avg"], exp_avg_sq=st["exp_avg_sq"])
    return out


def save_training_checkpoint(path: str, step: int, params, optimizer=None, up_projection=None, extra_pipeline=None,
                             extra_optimizers=None, extra_schedulers=None) -> None:
    torch.save(trainer_checkpoint(step, params, optimizer, up_projection, extra_pipeline, extra_optimizers,
                                  extra_schedulers), path)


def load_reference_checkpoint(path: str, map_location="cpu") -> Dict[str, torch.Tensor]:
    return params_from_reference(torch.load(path, map_location=map_location, weights_only=False))
