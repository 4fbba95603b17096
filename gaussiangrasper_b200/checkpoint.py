"""Reading / writing the Gaussian parameters of a reference checkpoint (formats row).

The reference trainer saves `{"step", "pipeline": pipeline.state_dict(), "optimizers", ...}` with torch.save
(nerfstudio/engine/trainer.py:427-456); the Gaussian model's parameters sit in the pipeline state dict as
`_model.means`, `_model.scales`, `_model.quats`, `_model.opacities`, `_model.colors_all`, `_model.feature`
(nerfstudio/models/gaussian_splatting.py:255-281, load_state_dict :300-312).  This module maps them to the
names `render_views` takes and back, so a scene trained with the reference can be rendered here and vice
versa.  Pure host code (the reference's is Python too); no GPU involved.
"""
from __future__ import annotations

from typing import Dict, Mapping

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


def reference_state_dict(params: Mapping[str, torch.Tensor], prefix: str = "_model.") -> Dict[str, torch.Tensor]:
    """The inverse: entries a reference `GaussianSplattingModel.load_state_dict` accepts (:300-312)."""
    out = {}
    for ref_name, ours in NAME_MAP.items():
        t = params[ours].detach().to(torch.float32)
        if ref_name == "opacities":
            t = t.reshape(-1, 1)
        out[prefix + ref_name] = t.contiguous()
    return out


def load_reference_checkpoint(path: str, map_location="cpu") -> Dict[str, torch.Tensor]:
    return params_from_reference(torch.load(path, map_location=map_location, weights_only=False))
