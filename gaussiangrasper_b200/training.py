"""The two per-Gaussian streaming steps directly behind the rasterizer's backward in a training
iteration ("next rows"): a fused Adam over the flat all-reduced gradient buffer and the
densification statistics of GaussianSplattingModel.after_train."""
from __future__ import annotations

import ctypes as C
from typing import Dict, Optional

import torch

from . import _lib
from .distributed import GradientBucket

# learning rates of the reference's optimizer groups (nerfstudio/configs/method_configs.py:618-664)
REFERENCE_LRS = dict(means=1.6e-4, log_scales=5e-3, quats=1e-3, opacity_logit=5e-2, sh_coeffs=5e-4, features=5e-4)


class FusedAdam:
    """torch.optim.Adam semantics (betas (0.9, 0.999), eps 1e-15 as the reference configures it, no weight
    decay) for all Gaussian parameters in ONE kernel over the flat gradient bucket."""

    def __init__(self, params: Dict[str, torch.Tensor], bucket: Optional[GradientBucket] = None,
                 lrs: Optional[Dict[str, float]] = None, betas=(0.9, 0.999), eps: float = 1e-15):
        self.params = params
        self.bucket = bucket if bucket is not None else GradientBucket(params)
        self.lrs = dict(REFERENCE_LRS if lrs is None else lrs)
        self.betas, self.eps, self.t = betas, eps, 0
        self.exp_avg = torch.zeros_like(self.bucket.flat)
        self.exp_avg_sq = torch.zeros_like(self.bucket.flat)
        for k in self.bucket.names:
            p = params[k]
            if not (p.is_cuda and p.is_contiguous() and p.dtype == torch.float32):
                raise _lib.GGError(f"parameter {k} must be a contiguous fp32 CUDA tensor")

    @torch.no_grad()
    def step(self, lrs: Optional[Dict[str, float]] = None) -> None:
        """Apply one update from the gradients currently in the bucket (already summed over ranks)."""
        if lrs:
            self.lrs.update(lrs)
        self.t += 1
        b = self.bucket
        names = b.names
        n = len(names)
        ptrs = (C.c_void_p * n)(*[self.params[k].data_ptr() for k in names])
        offs = (C.c_longlong * n)(*[b.offsets[k] for k in names])
        cnts = (C.c_longlong * n)(*[b.sizes[k] for k in names])
        lr = (C.c_float * n)(*[float(self.lrs[k]) for k in names])
        dev = b.flat.device
        with _lib.device_guard(dev):
            _lib.call("gg_adam_step", n, ptrs, offs, cnts, lr, b.flat.data_ptr(), self.exp_avg.data_ptr(),
                      self.exp_avg_sq.data_ptr(), float(self.betas[0]), float(self.betas[1]), float(self.eps), self.t,
                      _lib.stream_ptr(dev))


class DensifyStats:
    """xys_grad_norm / vis_counts / max_2Dsize of gaussian_splatting.py:373-393, fed from the holder of
    render_views after backward (v_geo columns 0..1 are d loss / d xys)."""

    def __init__(self, n: int, device):
        self.n = n
        self.xys_grad_norm = torch.zeros(n, dtype=torch.float32, device=device)
        self.vis_counts = torch.zeros(n, dtype=torch.float32, device=device)
        self.max_2Dsize = torch.zeros(n, dtype=torch.float32, device=device)
        self.first = True

    @torch.no_grad()
    def update(self, v_geo: torch.Tensor, radii: torch.Tensor, img_height: int, img_width: int) -> None:
        n_views = radii.numel() // self.n
        dev = v_geo.device
        radii = radii.contiguous()
        with _lib.device_guard(dev):
            _lib.call("gg_densify_stats", self.n, n_views, v_geo.data_ptr(), radii.data_ptr(), int(img_height),
                      int(img_width), 1 if self.first else 0, self.xys_grad_norm.data_ptr(), self.vis_counts.data_ptr(),
                      self.max_2Dsize.data_ptr(), _lib.stream_ptr(dev))
        self.first = False
