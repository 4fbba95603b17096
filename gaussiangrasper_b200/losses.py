"""The terms of GaussianSplattingModel.get_loss_dict (nerfstudio/models/gaussian_splatting.py:841-933) that act on
the blended image of `render_views` and on the Gaussian parameters, each as one fused value-and-gradient launch
(csrc/loss.cu).  training.pixel_loss / training.ssim_loss give `main_loss`; this module adds

    depth_loss, normal_loss  geom_loss            (:876-880)
    feature_loss             contrastive_feature_loss over sampled pixel pairs (:905-912)
    up_loss                  up_loss: CLIP up-projection MLP on sampled points vs the ground-truth features (:913-914)
    sh_reg, scale_reg        param_regs           (:917-925)

All functions work on CUDA tensors only and return (loss tensor, gradient) -- continue with `image.backward(grad)`.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence, Tuple

import torch

from . import _lib

_ws = {}


def _workspace(dev) -> torch.Tensor:
    key = (dev.type, dev.index)
    ws = _ws.get(key)
    if ws is None:
        ws = _ws[key] = torch.zeros(int(_lib.load().gg_loss_workspace_bytes()), dtype=torch.uint8, device=dev)
    return ws


@torch.no_grad()
def geom_loss(image: torch.Tensor, gt_depth: torch.Tensor, gt_normal: torch.Tensor, depth_mask: torch.Tensor,
              w_depth: float = 1.0, w_normal: float = 1.0, grad: Optional[torch.Tensor] = None):
    """depth_loss and normal_loss of one or more blended images.

    image [..., CP] (render_views' `image`: depth = channel 3, normal = channels 4..6); gt_depth [...] or [..., 1];
    gt_normal [..., 3] (unit length, as the model normalises it :859); depth_mask [...] bool (gt_depth > 0.05 and
    valid, :861-871).  grad (optional, like image): the gradient of channels 3..6 is ADDED to it; otherwise a
    zero-filled tensor like image is returned with those channels set.
    Returns (loss [2] = (w_depth * depth_loss, w_normal * normal_loss), grad)."""
    dev = _lib.require_cuda(image, gt_depth, gt_normal, depth_mask, grad)
    img = _lib.f32c(image.detach())
    stride = int(img.shape[-1])
    n_pix = img.numel() // stride
    gd, gn = _lib.f32c(gt_depth.detach()).reshape(-1), _lib.f32c(gt_normal.detach()).reshape(-1, 3)
    m = depth_mask.to(torch.uint8).contiguous().reshape(-1)
    if gd.numel() != n_pix or gn.shape[0] != n_pix or m.numel() != n_pix:
        raise ValueError("gt_depth / gt_normal / depth_mask must have one entry per pixel of image")
    accumulate = grad is not None
    if accumulate:
        if grad.shape != image.shape or not grad.is_contiguous() or grad.dtype != torch.float32:
            raise ValueError("grad must be a contiguous fp32 tensor shaped like image")
    else:
        grad = torch.zeros_like(img)
    count = m.sum(dtype=torch.int32).reshape(1)
    loss = torch.empty(2, dtype=torch.float32, device=dev)
    ws = _workspace(dev)
    with _lib.device_guard(dev):
        _lib.call("gg_geom_loss", n_pix, stride, img.data_ptr(), gd.data_ptr(), gn.data_ptr(), m.data_ptr(),
                  count.data_ptr(), float(w_depth), float(w_normal), grad.data_ptr(), stride, 1 if accumulate else 0,
                  loss.data_ptr(), ws.data_ptr(), ws.numel(), _lib.stream_ptr(dev))
    return loss, grad


@torch.no_grad()
def cosine_rows_loss(a: torch.Tensor, b: torch.Tensor, a_index: Optional[torch.Tensor] = None,
                     b_index: Optional[torch.Tensor] = None, weight: Optional[torch.Tensor] = None,
                     grad_a: Optional[torch.Tensor] = None, grad_b: Optional[torch.Tensor] = None,
                     dim: Optional[int] = None, loss: Optional[torch.Tensor] = None):
    """sum_k w_k (1 - cos(a[ia_k], b[ib_k])) over rows of 2-D (strided) fp32 tensors; gradients are ADDED to
    grad_a / grad_b (tensors laid out like a / b).  `dim` restricts the rows to their first `dim` columns.
    With `loss` ([1]) given the value is added to it.  Returns the loss tensor [1]."""
    dev = _lib.require_cuda(a, b, a_index, b_index, weight, grad_a, grad_b, loss)
    for t in (a, b, grad_a, grad_b):
        if t is not None and (t.dim() != 2 or t.dtype != torch.float32 or t.stride(1) != 1):
            raise ValueError("operands must be 2-D fp32 tensors with unit column stride")
    dim = int(dim if dim is not None else a.shape[1])
    if dim > a.shape[1] or dim > b.shape[1]:
        raise ValueError("dim exceeds the row length")
    ia = a_index.to(torch.int64).contiguous() if a_index is not None else None
    ib = b_index.to(torch.int64).contiguous() if b_index is not None else None
    k = int(ia.numel() if ia is not None else (ib.numel() if ib is not None else a.shape[0]))
    if ia is None and ib is None and a.shape[0] != b.shape[0]:
        raise ValueError("a and b need the same number of rows")
    w = _lib.f32c(weight).reshape(-1) if weight is not None else None
    if w is not None and w.numel() != k:
        raise ValueError("one weight per row pair")
    for g, t in ((grad_a, a), (grad_b, b)):
        if g is not None and g.shape != t.shape:
            raise ValueError("gradients must be shaped like their operands")
    accumulate = loss is not None
    if loss is None:
        loss = torch.empty(1, dtype=torch.float32, device=dev)
    ws = _workspace(dev)
    with _lib.device_guard(dev):
        _lib.call("gg_cosine_rows_loss", k, dim, a.data_ptr(), int(a.stride(0)), _lib.ptr(ia), b.data_ptr(),
                  int(b.stride(0)), _lib.ptr(ib), _lib.ptr(w), _lib.ptr(grad_a),
                  int(grad_a.stride(0)) if grad_a is not None else 0, _lib.ptr(grad_b),
                  int(grad_b.stride(0)) if grad_b is not None else 0, loss.data_ptr(), 1 if accumulate else 0,
                  ws.data_ptr(), ws.numel(), _lib.stream_ptr(dev))
    return loss


def sampling_in_mask(mask: torch.Tensor, sample_num: int, generator: Optional[torch.Generator] = None) -> torch.Tensor:
    """gaussian_splatting.py:120-132: up to sample_num // (segments) pixels (row, col) of every segment id > -1."""
    mask = mask.detach()
    nums = torch.unique(mask)
    points_num = sample_num // max(len(nums) - 1, 1)
    out = []
    for v in nums.tolist():
        if v > -1:
            x, y = torch.where(mask == v)
            k = min(points_num, x.shape[0])
            ids = torch.randperm(x.shape[0], generator=generator, device=x.device if generator is None else generator.device)[:k].to(x.device)
            out.append(torch.stack([x[ids], y[ids]], dim=1))
    return torch.cat(out) if out else torch.zeros((0, 2), dtype=torch.int64, device=mask.device)


def sampling_pairs_in_mask(mask: torch.Tensor, sample_num: int, generator: Optional[torch.Generator] = None):
    """gaussian_splatting.py:134-148: per segment id > -1, two independent random orders of its pixels, paired."""
    mask = mask.detach()
    pairs = []
    for v in torch.unique(mask).tolist():
        if v > -1:
            x, y = torch.where(mask == v)
            k = min(sample_num, x.shape[0])
            gdev = x.device if generator is None else generator.device
            i1 = torch.randperm(x.shape[0], generator=generator, device=gdev)[:k].to(x.device)
            i2 = torch.randperm(x.shape[0], generator=generator, device=gdev)[:k].to(x.device)
            pairs.append([torch.stack([x[i1], y[i1]], dim=1), torch.stack([x[i2], y[i2]], dim=1)])
    return pairs


def _resized(t_hwc: torch.Tensor, size: Tuple[int, int], mode: str) -> torch.Tensor:
    """[H, W, C] -> [h, w, C] with F.interpolate's `mode` (half-pixel centres, no anti-aliasing)."""
    return torch.nn.functional.interpolate(t_hwc.permute(2, 0, 1)[None], size=size, mode=mode)[0].permute(1, 2, 0)


def prepare_targets(batch: Dict[str, torch.Tensor], downscale: int = 1) -> Dict[str, torch.Tensor]:
    """The supervision of one view as the loss functions of this module take it, from the reference's batch
    (get_loss_dict, gaussian_splatting.py:849-875): `image` [H,W,3], `normal` [H,W,3], `depth` [H,W,1], `sam_mask`
    [H,W] segment ids, `valid_mask` [H,W], `feature` [H,W,F].  With downscale d > 1 (the resolution warm-up,
    scenes.downscale_factor) everything is brought to [H // d, W // d]: image, normal and depth bilinearly, the masks,
    the depth validity (depth > 0.05, decided BEFORE the resize) and the features by nearest neighbour.

    Returns channel-last tensors: image [h,w,3]; normal [h,w,3] (unit length); depth [h,w]; depth_mask [h,w] bool
    (valid depth and valid pixel); valid [h,w] bool; segments [h,w] float (-1 on invalid pixels: what the samplers
    take); feature [h,w,F]."""
    F = torch.nn.functional
    image = batch["image"]
    size = (image.shape[0] // downscale, image.shape[1] // downscale) if downscale > 1 else tuple(image.shape[:2])
    if downscale > 1:
        image = _resized(image, size, "bilinear")
    normal = F.normalize(_resized(batch["normal"], size, "bilinear"), dim=-1)
    depth_full = batch["depth"].reshape(batch["depth"].shape[0], batch["depth"].shape[1], 1)
    depth_ok = _resized((depth_full > 0.05).to(depth_full.dtype), size, "nearest")[..., 0]
    depth = _resized(depth_full, size, "bilinear")[..., 0]
    valid = _resized(batch["valid_mask"].float()[..., None], size, "nearest")[..., 0] > 0
    segments = _resized(batch["sam_mask"].float()[..., None], size, "nearest")[..., 0].clone()
    segments[~valid] = -1.0
    feature = _resized(batch["feature"].float(), size, "nearest")
    return dict(image=image, normal=normal, depth=depth, depth_mask=(depth_ok > 0) & valid, valid=valid,
                segments=segments, feature=feature)


def _pixel_rows(image: torch.Tensor, c0: int, D: int) -> Tuple[torch.Tensor, int]:
    """[pixels, D] strided view of the feature channels of a [H, W, CP] image."""
    H, W, CP = image.shape
    return image.reshape(H * W, CP)[:, c0:c0 + D], W


@torch.no_grad()
def contrastive_feature_loss(image: torch.Tensor, pairs: Sequence, grad: torch.Tensor, feature_channel0: int = 7,
                             feature_dim: Optional[int] = None, weight: float = 1.0):
    """feature_loss of :905-912: mean over the segments of cosine_similarity_loss between the features at the two
    pixel lists of each segment.  image, grad: [H, W, CP] (one view); pairs: list of [pixels_1 [K,2], pixels_2 [K,2]]
    (row, col) as sampling_pairs_in_mask returns.  The gradient is ADDED to grad.  Returns loss [1]."""
    H, W, CP = image.shape
    D = int(feature_dim if feature_dim is not None else CP - feature_channel0)
    rows, _ = _pixel_rows(_lib.f32c(image.detach()), feature_channel0, D)
    grows, _ = _pixel_rows(grad, feature_channel0, D)
    pairs = [p for p in pairs if p[0].shape[0] > 0]
    if not pairs:
        return torch.zeros(1, dtype=torch.float32, device=image.device)
    i1 = torch.cat([p[0][:, 0] * W + p[0][:, 1] for p in pairs])
    i2 = torch.cat([p[1][:, 0] * W + p[1][:, 1] for p in pairs])
    w = torch.cat([torch.full((p[0].shape[0],), weight / (len(pairs) * p[0].shape[0]), dtype=torch.float32,
                              device=image.device) for p in pairs])
    return cosine_rows_loss(rows, rows, i1, i2, w, grows, grows, dim=D)


class UpProjection(torch.nn.Module):
    """The CLIP up-projection `fea_up = MLP(feature_dim, 512, [128])` of the reference model
    (gaussian_splatting.py:198-213, :294): Linear(D,128) - ReLU - Linear(128,512); parameter names match the
    reference's `fea_up.layers.{0,2}.{weight,bias}` so its checkpoints load."""

    def __init__(self, in_dim: int = 32, out_dim: int = 512, hidden: int = 128):
        super().__init__()
        self.layers = torch.nn.Sequential(torch.nn.Linear(in_dim, hidden), torch.nn.ReLU(), torch.nn.Linear(hidden, out_dim))

    def forward(self, x):
        return self.layers(x)


def up_loss(image: torch.Tensor, points: torch.Tensor, gt_features: torch.Tensor, mlp: torch.nn.Module,
            grad: torch.Tensor, feature_channel0: int = 7, feature_dim: Optional[int] = None, weight: float = 1.0):
    """up_loss of :913-914: cosine_similarity_loss(fea_up(feature[points]), gt_fea[:, points]).
    image, grad [H, W, CP]; points [K, 2] (row, col); gt_features [H, W, F] (channel-last CLIP features).
    The 1000-point MLP runs through torch (its parameters receive `.grad`), the cosine loss and its gradient are
    one kernel; the gradient w.r.t. the sampled feature pixels is ADDED to grad.  Returns loss [1]."""
    H, W, CP = image.shape
    D = int(feature_dim if feature_dim is not None else CP - feature_channel0)
    K = int(points.shape[0])
    if K == 0:
        return torch.zeros(1, dtype=torch.float32, device=image.device)
    idx = points[:, 0] * W + points[:, 1]
    rows, _ = _pixel_rows(image.detach(), feature_channel0, D)
    x = rows[idx].clone().requires_grad_(True)
    with torch.enable_grad():
        up = mlp(x)
    gt = _lib.f32c(gt_features.detach()).reshape(H * W, -1)
    g_up = torch.zeros_like(up)
    w = torch.full((K,), weight / K, dtype=torch.float32, device=image.device)
    loss = cosine_rows_loss(up.detach().contiguous(), gt, None, idx, w, g_up, None)
    up.backward(g_up)
    grows, _ = _pixel_rows(grad, feature_channel0, D)
    grows.index_put_((idx,), x.grad, accumulate=True)
    return loss


@torch.no_grad()
def param_regs(sh_coeffs: torch.Tensor, log_scales: torch.Tensor, max_gauss_ratio: float = 10.0, w_sh: float = 1.0,
               w_scale: float = 1.0, v_sh_coeffs: Optional[torch.Tensor] = None,
               v_log_scales: Optional[torch.Tensor] = None):
    """sh_reg and scale_reg of :917-925 (the model applies them every 10th step); their gradients are ADDED to
    v_sh_coeffs / v_log_scales (e.g. the views of a GradientBucket).  Returns loss [2] = (sh_reg, scale_reg)."""
    dev = _lib.require_cuda(sh_coeffs, log_scales, v_sh_coeffs, v_log_scales)
    sh, ls = _lib.f32c(sh_coeffs.detach()), _lib.f32c(log_scales.detach())
    n, nb = sh.shape[0], sh.shape[1]
    for g, t in ((v_sh_coeffs, sh), (v_log_scales, ls)):
        if g is not None and (g.numel() != t.numel() or not g.is_contiguous() or g.dtype != torch.float32):
            raise ValueError("gradient buffers must be contiguous fp32 tensors sized like their parameters")
    loss = torch.empty(2, dtype=torch.float32, device=dev)
    ws = _workspace(dev)
    with _lib.device_guard(dev):
        _lib.call("gg_param_regs", n, nb, sh.data_ptr(), ls.data_ptr(), float(max_gauss_ratio), float(w_sh), float(w_scale),
                  _lib.ptr(v_sh_coeffs), _lib.ptr(v_log_scales), loss.data_ptr(), ws.data_ptr(), ws.numel(),
                  _lib.stream_ptr(dev))
    return loss


_packed_cache = {}


@torch.no_grad()
def up_project(features: torch.Tensor, mlp: torch.nn.Module, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """`model.fea_up(outputs["feature"])` for a whole feature map (base_pipeline.py:408): [..., 32] -> [..., 512]
    through the fused tcgen05 / TMEM kernel (csrc/mlp.cu: 3xTF32 split, fp32 accumulation, fp32-equivalent results).

    features: fp32 CUDA tensor whose last dimension (32) has unit stride and whose rows are equally spaced -- e.g.
    `render_views(...)["feature"]`, a channel slice of the blended image: no copy is made.  mlp: UpProjection (or
    any module with `layers.0` = Linear(32,128), `layers.2` = Linear(128,512)).  Inference only (no autograd)."""
    l0, l2 = mlp.layers[0], mlp.layers[2]
    if tuple(l0.weight.shape) != (128, 32) or tuple(l2.weight.shape) != (512, 128):
        raise ValueError("up_project implements the reference's MLP(32 -> 128 -> 512)")
    dev = _lib.require_cuda(features, l0.weight, l2.weight, out)
    if features.dtype != torch.float32 or features.shape[-1] != 32 or features.stride(-1) != 1:
        raise ValueError("features must be fp32 [..., 32] with unit stride in the last dimension")
    lead = features.shape[:-1]
    n_rows = 1
    for d in lead:
        n_rows *= int(d)
    # equally spaced rows: every leading stride must be the flattened multiple of the innermost leading stride
    row_stride = int(features.stride(-2)) if features.dim() >= 2 else 32
    expect = row_stride
    for d in range(features.dim() - 2, -1, -1):
        if features.shape[d] != 1 and features.stride(d) != expect:
            features = features.contiguous()
            row_stride = 32
            break
        expect *= int(features.shape[d])
    key = (dev.index, l0.weight.data_ptr(), l0.weight._version, l2.weight.data_ptr(), l2.weight._version)
    packed = _packed_cache.get(key)
    lib = _lib.load()
    with _lib.device_guard(dev):
        if packed is None:
            _packed_cache.clear()
            packed = torch.empty(int(lib.gg_mlp_packed_floats()), dtype=torch.float32, device=dev)
            _lib.call("gg_mlp_pack_weights", _lib.f32c(l0.weight.detach()).data_ptr(), _lib.f32c(l2.weight.detach()).data_ptr(),
                      packed.data_ptr(), _lib.stream_ptr(dev))
            _packed_cache[key] = packed
        if out is None:
            out = torch.empty(tuple(lead) + (512,), dtype=torch.float32, device=dev)
        elif out.numel() != n_rows * 512 or not out.is_contiguous() or out.dtype != torch.float32:
            raise ValueError("out must be a contiguous fp32 tensor of [..., 512]")
        b1, b2 = _lib.f32c(l0.bias.detach()), _lib.f32c(l2.bias.detach())
        _lib.call("gg_mlp_up", n_rows, features.data_ptr(), row_stride, packed.data_ptr(), b1.data_ptr(), b2.data_ptr(),
                  out.data_ptr(), _lib.stream_ptr(dev))
    return out
