"""CPU restatement (torch) of the image losses of GaussianSplattingModel.get_loss_dict.

TEST INFRASTRUCTURE ONLY.  PARITY UNPINNED: `pytorch_msssim` (requirements.txt of the reference) is not
installed here and the reference pins no vectors; `ssim` restates its published algorithm
(pytorch_msssim/ssim.py: _fspecial_gauss_1d, gaussian_filter, _ssim, ssim with size_average=True,
nonnegative_ssim=False, K=(0.01, 0.03), win_size 11, win_sigma 1.5) as the reference configures it
(nerfstudio/models/gaussian_splatting.py:284, :885).  Differentiable: gradients come from autograd.
"""
import torch
import torch.nn.functional as F


def _gauss_1d(size: int = 11, sigma: float = 1.5) -> torch.Tensor:
    coords = torch.arange(size, dtype=torch.float32) - size // 2
    g = torch.exp(-(coords ** 2) / (2 * sigma ** 2))
    return g / g.sum()


def _filter(x: torch.Tensor, win: torch.Tensor) -> torch.Tensor:
    """Separable, unpadded Gaussian blur of [B,C,H,W], H first then W (gaussian_filter)."""
    c = x.shape[1]
    w = win.to(x.dtype)
    x = F.conv2d(x, w.view(1, 1, -1, 1).repeat(c, 1, 1, 1), groups=c)
    return F.conv2d(x, w.view(1, 1, 1, -1).repeat(c, 1, 1, 1), groups=c)


def ssim(x: torch.Tensor, y: torch.Tensor, data_range: float = 1.0) -> torch.Tensor:
    """x, y: [B,C,H,W] -> scalar mean SSIM."""
    win = _gauss_1d()
    c1, c2 = (0.01 * data_range) ** 2, (0.03 * data_range) ** 2
    mu1, mu2 = _filter(x, win), _filter(y, win)
    s11 = _filter(x * x, win) - mu1 * mu1
    s22 = _filter(y * y, win) - mu2 * mu2
    s12 = _filter(x * y, win) - mu1 * mu2
    cs = (2 * s12 + c2) / (s11 + s22 + c2)
    m = ((2 * mu1 * mu2 + c1) / (mu1 * mu1 + mu2 * mu2 + c1)) * cs
    return torch.flatten(m, 2).mean(-1).mean()


def main_loss(pred_rgb: torch.Tensor, gt_rgb: torch.Tensor, ssim_lambda: float = 0.2) -> torch.Tensor:
    """(1 - lambda) * L1 + lambda * (1 - SSIM) on [H,W,3] images (gaussian_splatting.py:882-885, :931)."""
    l1 = (gt_rgb - pred_rgb).abs().mean()
    s = 1 - ssim(gt_rgb.permute(2, 0, 1)[None], pred_rgb.permute(2, 0, 1)[None])
    return (1 - ssim_lambda) * l1 + ssim_lambda * s
