"""CPU restatement (torch) of the image losses of GaussianSplattingModel.get_loss_dict.

TEST INFRASTRUCTURE ONLY.  Every term but SSIM is PINNED TO THE REFERENCE: tests/golden/ref_losses_small.npz holds the
values and gradients of the reference's own get_loss_dict (tests/golden/make_reference_golden.py) and
tests/test_reference_golden_cpu.py checks this file against them (values to 2e-6, gradients to 2e-5 of the largest
entry).  SSIM: PARITY UNPINNED -- `pytorch_msssim` (requirements.txt of the reference) is not
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


# ---------------------------------------------------------------------------------------------
# the other terms of get_loss_dict, restated line by line (differentiable; gradients by autograd)
# ---------------------------------------------------------------------------------------------
def cosine_similarity_loss(e1: torch.Tensor, e2: torch.Tensor) -> torch.Tensor:
    """gaussian_splatting.py:113-118: embeddings are [dim, K] (normalised and summed over dim 0)."""
    e1 = F.normalize(e1, dim=0)
    e2 = F.normalize(e2, dim=0)
    return 1 - torch.sum(e1 * e2, dim=0).mean()


def geom_losses(normal_hw3: torch.Tensor, depth_hw1: torch.Tensor, gt_normal_3hw: torch.Tensor,
                gt_depth_1hw: torch.Tensor, depth_mask_1hw: torch.Tensor):
    """:876-880 on outputs["normal"] [H,W,3], outputs["depth"] [H,W,1]; returns (depth_loss, normal_loss)."""
    normal = normal_hw3.permute(2, 0, 1)
    depth = depth_hw1.permute(2, 0, 1)
    m = depth_mask_1hw
    normal_loss = 0.5 * F.mse_loss(normal[:, m[0]], gt_normal_3hw[:, m[0]], reduction='mean') + \
        0.5 * cosine_similarity_loss(normal[:, m[0]], gt_normal_3hw[:, m[0]])
    depth_loss = F.l1_loss(depth[m], gt_depth_1hw[m], reduction='mean')
    return depth_loss, normal_loss


def feature_loss(feature_hwd: torch.Tensor, selected_pairs) -> torch.Tensor:
    """:905-912."""
    fea_loss = 0
    for i in range(len(selected_pairs)):
        f1 = feature_hwd[selected_pairs[i][0][:, 0], selected_pairs[i][0][:, 1]]
        f2 = feature_hwd[selected_pairs[i][1][:, 0], selected_pairs[i][1][:, 1]]
        fea_loss += cosine_similarity_loss(f1.permute(1, 0), f2.permute(1, 0))
    return fea_loss / len(selected_pairs)


def up_loss(feature_hwd: torch.Tensor, selected_points: torch.Tensor, gt_fea_fhw: torch.Tensor, mlp) -> torch.Tensor:
    """:913-914."""
    fea_up = mlp(feature_hwd[selected_points[:, 0], selected_points[:, 1], :]).permute(1, 0)
    return cosine_similarity_loss(fea_up, gt_fea_fhw[:, selected_points[:, 0], selected_points[:, 1]])


def regs(colors_all: torch.Tensor, scales: torch.Tensor, max_gauss_ratio: float = 10.0):
    """:917-923; returns (sh_reg, scale_reg)."""
    sh_reg = colors_all[:, 1:, :].norm(dim=1).mean()
    scale_exp = torch.exp(scales)
    scale_reg = torch.maximum(scale_exp.amax(dim=-1) / scale_exp.amin(dim=-1),
                              torch.tensor(max_gauss_ratio, dtype=scales.dtype)) - max_gauss_ratio
    return sh_reg, 0.1 * scale_reg.mean()


class MLP(torch.nn.Module):
    """:198-213 (in_dim -> hidden_list -> out_dim with ReLU between)."""

    def __init__(self, in_dim=8, out_dim=512, hidden_list=(128,)):
        super().__init__()
        layers = []
        lastv = in_dim
        for hidden in hidden_list:
            layers.append(torch.nn.Linear(lastv, hidden))
            layers.append(torch.nn.ReLU())
            lastv = hidden
        layers.append(torch.nn.Linear(lastv, out_dim))
        self.layers = torch.nn.Sequential(*layers)

    def forward(self, x):
        return self.layers(x)
