"""Vectorised, differentiable PyTorch-CPU restatement of the gsplat 0.1.0 path
(the role gsplat/_torch_impl.py plays upstream).

TEST INFRASTRUCTURE ONLY.  PARITY UNPINNED (see oracle/gg_oracle.c): gsplat 0.1.0
is not vendored in /root/reference and not installable here; this file restates
SURVEY.md Appendix A (A1..A11) and is anchored on the reference call sites
nerfstudio/models/gaussian_splatting.py:699-784.  It is used (a) as the fp64
autograd checker for gradients, (b) as the "port" CPU baseline timed by bench.py.

Everything works in the dtype of its inputs (fp32 or fp64).
"""
from __future__ import annotations

import math
from typing import Optional, Tuple

import torch

TILE = 16

SH_C0 = 0.28209479177387814
SH_C1 = 0.4886025119029199
SH_C2 = (1.0925484305920792, -1.0925484305920792, 0.31539156525252005, -1.0925484305920792, 0.5462742152960396)
SH_C3 = (-0.5900435899266435, 2.890611442640554, -0.4570457994644658, 0.3731763325901154,
         -0.4570457994644658, 1.445305721320277, -0.5900435899266435)
SH_C4 = (2.5033429417967046, -1.7701307697799304, 0.9461746957575601, -0.6690465435572892,
         0.10578554691520431, -0.6690465435572892, 0.47308734787878004, -1.7701307697799304,
         0.6258357354491761)


def num_sh_bases(degree: int) -> int:
    return {0: 1, 1: 4, 2: 9, 3: 16}.get(degree, 25)


def quat_to_rotmat(quat: torch.Tensor) -> torch.Tensor:
    """A2; gsplat._torch_impl.quat_to_rotmat (imported at gaussian_splatting.py:46)."""
    q = quat / quat.norm(dim=-1, keepdim=True)
    w, x, y, z = q.unbind(-1)
    R = torch.stack([
        1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y),
        2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x),
        2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)], dim=-1)
    return R.reshape(quat.shape[:-1] + (3, 3))


def _trunc_sat_i32(x: torch.Tensor) -> torch.Tensor:
    x = torch.nan_to_num(x, nan=0.0)
    return torch.clamp(torch.trunc(x), -2147483648.0, 2147483647.0).to(torch.int64)


def tile_bbox(xys, radius, tile_bounds):
    tc = xys / TILE
    tr = (radius / TILE)[:, None]
    bound = torch.tensor([tile_bounds[0], tile_bounds[1]], dtype=torch.int64)
    tmin = torch.minimum(torch.clamp(_trunc_sat_i32(tc - tr), min=0), bound)
    tmax = torch.minimum(torch.clamp(_trunc_sat_i32(tc + tr + 1), min=0), bound)
    return tmin, tmax


def project_gaussians(means3d, scales, glob_scale, quats, viewmat, fullmat, fx, fy, cx, cy, H, W, tile_bounds,
                      clip_thresh=0.01):
    """A1..A6.  Returns (xys, depths, radii, conics, num_tiles_hit, cov3d); culled rows are zero."""
    dt = means3d.dtype
    viewmat = viewmat.to(dt)
    fullmat = fullmat.to(dt)
    Wm = viewmat[:3, :3]
    t = means3d @ Wm.T + viewmat[:3, 3]
    depth = t[:, 2]
    keep = depth > clip_thresh
    R = quat_to_rotmat(quats)
    M = R * (glob_scale * scales)[:, None, :]
    cov3d_full = M @ M.transpose(1, 2)
    cov3d = torch.stack([cov3d_full[:, 0, 0], cov3d_full[:, 0, 1], cov3d_full[:, 0, 2],
                         cov3d_full[:, 1, 1], cov3d_full[:, 1, 2], cov3d_full[:, 2, 2]], dim=-1)
    tan_fovx, tan_fovy = 0.5 * W / fx, 0.5 * H / fy
    limx, limy = 1.3 * tan_fovx, 1.3 * tan_fovy
    safe_z = torch.where(keep, depth, torch.ones_like(depth))
    tx = safe_z * torch.clamp(t[:, 0] / safe_z, -limx, limx)
    ty = safe_z * torch.clamp(t[:, 1] / safe_z, -limy, limy)
    rz = 1.0 / safe_z
    rz2 = rz * rz
    zero = torch.zeros_like(rz)
    J = torch.stack([fx * rz, zero, -fx * tx * rz2, zero, fy * rz, -fy * ty * rz2], dim=-1).reshape(-1, 2, 3)
    T = J @ Wm
    cov2d = T @ cov3d_full @ T.transpose(1, 2)
    a = cov2d[:, 0, 0] + 0.3
    b = cov2d[:, 0, 1]
    c = cov2d[:, 1, 1] + 0.3
    det = a * c - b * b
    keep = keep & (det != 0)
    sdet = torch.where(det != 0, det, torch.ones_like(det))
    conic = torch.stack([c / sdet, -b / sdet, a / sdet], dim=-1)
    mid = 0.5 * (a + c)
    sq = torch.sqrt(torch.clamp(mid * mid - det, min=0.1))
    radius = torch.ceil(3.0 * torch.sqrt(torch.maximum(mid + sq, mid - sq))).detach()
    hom = means3d @ fullmat[:, :3].T + fullmat[:, 3]
    rw = 1.0 / (hom[:, 3] + 1e-6)
    ux = 0.5 * W * (hom[:, 0] * rw) + cx - 0.5
    uy = 0.5 * H * (hom[:, 1] * rw) + cy - 0.5
    xys = torch.stack([ux, uy], dim=-1)
    tmin, tmax = tile_bbox(xys.detach(), radius, tile_bounds)
    area = (tmax[:, 0] - tmin[:, 0]) * (tmax[:, 1] - tmin[:, 1])
    keep = keep & (area > 0)
    kf = keep.to(dt)
    nth = torch.where(keep, area, torch.zeros_like(area)).to(torch.int32)
    radii = torch.where(keep, _trunc_sat_i32(radius), torch.zeros_like(area)).to(torch.int32)
    z = lambda v: torch.where(keep[:, None] if v.dim() == 2 else keep, v, torch.zeros_like(v))
    return z(xys), z(depth), radii, z(conic), nth, torch.where((depth > clip_thresh)[:, None], cov3d, torch.zeros_like(cov3d))


def sh_basis(degrees_to_use: int, dirs: torch.Tensor) -> torch.Tensor:
    """A10 basis values [N, (deg+1)^2]."""
    n = dirs.shape[0]
    Y = [torch.full((n,), SH_C0, dtype=dirs.dtype)]
    if degrees_to_use >= 1:
        d = dirs / dirs.norm(dim=-1, keepdim=True)
        x, y, z = d.unbind(-1)
        Y += [SH_C1 * (-y), SH_C1 * z, SH_C1 * (-x)]
    if degrees_to_use >= 2:
        xx, yy, zz, xy, yz, xz = x * x, y * y, z * z, x * y, y * z, x * z
        Y += [SH_C2[0] * xy, SH_C2[1] * yz, SH_C2[2] * (2 * zz - xx - yy), SH_C2[3] * xz, SH_C2[4] * (xx - yy)]
    if degrees_to_use >= 3:
        Y += [SH_C3[0] * y * (3 * xx - yy), SH_C3[1] * xy * z, SH_C3[2] * y * (4 * zz - xx - yy),
              SH_C3[3] * z * (2 * zz - 3 * xx - 3 * yy), SH_C3[4] * x * (4 * zz - xx - yy),
              SH_C3[5] * z * (xx - yy), SH_C3[6] * x * (xx - 3 * yy)]
    if degrees_to_use >= 4:
        Y += [SH_C4[0] * xy * (xx - yy), SH_C4[1] * yz * (3 * xx - yy), SH_C4[2] * xy * (7 * zz - 1),
              SH_C4[3] * yz * (7 * zz - 3), SH_C4[4] * (zz * (35 * zz - 30) + 3), SH_C4[5] * xz * (7 * zz - 3),
              SH_C4[6] * (xx - yy) * (7 * zz - 1), SH_C4[7] * xz * (xx - 3 * yy),
              SH_C4[8] * (xx * (xx - 3 * yy) - yy * (3 * xx - yy))]
    return torch.stack(Y, dim=-1)


def spherical_harmonics(degrees_to_use: int, dirs: torch.Tensor, coeffs: torch.Tensor) -> torch.Tensor:
    """colour[N,3] = sum_b Y_b(dir) coeffs[N,b,3]; no gradient to dirs (as upstream)."""
    Y = sh_basis(degrees_to_use, dirs.detach())
    nb = Y.shape[1]
    return (Y[:, :, None] * coeffs[:, :nb, :]).sum(dim=1)


def bin_and_sort(xys, depths, radii, num_tiles_hit, tile_bounds):
    """A7/A8 with torch ops.  Returns (cum, keys_sorted[int64], ids_sorted[int32], tile_ranges[T,2] int32)."""
    n = xys.shape[0]
    tx_n, ty_n = int(tile_bounds[0]), int(tile_bounds[1])
    cum = torch.cumsum(num_tiles_hit.to(torch.int32), dim=0, dtype=torch.int32)
    tmin, tmax = tile_bbox(xys.detach().float(), radii.float(), tile_bounds)
    vis = radii > 0
    w = torch.where(vis, tmax[:, 0] - tmin[:, 0], torch.zeros_like(tmin[:, 0]))
    cnt = num_tiles_hit.to(torch.int64)
    ids = torch.repeat_interleave(torch.arange(n, dtype=torch.int64), cnt)
    start = torch.cumsum(cnt, 0) - cnt
    k = torch.arange(ids.numel(), dtype=torch.int64) - start[ids]
    wi = torch.clamp(w[ids], min=1)
    ty = tmin[ids, 1] + k // wi
    tx = tmin[ids, 0] + k % wi
    dbits = depths.detach().float().contiguous().view(torch.int32).to(torch.int64) & 0xFFFFFFFF
    keys = ((ty * tx_n + tx) << 32) | dbits[ids]
    keys_sorted, order = torch.sort(keys, stable=True)
    ids_sorted = ids[order].to(torch.int32)
    tiles = (keys_sorted >> 32)
    ranges = torch.zeros((tx_n * ty_n, 2), dtype=torch.int32)
    if keys_sorted.numel() > 0:
        m = keys_sorted.numel()
        first = torch.ones(m, dtype=torch.bool)
        first[1:] = tiles[1:] != tiles[:-1]
        last = torch.ones(m, dtype=torch.bool)
        last[:-1] = tiles[1:] != tiles[:-1]
        idx = torch.arange(m, dtype=torch.int32)
        ranges[tiles[first], 0] = idx[first]
        ranges[tiles[last], 1] = idx[last] + 1
    return cum, keys_sorted, ids_sorted, ranges


def _blend_tile(px, py, g, xys, conics, opac, colors, bg, chunk=2048):
    """A9 for one tile.  px,py: [P] pixel coords; g: [L] int64 gaussian ids in depth order."""
    dt = xys.dtype
    P = px.shape[0]
    out = torch.zeros((P, colors.shape[1]), dtype=dt)
    T = torch.ones((P,), dtype=dt)
    alive = torch.ones((P,), dtype=torch.bool)
    last = torch.zeros((P,), dtype=torch.int64)
    for s in range(0, g.numel(), chunk):
        gi = g[s:s + chunk]
        dx = xys[gi, 0][None, :] - px[:, None]
        dy = xys[gi, 1][None, :] - py[:, None]
        A, B, C = conics[gi, 0][None, :], conics[gi, 1][None, :], conics[gi, 2][None, :]
        sigma = 0.5 * (A * dx * dx + C * dy * dy) + B * dx * dy
        alpha = torch.clamp(opac[gi][None, :] * torch.exp(-sigma), max=0.999)
        valid = (sigma >= 0) & (alpha >= 1.0 / 255.0)
        a_eff = torch.where(valid, alpha, torch.zeros_like(alpha))
        om = 1.0 - a_eff
        cp = torch.cumprod(om, dim=1)
        T_incl = T[:, None] * cp
        live = (T_incl > 1e-4) & alive[:, None]
        T_excl = torch.cat([T[:, None], T_incl[:, :-1]], dim=1)
        w = a_eff * T_excl * live.to(dt)
        out = out + w @ colors[gi]
        contrib = live & valid
        pos = torch.arange(gi.numel(), dtype=torch.int64)[None, :] + (s + 1)
        last = torch.maximum(last, torch.where(contrib, pos, torch.zeros_like(pos)).amax(dim=1))
        T = T * torch.where(live, om, torch.ones_like(om)).prod(dim=1)
        alive = alive & live[:, -1]
        if not bool(alive.any()):
            break
    out = out + T[:, None] * bg[None, :]
    return out, T, last


def rasterize(xys, conics, opac, colors, ids_sorted, tile_ranges, H, W, bg, tiles=None):
    """Differentiable A9 over the whole image (or only the listed tile ids).

    Returns (out[H,W,C], final_T[H,W], final_idx[H,W] int32)."""
    dt = xys.dtype
    opac = opac.reshape(-1)
    ch = colors.shape[1]
    tiles_x = (W + TILE - 1) // TILE
    tiles_y = (H + TILE - 1) // TILE
    out = torch.zeros((H, W, ch), dtype=dt) + bg.to(dt) * 0  # keeps graph simple
    rows = [[None] * tiles_x for _ in range(tiles_y)]
    fT = torch.ones((H, W), dtype=dt)
    fidx = torch.zeros((H, W), dtype=torch.int32)
    ids64 = ids_sorted.to(torch.int64)
    tile_iter = range(tiles_x * tiles_y) if tiles is None else tiles
    pieces = {}
    for t in tile_iter:
        ty, tx = divmod(int(t), tiles_x)
        y0, x0 = ty * TILE, tx * TILE
        y1, x1 = min(y0 + TILE, H), min(x0 + TILE, W)
        yy, xx = torch.meshgrid(torch.arange(y0, y1), torch.arange(x0, x1), indexing="ij")
        px, py = xx.reshape(-1).to(dt), yy.reshape(-1).to(dt)
        s, e = int(tile_ranges[t, 0]), int(tile_ranges[t, 1])
        o, T, last = _blend_tile(px, py, ids64[s:e], xys, conics, opac, colors, bg.to(dt))
        pieces[(ty, tx)] = o.reshape(y1 - y0, x1 - x0, ch)
        fT[y0:y1, x0:x1] = T.detach().reshape(y1 - y0, x1 - x0)
        fidx[y0:y1, x0:x1] = (last + s).to(torch.int32).reshape(y1 - y0, x1 - x0)
    # assemble (tiles not rendered are background-free zeros)
    row_tensors = []
    for ty in range(tiles_y):
        y0 = ty * TILE
        y1 = min(y0 + TILE, H)
        cols = []
        for tx in range(tiles_x):
            x0 = tx * TILE
            x1 = min(x0 + TILE, W)
            cols.append(pieces.get((ty, tx), torch.zeros((y1 - y0, x1 - x0, ch), dtype=dt)))
        row_tensors.append(torch.cat(cols, dim=1))
    out = torch.cat(row_tensors, dim=0)
    return out, fT, fidx


def rasterize_literal(xys, conics, opac, colors, ids_sorted, tile_ranges, H, W, bg, crop):
    """A9 as the literal per-pixel, per-entry loop (the shape of gsplat's `_torch_impl` rasterizer: Python loops
    over pixels and intersections), on the pixel rectangle crop = (x0, y0, w, h).  Plain Python floats rounded to
    fp32 after every operation would cost more than the loop itself, so the arithmetic is fp64; it is used to
    validate the vectorised `rasterize` on small crops and as SURVEY 8d's CPU baseline (i).
    Returns (out [h, w, C] float64 tensor, pairs visited)."""
    x0, y0, w, h = crop
    X, Cn, O = xys.double().tolist(), conics.double().tolist(), opac.reshape(-1).double().tolist()
    Col, Bg = colors.double().tolist(), bg.double().tolist()
    ids, ranges = ids_sorted.tolist(), tile_ranges.tolist()
    ch = len(Bg)
    tiles_x = (W + TILE - 1) // TILE
    out = torch.zeros((h, w, ch), dtype=torch.float64)
    pairs = 0
    exp = math.exp
    for py in range(y0, y0 + h):
        for px in range(x0, x0 + w):
            s, e = ranges[(py // TILE) * tiles_x + px // TILE]
            acc = [0.0] * ch
            T = 1.0
            for k in range(s, e):
                pairs += 1
                g = ids[k]
                dx, dy = X[g][0] - px, X[g][1] - py
                A, B, C = Cn[g]
                sigma = 0.5 * (A * dx * dx + C * dy * dy) + B * dx * dy
                if sigma < 0.0:
                    continue
                alpha = min(0.999, O[g] * exp(-sigma))
                if alpha < 1.0 / 255.0:
                    continue
                next_T = T * (1.0 - alpha)
                if next_T <= 1e-4:
                    break
                vis = alpha * T
                cg = Col[g]
                for c in range(ch):
                    acc[c] += vis * cg[c]
                T = next_T
            out[py - y0, px - x0] = torch.tensor([acc[c] + T * Bg[c] for c in range(ch)], dtype=torch.float64)
    return out, pairs


def rasterize_grads(xys, conics, opac, colors, ids_sorted, tile_ranges, H, W, bg, v_out, tiles=None):
    """Tile-by-tile autograd of A9 with bounded memory.

    Returns (out, v_xys, v_conics, v_opac, v_colors) for detached leaf copies of the inputs."""
    dt = xys.dtype
    leaves = [t.detach().clone().requires_grad_(True) for t in (xys, conics, opac.reshape(-1), colors)]
    lx, lc, lo, lcol = leaves
    ch = colors.shape[1]
    tiles_x = (W + TILE - 1) // TILE
    tiles_y = (H + TILE - 1) // TILE
    grads = [torch.zeros_like(t) for t in leaves]
    out = torch.zeros((H, W, ch), dtype=dt)
    ids64 = ids_sorted.to(torch.int64)
    bg = bg.to(dt)
    v_out = v_out.to(dt)
    tile_iter = range(tiles_x * tiles_y) if tiles is None else tiles
    for t in tile_iter:
        ty, tx = divmod(int(t), tiles_x)
        y0, x0 = ty * TILE, tx * TILE
        y1, x1 = min(y0 + TILE, H), min(x0 + TILE, W)
        yy, xx = torch.meshgrid(torch.arange(y0, y1), torch.arange(x0, x1), indexing="ij")
        px, py = xx.reshape(-1).to(dt), yy.reshape(-1).to(dt)
        s, e = int(tile_ranges[t, 0]), int(tile_ranges[t, 1])
        o, _, _ = _blend_tile(px, py, ids64[s:e], lx, lc, lo, lcol, bg)
        out[y0:y1, x0:x1] = o.detach().reshape(y1 - y0, x1 - x0, ch)
        if e > s:
            g = torch.autograd.grad(o, leaves, v_out[y0:y1, x0:x1].reshape(-1, ch), allow_unused=True)
            for acc, gi in zip(grads, g):
                if gi is not None:
                    acc += gi
    return out, grads[0], grads[1], grads[2], grads[3]


# --------------------------------------------------------------------------
# camera helpers following gaussian_splatting.py:87-105 and :658-676
# --------------------------------------------------------------------------
def projection_matrix(znear, zfar, fovx, fovy):
    t = znear * math.tan(0.5 * fovy)
    b = -t
    r = znear * math.tan(0.5 * fovx)
    l = -r
    n, f = znear, zfar
    return torch.tensor([
        [2 * n / (r - l), 0.0, (r + l) / (r - l), 0.0],
        [0.0, 2 * n / (t - b), (t + b) / (t - b), 0.0],
        [0.0, 0.0, (f + n) / (f - n), -1.0 * f * n / (f - n)],
        [0.0, 0.0, 1.0, 0.0]], dtype=torch.float32)
