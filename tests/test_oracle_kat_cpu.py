"""Known-answer tests that pin the oracle itself (CPU).

PARITY UNPINNED: there is no gsplat 0.1.0 binary or golden vector.  What CAN be pinned without
one is that the two restatements compute the published formulas -- here they are checked
against closed forms worked out by hand in this file (float64, no shared code with oracle/), on cases small
enough to reason about: one isotropic Gaussian on the optical axis, two Gaussians occluding each other, every
branch threshold of the blend (alpha clamp 0.999, alpha < 1/255, T(1-alpha) <= 1e-4, sigma < 0), hand-placed
tile boxes / keys / ranges, and central finite differences of the fp64 backward.
"""
import math

import numpy as np
import pytest
import torch

from gaussiangrasper_b200 import scenes
from oracle import c_oracle, torch_oracle

W, H = 64, 48
FX, FY, CX, CY = 50.0, 40.0, 31.5, 24.5   # cx - 0.5 = 31, cy - 0.5 = 24: the axis hits the centre of pixel (24, 31)
TB = ((W + 15) // 16, (H + 15) // 16, 1)


def axis_camera():
    """Identity view matrix (camera at the origin, +z forward) with the reference's projection matrix."""
    viewmat = torch.eye(4)
    proj = scenes.projection_matrix(0.001, 1000.0, 2 * math.atan(W / (2 * FX)), 2 * math.atan(H / (2 * FY)))
    return viewmat, proj @ viewmat


def project_both(means, scales, quats):
    viewmat, fullmat = axis_camera()
    c = c_oracle.project_fwd(means, scales, 1.0, quats, viewmat[:3].numpy(), fullmat.numpy(), FX, FY, CX, CY, H, W, TB)
    t = torch_oracle.project_gaussians(torch.from_numpy(means).double(), torch.from_numpy(scales).double(), 1.0,
                                       torch.from_numpy(quats).double(), viewmat, fullmat, FX, FY, CX, CY, H, W, TB)
    return c, tuple(x.numpy() for x in t)


def blend_both(xys, conics, opac, colors, bg, depths, radii, nth):
    _, _, _, _, ids_s, ranges = c_oracle.bin_and_sort(xys, depths, radii, nth, TB)
    out_c, T_c, idx_c, _, _ = c_oracle.blend_fwd(H, W, TB, ids_s, ranges, xys, conics, opac, colors, bg)
    f64 = lambda a: torch.from_numpy(np.asarray(a)).double()
    out_t, T_t, idx_t = torch_oracle.rasterize(f64(xys), f64(conics), f64(opac), f64(colors), torch.from_numpy(ids_s),
                                               torch.from_numpy(ranges), H, W, f64(bg))
    return (out_c, T_c, idx_c), (out_t.numpy(), T_t.numpy(), idx_t.numpy()), ids_s, ranges


def test_single_isotropic_gaussian_closed_form():
    z0, s, o = 4.0, 0.5, 0.8
    means = np.array([[0.0, 0.0, z0]], np.float32)
    scales = np.full((1, 3), s, np.float32)
    quats = np.array([[1.0, 0.0, 0.0, 0.0]], np.float32)
    # closed form (Appendix A1-A6): J = diag(fx/z, fy/z) on the axis, Sigma = s^2 I
    a = (FX * s / z0) ** 2 + 0.3
    c = (FY * s / z0) ** 2 + 0.3
    conic = np.array([1 / a, 0.0, 1 / c])
    radius = math.ceil(3 * math.sqrt(max(a, c)))
    xy = np.array([CX - 0.5, CY - 0.5])
    x0, x1 = int(xy[0] / 16 - radius / 16), int(xy[0] / 16 + radius / 16 + 1)
    y0, y1 = int(xy[1] / 16 - radius / 16), int(xy[1] / 16 + radius / 16 + 1)
    area = (min(x1, TB[0]) - max(x0, 0)) * (min(y1, TB[1]) - max(y0, 0))
    for got in project_both(means, scales, quats):
        xys, depths, radii, conics, nth, cov3d = got
        np.testing.assert_allclose(xys[0], xy, atol=2e-4)
        np.testing.assert_allclose(depths[0], z0, rtol=1e-6)
        assert radii[0] == radius and nth[0] == area
        np.testing.assert_allclose(conics[0], conic, rtol=1e-5, atol=1e-9)
        np.testing.assert_allclose(cov3d[0], [s * s, 0, 0, s * s, 0, s * s], rtol=1e-6, atol=1e-9)
    # blend (A9): one Gaussian -> out = alpha c + (1 - alpha) bg with alpha = o exp(-sigma), cut below 1/255
    colors = np.array([[0.2, 0.7, 0.4]], np.float32)
    bg = np.array([0.05, 0.1, 0.9], np.float32)
    xys, depths, radii, conics, nth, _ = project_both(means, scales, quats)[0]
    jj, ii = np.meshgrid(np.arange(W), np.arange(H))
    sigma = 0.5 * (conic[0] * (xy[0] - jj) ** 2 + conic[2] * (xy[1] - ii) ** 2)
    alpha = np.minimum(0.999, o * np.exp(-sigma))
    alpha[alpha < 1 / 255] = 0.0
    # pixels outside the Gaussian's tile box never see it
    tile_mask = np.zeros((H, W), bool)
    tile_mask[max(y0, 0) * 16:min(y1, TB[1]) * 16, max(x0, 0) * 16:min(x1, TB[0]) * 16] = True
    alpha[~tile_mask] = 0.0
    ref = alpha[..., None] * colors[0].astype(np.float64) + (1 - alpha[..., None]) * bg.astype(np.float64)
    (out_c, T_c, idx_c), (out_t, T_t, idx_t), _, _ = blend_both(xys, conics, np.array([o], np.float32), colors, bg,
                                                               depths, radii, nth)
    near = np.abs(o * np.exp(-sigma) - 1 / 255) < 1e-6   # the cut itself: fp32 vs fp64 may disagree there
    assert near.sum() < 8
    np.testing.assert_allclose(out_c[~near], ref[~near], atol=3e-6)
    np.testing.assert_allclose(out_t[~near], ref[~near], atol=1e-6)
    np.testing.assert_allclose(T_c[~near], (1 - alpha)[~near], atol=3e-6)
    assert alpha.max() > 0.79 and (alpha > 0).sum() > 300


def test_two_gaussian_occlusion_closed_form():
    """Front-to-back order is by depth, not by id: the farther Gaussian has the LOWER id here."""
    xy = np.array([[31.0, 24.0], [33.0, 25.0]], np.float32)
    conics = np.array([[0.02, 0.0, 0.03], [0.04, 0.01, 0.02]], np.float32)
    depths = np.array([5.0, 2.0], np.float32)   # id 0 is behind id 1
    radii = np.array([30, 30], np.int32)
    nth = c_oracle.tile_counts(xy, radii, TB)
    opac = np.array([0.9, 0.6], np.float32)
    colors = np.array([[1.0, 0.0, 0.25], [0.0, 1.0, 0.5]], np.float32)
    bg = np.array([0.3, 0.3, 0.3], np.float32)
    jj, ii = np.meshgrid(np.arange(W), np.arange(H))

    def alpha_of(g):
        dx, dy = xy[g, 0] - jj, xy[g, 1] - ii
        A, B, C = conics[g].astype(np.float64)
        sig = 0.5 * (A * dx * dx + C * dy * dy) + B * dx * dy
        al = np.minimum(0.999, float(opac[g]) * np.exp(-sig))
        al[(al < 1 / 255) | (sig < 0)] = 0.0
        return al

    a_front, a_back = alpha_of(1), alpha_of(0)
    ref = (a_front[..., None] * colors[1] + ((1 - a_front) * a_back)[..., None] * colors[0]
           + ((1 - a_front) * (1 - a_back))[..., None] * bg)
    (out_c, T_c, _), (out_t, T_t, _), ids_s, ranges = blend_both(xy, conics, opac, colors, bg, depths, radii, nth)
    assert nth[0] == nth[1] == TB[0] * TB[1]          # both cover every tile
    assert ids_s[:2].tolist() == [1, 0]               # depth order inside a tile
    edge = (np.abs(a_front - 1 / 255) < 2e-6) | (np.abs(a_back - 1 / 255) < 2e-6)
    np.testing.assert_allclose(out_c[~edge], ref[~edge], atol=5e-6)
    np.testing.assert_allclose(out_t[~edge], ref[~edge], atol=1e-6)
    np.testing.assert_allclose(T_t[~edge], ((1 - a_front) * (1 - a_back))[~edge], atol=1e-6)


def test_blend_thresholds_closed_form():
    """alpha is clamped at 0.999; alpha < 1/255 is skipped; the entry that would take T(1-alpha) to <= 1e-4 is
    NOT added and ends the pixel; sigma < 0 (indefinite conic) is skipped (Appendix A9)."""
    centre = np.array([[31.0, 24.0]], np.float32)
    wide = np.array([[1e-6, 0.0, 1e-6]], np.float32)          # sigma ~ 0 everywhere: alpha = opacity
    col = lambda *c: np.array([c], np.float32)
    bg = np.array([0.5, 0.25, 0.125], np.float32)

    def run(opacs, colors, conics=None, depths=None):
        n = len(opacs)
        xy = np.repeat(centre, n, 0)
        con = np.repeat(wide, n, 0) if conics is None else conics
        d = np.arange(1, n + 1, dtype=np.float32) if depths is None else depths
        radii = np.full(n, 100, np.int32)
        nth = c_oracle.tile_counts(xy, radii, TB)
        return blend_both(xy, con, np.asarray(opacs, np.float32), np.concatenate(colors), bg, d, radii, nth)

    # clamp: opacity 1 -> alpha 0.999
    (oc, Tc, ic), (ot, Tt, it), _, ranges = run([1.0], [col(1, 0, 0)])
    for out, T in ((oc, Tc), (ot, Tt)):
        np.testing.assert_allclose(out[24, 31], 0.999 * np.array([1, 0, 0]) + 0.001 * bg, atol=2e-6)
        np.testing.assert_allclose(T[24, 31], 0.001, rtol=2e-3)
    # below 1/255: contributes nothing at all
    (oc, Tc, ic), (ot, Tt, it), _, ranges = run([1 / 300], [col(1, 0, 0)])
    for out, T, idx in ((oc, Tc, ic), (ot, Tt, it)):
        np.testing.assert_array_equal(out[24, 31], bg)
        assert T[24, 31] == 1.0
        assert idx[24, 31] == ranges[(24 // 16) * TB[0] + 31 // 16, 0]   # no contributor: the tile's start
    # stop rule: 0.999 then 0.999 -> the second would leave T = 1e-6 <= 1e-4: not added, pixel ends with T = 0.001
    (oc, Tc, ic), (ot, Tt, it), _, ranges = run([1.0, 1.0, 0.5], [col(1, 0, 0), col(0, 1, 0), col(0, 0, 1)])
    for out, T, idx in ((oc, Tc, ic), (ot, Tt, it)):
        np.testing.assert_allclose(out[24, 31], 0.999 * np.array([1, 0, 0]) + 0.001 * bg, atol=2e-6)
        np.testing.assert_allclose(T[24, 31], 0.001, rtol=2e-3)
        assert idx[24, 31] == ranges[(24 // 16) * TB[0] + 31 // 16, 0] + 1
    # just above the stop: 0.9 then 0.9989 leaves 0.1 * 0.0011 = 1.1e-4 > 1e-4 -> both are added
    (oc, Tc, ic), (ot, Tt, it), _, _ = run([0.9, 0.9989], [col(1, 0, 0), col(0, 1, 0)])
    ref = 0.9 * np.array([1, 0, 0]) + 0.1 * 0.9989 * np.array([0, 1, 0]) + 0.1 * 0.0011 * bg
    np.testing.assert_allclose(oc[24, 31], ref, atol=2e-6)
    np.testing.assert_allclose(ot[24, 31], ref, atol=1e-6)
    # sigma < 0 (indefinite conic, away from the centre): skipped
    neg = np.array([[-0.01, 0.0, -0.01]], np.float32)
    (oc, Tc, _), (ot, Tt, _), _, _ = run([0.9], [col(1, 0, 0)], conics=neg)
    for out in (oc, ot):
        np.testing.assert_array_equal(out[10, 10], bg)


def test_keys_ranges_by_hand():
    """4x3 tiles of a 64x48 image; three hand-placed discs."""
    xys = np.array([[8.0, 8.0], [40.0, 20.0], [63.0, 47.0]], np.float32)
    radii = np.array([4, 20, 1], np.int32)
    depths = np.array([3.0, 1.5, 2.0], np.float32)
    # A6 by hand: tile box = [trunc((c - r)/16), trunc((c + r)/16 + 1)) clamped to the grid
    #   g0: x [0.25, 0.75+1) -> [0,1), y likewise -> 1 tile: (0,0)
    #   g1: x [1.25, 3.75+1) -> [1,4), y [0, 2.5+1) -> [0,3)  -> 9 tiles
    #   g2: x [3.875, 4+1) -> [3,4) (clamped), y [2.875, 3+1) -> [2,3) -> 1 tile: (2,3)
    nth = c_oracle.tile_counts(xys, radii, TB)
    assert nth.tolist() == [1, 9, 1]
    cum, keys, ids, keys_s, ids_s, ranges = c_oracle.bin_and_sort(xys, depths, radii, nth, TB)
    assert cum.tolist() == [1, 10, 11]
    bits = lambda d: int(np.float32(d).view(np.int32))
    g1_tiles = [ty * 4 + tx for ty in range(3) for tx in range(1, 4)]
    want_keys = [(0 << 32) | bits(3.0)] + [(t << 32) | bits(1.5) for t in g1_tiles] + [(11 << 32) | bits(2.0)]
    assert keys.tolist() == want_keys and ids.tolist() == [0] + [1] * 9 + [2]
    # tile 11 holds g1 (depth 1.5) before g2 (depth 2.0)
    assert ids_s.tolist() == [0, 1, 1, 1, 1, 1, 1, 1, 1, 1, 2]
    want_ranges = np.zeros((12, 2), np.int32)
    want_ranges[0] = (0, 1)
    for k, t in enumerate(g1_tiles[:-1]):
        want_ranges[t] = (1 + k, 2 + k)
    want_ranges[11] = (9, 11)
    np.testing.assert_array_equal(ranges, want_ranges)
    # the torch restatement agrees
    _, keys_t, ids_t, ranges_t = torch_oracle.bin_and_sort(torch.from_numpy(xys), torch.from_numpy(depths),
                                                           torch.from_numpy(radii), torch.from_numpy(nth), TB)
    assert keys_t.tolist() == sorted(want_keys) and ids_t.tolist() == ids_s.tolist()
    np.testing.assert_array_equal(ranges_t.numpy(), want_ranges)


def test_sh_closed_form_low_bands():
    """Band 0 and 1 by hand (Appendix A10): colour = C0 c0 + C1 (-y c1 + z c2 - x c3)."""
    d = np.array([[0.0, 0.0, 2.0], [3.0, 0.0, 0.0], [0.0, -1.0, 0.0], [1.0, 2.0, 2.0]], np.float32)
    rng = np.random.default_rng(0)
    coeffs = rng.normal(size=(4, 25, 3)).astype(np.float32)
    u = d / np.linalg.norm(d, axis=1, keepdims=True)
    C0, C1 = 0.28209479177387814, 0.4886025119029199
    ref = C0 * coeffs[:, 0] + C1 * (-u[:, 1:2] * coeffs[:, 1] + u[:, 2:3] * coeffs[:, 2] - u[:, 0:1] * coeffs[:, 3])
    np.testing.assert_allclose(c_oracle.sh_fwd(1, d, coeffs), ref, rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(torch_oracle.spherical_harmonics(1, torch.from_numpy(d), torch.from_numpy(coeffs)).numpy(),
                               ref, rtol=1e-5, atol=1e-6)
    # orthonormality of all 25 basis functions over the sphere (Monte-Carlo): pins constants and signs of bands 2-4
    g = torch.Generator().manual_seed(1)
    v = torch.randn((400_000, 3), generator=g, dtype=torch.float64)
    Y = torch_oracle.sh_basis(4, v)
    gram = (Y.T @ Y) / v.shape[0] * (4 * math.pi)
    np.testing.assert_allclose(gram.numpy(), np.eye(25), atol=0.03)


@pytest.mark.parametrize("seed", [0, 1])
def test_blend_backward_vs_finite_differences(seed):
    """gg_oracle_blend_bwd (the fp64 analytic gradient the GPU backward is held to) against central differences
    of the fp64 forward, on a scene whose branch decisions do not change under the perturbation."""
    rng = np.random.default_rng(seed)
    n, C = 14, 5
    Hs, Ws = 32, 32
    tb = (2, 2, 1)
    xys = rng.uniform(4, 28, (n, 2)).astype(np.float32)
    base = rng.uniform(0.01, 0.06, (n, 2))
    conics = np.stack([base[:, 0], rng.uniform(-0.005, 0.005, n), base[:, 1]], 1).astype(np.float32)
    opac = rng.uniform(0.2, 0.9, n).astype(np.float32)
    colors = rng.uniform(-1, 1, (n, C)).astype(np.float32)
    depths = rng.uniform(1, 5, n).astype(np.float32)
    radii = np.full(n, 40, np.int32)
    bg = rng.uniform(0, 1, C).astype(np.float32)
    v_out = rng.normal(size=(Hs, Ws, C)).astype(np.float32)
    nth = c_oracle.tile_counts(xys, radii, tb)
    _, _, _, _, ids_s, ranges = c_oracle.bin_and_sort(xys, depths, radii, nth, tb)
    v_xy, v_conic, v_colors, v_opac = c_oracle.blend_bwd(Hs, Ws, tb, ids_s, ranges, xys, conics, opac, colors, bg, v_out)
    ids_t, ranges_t = torch.from_numpy(ids_s), torch.from_numpy(ranges)
    f64 = lambda a: torch.from_numpy(np.asarray(a)).double()
    leaves = dict(xys=f64(xys), conics=f64(conics), opac=f64(opac), colors=f64(colors))
    vo = f64(v_out)

    def loss(lv):
        out, T, idx = torch_oracle.rasterize(lv["xys"], lv["conics"], lv["opac"], lv["colors"], ids_t, ranges_t, Hs, Ws,
                                             f64(bg))
        return float((out * vo).sum()), idx

    grads = dict(xys=v_xy, conics=v_conic, opac=v_opac, colors=v_colors)
    eps = 1e-6
    checked = 0
    for name, g in grads.items():
        flat = leaves[name].reshape(-1)
        for k in rng.choice(flat.numel(), size=min(12, flat.numel()), replace=False):
            old = float(flat[k])
            flat[k] = old + eps
            lp, ip = loss(leaves)
            flat[k] = old - eps
            lm, im = loss(leaves)
            flat[k] = old
            if not torch.equal(ip, im):
                continue   # a pixel's last contributor changed: a branch flipped inside the stencil
            fd = (lp - lm) / (2 * eps)
            ref = float(np.asarray(g).reshape(-1)[k])
            assert abs(fd - ref) <= 2e-5 * max(1.0, abs(ref)), (name, int(k), fd, ref)
            checked += 1
    assert checked >= 30


def test_offaxis_rotated_gaussians_ewa_from_a_numerical_jacobian():
    """Projection (Appendix A1-A5) of rotated, anisotropic, off-axis Gaussians against a derivation that shares no
    formula with the oracle: the rotation matrices are the reference's own (nerfstudio's quaternion_matrix, stored in
    tests/golden/ref_init_small.npz), and the Jacobian of the pinhole projection is taken NUMERICALLY (central
    differences of (fx x/z + cx, fy y/z + cy) in float64) instead of from the analytic EWA expression:
    cov2d = J (W Sigma W^T) J^T + 0.3 I, conic = cov2d^-1, radius = ceil(3 sqrt(lambda_max)), xys = projection - 0.5."""
    import os
    fix = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_init_small.npz"))
    q, R = fix["quat_wxyz"][:24].astype(np.float32), fix["quat_rotmat"][:24]
    n = len(q)
    g = np.random.default_rng(3)
    means = np.stack([g.uniform(-1.2, 1.2, n), g.uniform(-0.9, 0.9, n), g.uniform(2.5, 6.0, n)], axis=1).astype(np.float32)
    scales = g.uniform(0.05, 0.4, (n, 3)).astype(np.float32)
    qn = (q / np.linalg.norm(q, axis=1, keepdims=True)).astype(np.float32)

    def pinhole(p):
        return np.array([FX * p[0] / p[2] + CX, FY * p[1] / p[2] + CY])
    for got in project_both(means, scales, qn):
        xys, depths, radii, conics, nth, cov3d = got
        for i in range(n):
            p = means[i].astype(np.float64)                      # identity view matrix: camera space = world space
            h = 1e-6
            J = np.stack([(pinhole(p + h * e) - pinhole(p - h * e)) / (2 * h) for e in np.eye(3)], axis=1)   # [2,3]
            Sigma = R[i] @ np.diag(scales[i].astype(np.float64) ** 2) @ R[i].T
            cov2d = J @ Sigma @ J.T + 0.3 * np.eye(2)
            det = cov2d[0, 0] * cov2d[1, 1] - cov2d[0, 1] ** 2
            conic = np.array([cov2d[1, 1], -cov2d[0, 1], cov2d[0, 0]]) / det
            b = 0.5 * (cov2d[0, 0] + cov2d[1, 1])
            lam = b + math.sqrt(max(0.1, b * b - det))
            np.testing.assert_allclose(xys[i], pinhole(p) - 0.5, atol=5e-4)
            np.testing.assert_allclose(depths[i], p[2], rtol=1e-6)
            np.testing.assert_allclose(conics[i], conic, rtol=2e-4, atol=1e-7)
            assert abs(int(radii[i]) - math.ceil(3 * math.sqrt(lam))) <= (1 if abs(3 * math.sqrt(lam) % 1) < 1e-3 else 0)
            assert radii[i] > 0 and nth[i] > 0
