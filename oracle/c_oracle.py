"""numpy/ctypes front-end of the C oracle (oracle/gg_oracle.c).

TEST INFRASTRUCTURE ONLY -- see the header of gg_oracle.c.  PARITY UNPINNED: no
gsplat 0.1.0 binary or golden vector exists for this path.
Only tests/, bench.py's CPU legs and __graft_entry__.smoke() import this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libgg_oracle.so")
_lib = None


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "gg_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.run(["make", "-C", _HERE, "-B", "libgg_oracle.so"], check=True, capture_output=True)
    return _SO


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_SO)
        _lib.gg_oracle_blend_fwd.restype = C.c_int64
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def project_fwd(means, scales, glob_scale, quats, viewmat, fullmat, fx, fy, cx, cy, H, W, tile_bounds,
                clip_thresh=0.01):
    means, scales, quats = _f32(means), _f32(scales), _f32(quats)
    vm, fm = _f32(viewmat).reshape(-1), _f32(fullmat).reshape(-1)
    n = means.shape[0]
    cov3d = np.empty((n, 6), np.float32)
    xys = np.empty((n, 2), np.float32)
    depths = np.empty((n,), np.float32)
    radii = np.empty((n,), np.int32)
    conics = np.empty((n, 3), np.float32)
    nth = np.empty((n,), np.int32)
    lib().gg_oracle_project_fwd(
        C.c_int(n), _p(means), _p(scales), C.c_float(glob_scale), _p(quats), _p(vm), _p(fm),
        C.c_float(fx), C.c_float(fy), C.c_float(cx), C.c_float(cy), C.c_int(H), C.c_int(W),
        C.c_int(tile_bounds[0]), C.c_int(tile_bounds[1]), C.c_float(clip_thresh),
        _p(cov3d), _p(xys), _p(depths), _p(radii), _p(conics), _p(nth))
    return xys, depths, radii, conics, nth, cov3d


def num_sh_bases(degree):
    return int(lib().gg_oracle_num_sh_bases(C.c_int(degree)))


def sh_fwd(degrees_to_use, dirs, coeffs):
    dirs, coeffs = _f32(dirs), _f32(coeffs)
    n, nb = coeffs.shape[0], coeffs.shape[1]
    degree = {1: 0, 4: 1, 9: 2, 16: 3, 25: 4}[nb]
    out = np.empty((n, 3), np.float32)
    lib().gg_oracle_sh_fwd(C.c_int(n), C.c_int(degree), C.c_int(degrees_to_use), _p(dirs), _p(coeffs), _p(out))
    return out


def sh_bwd(degree, degrees_to_use, dirs, v_colors):
    dirs, v_colors = _f32(dirs), _f32(v_colors)
    n = dirs.shape[0]
    nb = num_sh_bases(degree)
    out = np.empty((n, nb, 3), np.float32)
    lib().gg_oracle_sh_bwd(C.c_int(n), C.c_int(degree), C.c_int(degrees_to_use), _p(dirs), _p(v_colors), _p(out))
    return out


def tile_counts(xys, radii, tile_bounds):
    """num_tiles_hit of made-up projection outputs (A6 only)."""
    xys, radii = _f32(xys), _i32(radii)
    out = np.empty((xys.shape[0],), np.int32)
    lib().gg_oracle_tile_counts(C.c_int(xys.shape[0]), _p(xys), _p(radii), C.c_int(int(tile_bounds[0])),
                                C.c_int(int(tile_bounds[1])), _p(out))
    return out


def bin_and_sort(xys, depths, radii, num_tiles_hit, tile_bounds):
    """Returns (cum, keys_unsorted, ids_unsorted, keys_sorted, ids_sorted, tile_ranges[T,2])."""
    xys, depths, radii, nth = _f32(xys), _f32(depths), _i32(radii), _i32(num_tiles_hit)
    n = xys.shape[0]
    tx, ty = int(tile_bounds[0]), int(tile_bounds[1])
    cum = np.empty((n,), np.int32)
    lib().gg_oracle_cumsum(C.c_int(n), _p(nth), _p(cum))
    m = int(cum[-1]) if n > 0 else 0
    keys = np.empty((max(m, 1),), np.int64)
    ids = np.empty((max(m, 1),), np.int32)
    lib().gg_oracle_map_to_intersects(C.c_int(n), _p(xys), _p(depths), _p(radii), _p(cum), C.c_int(tx), C.c_int(ty),
                                      _p(keys), _p(ids))
    keys_s = np.empty_like(keys)
    ids_s = np.empty_like(ids)
    lib().gg_oracle_sort(C.c_int64(m), _p(keys), _p(ids), _p(keys_s), _p(ids_s))
    ranges = np.empty((tx * ty, 2), np.int32)
    lib().gg_oracle_tile_ranges(C.c_int64(m), _p(keys_s), C.c_int(tx * ty), _p(ranges))
    return cum, keys[:m], ids[:m], keys_s[:m], ids_s[:m], ranges


def blend_fwd(H, W, tile_bounds, ids_sorted, tile_ranges, xys, conics, opac, colors, bg, eps=1e-4):
    """Returns (out[H,W,C], final_T[H,W], final_idx[H,W], fragile[H,W] bool, pairs)."""
    ids_sorted, tile_ranges = _i32(ids_sorted), _i32(tile_ranges)
    xys, conics, opac, colors, bg = _f32(xys), _f32(conics), _f32(opac).reshape(-1), _f32(colors), _f32(bg)
    ch = colors.shape[1]
    out = np.empty((H, W, ch), np.float32)
    fT = np.empty((H, W), np.float32)
    fi = np.empty((H, W), np.int32)
    frag = np.zeros((H, W), np.uint8)
    pairs = lib().gg_oracle_blend_fwd(
        C.c_int(ch), C.c_int(H), C.c_int(W), C.c_int(tile_bounds[0]), C.c_int(tile_bounds[1]),
        _p(ids_sorted), _p(tile_ranges), _p(xys), _p(conics), _p(opac), _p(colors), _p(bg),
        _p(out), _p(fT), _p(fi), _p(frag), C.c_float(eps))
    return out, fT, fi, frag.astype(bool), int(pairs)


def blend_bwd(H, W, tile_bounds, ids_sorted, tile_ranges, xys, conics, opac, colors, bg, v_out):
    """fp64 analytic gradients (v_xy[N,2], v_conic[N,3], v_colors[N,C], v_opac[N])."""
    ids_sorted, tile_ranges = _i32(ids_sorted), _i32(tile_ranges)
    xys, conics, opac, colors, bg = _f32(xys), _f32(conics), _f32(opac).reshape(-1), _f32(colors), _f32(bg)
    v_out = _f32(v_out)
    n, ch = colors.shape
    v_xy = np.empty((n, 2), np.float64)
    v_conic = np.empty((n, 3), np.float64)
    v_colors = np.empty((n, ch), np.float64)
    v_opac = np.empty((n,), np.float64)
    lib().gg_oracle_blend_bwd(
        C.c_int(n), C.c_int(ch), C.c_int(H), C.c_int(W), C.c_int(tile_bounds[0]),
        _p(ids_sorted), _p(tile_ranges), _p(xys), _p(conics), _p(opac), _p(colors), _p(bg), _p(v_out),
        _p(v_xy), _p(v_conic), _p(v_colors), _p(v_opac))
    return v_xy, v_conic, v_colors, v_opac


def blend_bwd_ex(H, W, tile_bounds, ids_sorted, tile_ranges, xys, conics, opac, colors, bg, v_out, eps=2e-5):
    """blend_bwd plus the two element-wise error scales of gg_oracle_blend_bwd_ex:
    returns dict(v_xy, v_conic, v_colors, v_opac, abs_geo[N,6], abs_colors[N,C], taint_geo[N,6], taint_colors[N,C])
    with the geo columns ordered (x, y, A, B, C, opacity)."""
    ids_sorted, tile_ranges = _i32(ids_sorted), _i32(tile_ranges)
    xys, conics, opac, colors, bg = _f32(xys), _f32(conics), _f32(opac).reshape(-1), _f32(colors), _f32(bg)
    v_out = _f32(v_out)
    n, ch = colors.shape
    r = dict(v_xy=np.empty((n, 2), np.float64), v_conic=np.empty((n, 3), np.float64),
             v_colors=np.empty((n, ch), np.float64), v_opac=np.empty((n,), np.float64),
             abs_geo=np.empty((n, 6), np.float64), abs_colors=np.empty((n, ch), np.float64),
             taint_geo=np.empty((n, 6), np.float64), taint_colors=np.empty((n, ch), np.float64))
    lib().gg_oracle_blend_bwd_ex(
        C.c_int(n), C.c_int(ch), C.c_int(H), C.c_int(W), C.c_int(tile_bounds[0]),
        _p(ids_sorted), _p(tile_ranges), _p(xys), _p(conics), _p(opac), _p(colors), _p(bg), _p(v_out),
        _p(r["v_xy"]), _p(r["v_conic"]), _p(r["v_colors"]), _p(r["v_opac"]), C.c_float(eps),
        _p(r["abs_geo"]), _p(r["abs_colors"]), _p(r["taint_geo"]), _p(r["taint_colors"]))
    return r
