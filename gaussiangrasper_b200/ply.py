"""`.ply` export of a trained scene with the properties the reference's exporter writes
(nerfstudio/scripts/exporter.py:482-530, `ns-export gaussian-splat`; SURVEY 8-f4 formats row).

The reference hands a name -> array map to open3d (`o3d.t.geometry.PointCloud` / `o3d.t.io.write_point_cloud`):
    positions -> x y z (float32), normals -> nx ny nz (zeros), colors -> red green blue (uint8, = model.colors * 255),
    f_dc_0..2 = model.colors = SH2RGB(colors_all[:, 0, :]), f_rest_i, opacity (logit), scale_0..2 (log), rot_0..3 (raw wxyz).
open3d is not installable here and the order in which it emits the custom attributes is its own (an unordered
map); PLY readers address properties BY NAME, so this writer fixes the order listed above, binary little endian.

One quirk of the reference is reproduced by default: it reshapes shs_rest to (N, -1, 1) and then loops over the
LAST axis (:513-516), so exactly one property `f_rest_0` -- the first higher-band coefficient -- is written.
`full_sh=True` writes all 3*(K-1) coefficients channel-major (the layout 3DGS viewers expect) instead.
"""
from __future__ import annotations

from typing import Dict, List, Mapping, Tuple

import numpy as np
import torch

SH_C0 = 0.28209479177387814


def sh2rgb(sh: np.ndarray) -> np.ndarray:
    """gaussian_splatting.py:80-85."""
    return sh * SH_C0 + 0.5


def ply_properties(params: Mapping[str, torch.Tensor], full_sh: bool = False) -> List[Tuple[str, np.ndarray]]:
    """Ordered (name, column) list; every column is [N] float32 except red/green/blue (uint8)."""
    g = lambda k: params[k].detach().to(torch.float32).cpu().numpy()
    means, sh = g("means"), g("sh_coeffs")
    n = means.shape[0]
    cols: List[Tuple[str, np.ndarray]] = [("x", means[:, 0]), ("y", means[:, 1]), ("z", means[:, 2])]
    cols += [(k, np.zeros(n, np.float32)) for k in ("nx", "ny", "nz")]
    colors = sh2rgb(sh[:, 0, :]).astype(np.float32)
    rgb8 = (colors * 255).astype(np.uint8)             # the reference's cast: no clamp, C wrap-around semantics
    cols += [("red", rgb8[:, 0]), ("green", rgb8[:, 1]), ("blue", rgb8[:, 2])]
    cols += [(f"f_dc_{i}", colors[:, i]) for i in range(3)]
    if sh.shape[1] > 1:
        rest = sh[:, 1:, :]
        if full_sh:
            flat = np.ascontiguousarray(rest.transpose(0, 2, 1)).reshape(n, -1)   # channel-major, as 3DGS stores it
            cols += [(f"f_rest_{i}", flat[:, i]) for i in range(flat.shape[1])]
        else:
            cols.append(("f_rest_0", rest.reshape(n, -1)[:, 0]))                  # exporter.py:513-516 as written
    cols.append(("opacity", g("opacity_logit").reshape(n)))
    ls, q = g("log_scales"), g("quats")
    cols += [(f"scale_{i}", ls[:, i]) for i in range(3)]
    cols += [(f"rot_{i}", q[:, i]) for i in range(4)]
    return [(k, np.ascontiguousarray(v)) for k, v in cols]


def write_ply(path: str, params: Mapping[str, torch.Tensor], full_sh: bool = False) -> int:
    """Write `point_cloud.ply`; returns the number of vertices."""
    cols = ply_properties(params, full_sh)
    n = int(cols[0][1].shape[0])
    dt = np.dtype([(k, "u1" if v.dtype == np.uint8 else "<f4") for k, v in cols])
    rec = np.empty(n, dtype=dt)
    for k, v in cols:
        rec[k] = v
    header = ["ply", "format binary_little_endian 1.0", f"element vertex {n}"]
    header += [f"property {'uchar' if v.dtype == np.uint8 else 'float'} {k}" for k, v in cols]
    header.append("end_header")
    with open(path, "wb") as f:
        f.write(("\n".join(header) + "\n").encode("ascii"))
        f.write(rec.tobytes())
    return n


def read_ply(path: str) -> Dict[str, np.ndarray]:
    """Minimal reader of what write_ply produces (binary little endian, one vertex element)."""
    with open(path, "rb") as f:
        data = f.read()
    end = data.index(b"end_header\n") + len(b"end_header\n")
    lines = data[:end].decode("ascii").splitlines()
    if lines[0] != "ply" or "binary_little_endian" not in lines[1]:
        raise ValueError("not a binary little-endian PLY file")
    n = int(next(l for l in lines if l.startswith("element vertex")).split()[-1])
    types = {"float": "<f4", "uchar": "u1", "double": "<f8", "int": "<i4"}
    dt = np.dtype([(l.split()[2], types[l.split()[1]]) for l in lines if l.startswith("property")])
    rec = np.frombuffer(data, dtype=dt, count=n, offset=end)
    return {k: np.array(rec[k]) for k in dt.names}
