#!/usr/bin/env python
"""Seeded inputs of the refinement fixture tests/golden/refine_small.npz, and the oracle's version of it.

The committed refine_small.npz is written by tests/golden/make_reference_golden.py: the REFERENCE's own
GaussianSplattingModel.refinement_after (nerfstudio/models/gaussian_splatting.py:402-464) run on `inputs(SEED)`.  This
script runs oracle/refine_oracle.py on the same inputs and reports whether it reproduces the committed file (it does,
array by array, bit for bit -- tests/test_golden.py::test_refine_oracle_reproduces_golden keeps checking that).

    python tests/golden/make_refine_golden.py
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import refine_oracle  # noqa: E402

N, D, K, SEED = 601, 3, 4, 77
RULES = dict(max_dim=640.0, densify_grad_thresh=0.0002, densify_size_thresh=0.01, split_screen_size=0.05,
             cull_alpha_thresh=0.1, cull_scale_thresh=0.5, cull_screen_size=0.15, do_densify=1, split_by_screen=1,
             do_cull=1, cull_by_scale=1, cull_by_screen=1)


def inputs(seed=SEED):
    g = torch.Generator().manual_seed(seed)
    P = dict(means=torch.randn(N, 3, generator=g), log_scales=torch.log(torch.rand(N, 3, generator=g) * 0.03 + 0.002),
             quats=torch.randn(N, 4, generator=g), opacity_logit=torch.logit(torch.rand(N, 1, generator=g) * 0.5 + 0.01),
             sh_coeffs=torch.randn(N, K, 3, generator=g), features=torch.randn(N, D, generator=g))
    P["log_scales"][::41] = torch.log(torch.tensor(0.9))
    M = {k: (torch.randn(v.shape, generator=g), torch.rand(v.shape, generator=g)) for k, v in P.items()}
    st = dict(xys_grad_norm=torch.rand(N, generator=g) * 1.2e-5, vis_counts=torch.randint(1, 4, (N,), generator=g).float(),
              max_2dsize=torch.rand(N, generator=g) * 0.2)
    z = torch.randn((2 * N, 3), generator=g)
    return P, M, st, z


def build():
    for seed in range(SEED, SEED + 50):
        P, M, st, z = inputs(seed)
        Pr = dict(P); Pr["opacity_logit"] = P["opacity_logit"].reshape(-1)
        Mr = dict(M); Mr["opacity_logit"] = tuple(t.reshape(-1) for t in M["opacity_logit"])
        p, m, info = refine_oracle.refine(Pr, Mr, st["xys_grad_norm"], st["vis_counts"], st["max_2dsize"], RULES,
                                          lambda k: z[:k])
        if info["fragile"] == 0:
            break
    else:
        raise SystemExit("no seed without threshold-fragile decisions")
    out = dict(seed=np.array([seed]), z=z.numpy(), n_out=np.array([info["n_out"]]), n_cat=np.array([info["n_cat"]]))
    for k in refine_oracle.PARAMS:
        out["in_" + k] = P[k].numpy()
        out["in_m0_" + k], out["in_m1_" + k] = M[k][0].numpy(), M[k][1].numpy()
        out["out_" + k] = p[k].numpy().reshape((info["n_out"],) + tuple(P[k].shape[1:]))
        out["out_m0_" + k] = m[k][0].numpy().reshape(out["out_" + k].shape)
        out["out_m1_" + k] = m[k][1].numpy().reshape(out["out_" + k].shape)
    for k, v in st.items():
        out["st_" + k] = v.numpy()
    return out


if __name__ == "__main__":
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "refine_small.npz")
    ref, mine = np.load(path), build()
    same = all(np.array_equal(ref[k], v) for k, v in mine.items())
    print("oracle/refine_oracle.py", "reproduces" if same else "DIFFERS FROM", path)
    raise SystemExit(0 if same else 1)
