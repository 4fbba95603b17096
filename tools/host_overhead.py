"""Diagnostic: host-side time of the pieces of one training step (config 1), GPU idle at each start."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gaussiangrasper_b200 import scenes, _lib
from gaussiangrasper_b200.render import ViewBatch, render_views

dev = torch.device("cuda:0")
cfg = scenes.CONFIGS[1]
sc = scenes.random_scene(cfg["n"], cfg["D"], seed=0)
P = {k: v.to(dev).requires_grad_(True) for k, v in sc.items()}
W, H = cfg["W"], cfg["H"]
cams = scenes.orbit_cameras(1, W, H)
names = ["means", "log_scales", "quats", "opacity_logit", "sh_coeffs", "features"]
CP = (7 + cfg["D"] + 3) // 4 * 4
v_img = torch.randn((1, H, W, CP), device=dev)
T = {}
def tick(name, t0):
    T.setdefault(name, []).append((time.perf_counter() - t0) * 1e6)
for it in range(30):
    for p in P.values():
        p.grad = None
    torch.cuda.synchronize()
    t0 = time.perf_counter(); vb = ViewBatch.from_cameras(cams, dev); tick("from_cameras(host)", t0)
    torch.cuda.synchronize()
    t0 = time.perf_counter(); out = render_views(*(P[k] for k in names), vb); tick("render_views fwd returns (host, incl. M wait)", t0)
    torch.cuda.synchronize(); tick("fwd complete on GPU", t0)
    t0 = time.perf_counter(); out["image"].backward(v_img); tick("backward returns (host)", t0)
    torch.cuda.synchronize(); tick("bwd complete on GPU", t0)
    t0 = time.perf_counter(); torch.cuda.synchronize(); tick("empty sync", t0)
for k, v in T.items():
    v = sorted(v[5:]); print(f"{k:50s} median {v[len(v)//2]:8.1f} us   min {v[0]:8.1f}")

# ---- e2e-like step (sync at the end of every step): host time points and GPU event times ----
print("--- e2e-like step: host time points (us since step start) and GPU events (us since first event)")
target = torch.randn((1, H, W, CP), device=dev)
loss_host = torch.empty((1,)).pin_memory()
HT, GT = {}, {}
for it in range(40):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    ev = {}
    def mark(name):
        HT.setdefault(name, []).append((time.perf_counter() - t0) * 1e6)
        e = torch.cuda.Event(enable_timing=True); e.record(); ev[name] = e
    for p in P.values():
        p.grad = None
    mark("0 start")
    vb = ViewBatch.from_cameras(cams, dev)
    mark("1 cameras")
    out = render_views(*(P[k] for k in names), vb)
    mark("2 fwd enqueued")
    img = out["image"]
    diff = img.detach() - target
    loss = (diff * diff).mean()
    g = diff * (2.0 / diff.numel())
    mark("3 loss enqueued")
    img.backward(g)
    mark("4 bwd enqueued")
    loss_host.copy_(loss.detach().reshape(1), non_blocking=True)
    torch.cuda.synchronize()
    HT.setdefault("5 synced", []).append((time.perf_counter() - t0) * 1e6)
    for k, e in ev.items():
        GT.setdefault(k, []).append(ev["0 start"].elapsed_time(e) * 1e3)
for k in sorted(HT):
    v = sorted(HT[k][5:]); g = sorted(GT[k][5:]) if k in GT else None
    print(f"{k:20s} host {v[len(v)//2]:8.1f}   gpu-event {g[len(g)//2] if g else float('nan'):8.1f}")
