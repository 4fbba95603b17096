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

# ---- which part of bench.py's e2e step costs what: add the side-stream copies one at a time ----
print("--- e2e variants: ms per step (sync'd every step)")
copy_stream = torch.cuda.Stream(device=dev)
target_host = torch.randn((1, H, W, CP)).pin_memory()
tbuf = [torch.empty((1, H, W, CP), device=dev) for _ in range(2)]
rgb_host = torch.empty((1, H, W, 3)).pin_memory()
def run_variant(h2d, d2h, steps=30):
    main = torch.cuda.current_stream(dev)
    ready = [None, None]
    def one(q):
        for p in P.values():
            p.grad = None
        if h2d:
            with torch.cuda.stream(copy_stream):
                tbuf[(q + 1) & 1].copy_(target_host, non_blocking=True)
                ready[(q + 1) & 1] = copy_stream.record_event()
        vb = ViewBatch.from_cameras(cams, dev)
        out = render_views(*(P[k] for k in names), vb)
        if d2h:
            fwd_done = main.record_event()
            with torch.cuda.stream(copy_stream):
                copy_stream.wait_event(fwd_done)
                rgb_host.copy_(out["rgb"].detach(), non_blocking=True)
                out["rgb"].record_stream(copy_stream)
        if h2d and ready[q & 1] is not None:
            main.wait_event(ready[q & 1])
        img = out["image"]
        diff = img.detach() - tbuf[q & 1]
        loss = (diff * diff).mean()
        img.backward(diff * (2.0 / diff.numel()))
        loss_host.copy_(loss.detach().reshape(1), non_blocking=True)
        torch.cuda.synchronize()
    for q in range(5):
        one(q)
    t0 = time.perf_counter()
    for q in range(steps):
        one(q)
    return (time.perf_counter() - t0) / steps * 1e3
for h2d, d2h in ((False, False), (True, False), (False, True), (True, True)):
    print(f"h2d_prefetch={h2d!s:5} rgb_d2h={d2h!s:5}  {run_variant(h2d, d2h):.3f} ms/step")
