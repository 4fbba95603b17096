#!/usr/bin/env python
"""bench.py -- Mpixels/s of the Gaussian-splatting hot path (BASELINE.json metric).

  python bench.py --gpus N --steps K --warmup W            our arm (CUDA, sm_100a)
  python bench.py --impl reference --steps K --warmup W    the reference's PyTorch-CPU maths

A step is one render of the configured views -- SH + projection + binning + one blend of
RGB(3)+depth(1)+normal(3)+feature(D) -- forward AND backward to the leaf gradients (means,
log-scales, quats, opacity logits, SH coefficients, features), plus, at N>1, the NCCL all-reduce
of those gradients.  N=1 runs BASELINE.json configs[1]: 500k Gaussians, one 640x480 view, D=16.
At N>1 every rank renders its own view of the same (replicated) scene: weak scaling by view.
"""
from __future__ import annotations

import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

METRIC = "Mpixels/s rendered (RGB+depth+16-ch feature, fwd+bwd)"
UNIT = "Mpixels/s"


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return float(d.get("hbm_gbs", 6650.0)), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region: NVML polled every ~2 ms from a
    thread (the region lasts tens of ms, too short for `nvidia-smi -lms`), nvidia-smi as a fallback."""
    REASONS = {"hw_slowdown": 0x8, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20, "sw_power_cap": 0x4}

    def __init__(self, index):
        self.index = index
        self.samples, self.reason_bits, self.max_mhz = [], 0, None
        self._stop = threading.Event()
        self.thread = None
        self.mode = None

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            idx = int(vis.split(",")[self.index]) if vis and vis.split(",")[self.index].isdigit() else self.index
            h = pynvml.nvmlDeviceGetHandleByIndex(idx)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))

            def poll():
                while not self._stop.is_set():
                    try:
                        self.samples.append(float(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)))
                        self.reason_bits |= int(pynvml.nvmlDeviceGetCurrentClocksEventReasons(h))
                    except Exception:
                        pass
                    time.sleep(0.002)

            self.mode = "nvml"
            self.thread = threading.Thread(target=poll, daemon=True)
            self.thread.start()
        except Exception:
            self.mode = "nvidia-smi"
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index),
                                      "--query-gpu=clocks.sm,clocks.max.sm,clocks_event_reasons.active",
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=10).stdout
                parts = [x.strip() for x in out.strip().split(",")]
                self.samples.append(float(parts[0])); self.max_mhz = float(parts[1])
                self.reason_bits = int(parts[2], 16) if parts[2].startswith("0x") else 0
            except Exception:
                self.mode = None

    def stop(self):
        self._stop.set()
        if self.thread is not None:
            self.thread.join(timeout=1)
        sm = sorted(self.samples)
        reasons = [k for k, bit in self.REASONS.items() if self.reason_bits & bit]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": self.max_mhz, "samples": len(sm),
                "reasons": reasons, "source": self.mode}


# ---------------------------------------------------------------------------------------------
# the reference's CPU maths (oracle/torch_oracle.py is the port of gsplat's _torch_impl)
# ---------------------------------------------------------------------------------------------
def cpu_reference_step(sc, cam, D, n_tiles, threads):
    """One fwd+bwd of the reference maths on the host: full SH + projection + binning for all
    Gaussians (timed), blend fwd+bwd on `n_tiles` sampled tiles (timed); returns
    (t_geom, t_blend_sample, scale, n_sampled, n_total) with scale = frame pairs / sample pairs."""
    from oracle import torch_oracle as to
    torch.set_num_threads(threads)
    P = {k: v.clone().requires_grad_(True) for k, v in sc.items()}
    t0 = time.perf_counter()
    scales = torch.exp(P["log_scales"])
    q = P["quats"] / P["quats"].norm(dim=-1, keepdim=True)
    xys, depths, radii, conics, nth, _ = to.project_gaussians(P["means"], scales, 1.0, q, cam.viewmat, cam.fullmat,
                                                              cam.fx, cam.fy, cam.cx, cam.cy, cam.H, cam.W,
                                                              cam.tile_bounds)
    dirs = P["means"].detach() - cam.position
    rgbs = torch.clamp(to.spherical_harmonics(4, dirs, P["sh_coeffs"]) + 0.5, 0.0, 1.0)
    op = torch.sigmoid(P["opacity_logit"]).reshape(-1)
    R = to.quat_to_rotmat(P["quats"])
    idx = P["log_scales"].min(dim=-1)[1][..., None, None].expand(-1, 3, -1)
    normals = R.gather(2, idx).squeeze(dim=2)
    cols = torch.cat([rgbs, depths[:, None], normals, P["features"]], dim=1)
    _, _, ids_s, ranges = to.bin_and_sort(xys, depths, radii, nth, cam.tile_bounds)
    t1 = time.perf_counter()
    tiles, scale = sample_tiles(ranges, n_tiles)
    t1b = time.perf_counter()
    bg = torch.zeros(cols.shape[1]); bg[3] = 10.0
    g = torch.Generator().manual_seed(0)
    v_out = torch.randn((cam.H, cam.W, cols.shape[1]), generator=g)
    out, v_xys, v_con, v_op, v_col = to.rasterize_grads(xys.detach(), conics.detach(), op.detach(), cols.detach(),
                                                        ids_s, ranges, cam.H, cam.W, bg, v_out, tiles=tiles)
    t2 = time.perf_counter()
    torch.autograd.backward([xys, conics, op, cols], [v_xys, v_con, v_op, v_col])
    t3 = time.perf_counter()
    return (t1 - t0) + (t3 - t2), (t2 - t1b), scale, len(tiles), int(ranges.shape[0])


def sample_tiles(ranges, count):
    """Tiles at evenly spaced quantiles of the tile-list length; returns (tile ids, scale) where
    scale = total pixel-Gaussian pairs of the frame / pairs of the sample (the vectorised CPU blend
    costs time proportional to the list length of a tile)."""
    lens = (ranges[:, 1] - ranges[:, 0]).to(torch.float64)
    order = torch.argsort(lens)
    nz = order[lens[order] > 0]
    if nz.numel() == 0:
        return [0], 1.0
    pick = nz[torch.linspace(0, nz.numel() - 1, min(count, nz.numel())).round().long()]
    return pick.tolist(), float(lens.sum() / lens[pick].sum())


def cpu_sample_size(sc, cam, D, threads, budget_s):
    """Number of tiles whose CPU blend takes about `budget_s` seconds (from an 8-tile probe), and the tile total."""
    cpu_reference_step(sc, cam, D, 8, threads)       # first call pays thread-pool start-up and lazy initialisation
    tg, tb, scale, ns, nt = cpu_reference_step(sc, cam, D, 8, threads)
    full = max(tb * scale, 1e-6)                    # estimated blend time of the whole frame
    return int(min(nt, max(8, round(nt * budget_s / full)))), nt


def run_reference(args):
    from gaussiangrasper_b200 import scenes
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cfg = scenes.CONFIGS[1]
    threads = os.cpu_count() or 1
    sc = scenes.random_scene(cfg["n"], feature_dim=cfg["D"], seed=1235)
    cam = scenes.orbit_cameras(1, cfg["W"], cfg["H"])[0]
    times = []
    # bounded sample per step: the whole run (warm-up + steps) stays within ~3 minutes, one step within ~12 s
    n_sample, _ = cpu_sample_size(sc, cam, cfg["D"], threads, min(12.0, 170.0 / max(1, args.warmup + args.steps)))
    for s in range(args.warmup + args.steps):
        tg, tb, scale, ns, nt = cpu_reference_step(sc, cam, cfg["D"], n_sample, threads)
        if s >= args.warmup:
            times.append(tg + tb * scale)
    t = sum(times) / len(times)
    mpix = cfg["W"] * cfg["H"] / 1e6
    val = mpix / t
    sample = (f"each step: SH+projection+binning of all {cfg['n']} Gaussians fwd+bwd (timed in full) + blend "
              f"fwd+bwd of {ns} of {nt} tiles (length quantiles), blend time scaled by frame pairs / sample pairs "
              f"= {scale:.1f}")
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": t * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": cfg["name"], "gaussians": cfg["n"], "image": [cfg["W"], cfg["H"]], "views": 1,
                       "channels": 7 + cfg["D"]},
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------
def run_ours(args):
    import torch.distributed as dist

    from gaussiangrasper_b200 import _lib, scenes
    from gaussiangrasper_b200.render import ViewBatch, render_views

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py (our arm) needs a CUDA device; there is no CPU fallback"
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    _lib.check(_lib.load().gg_check_device(), "gg_check_device")

    cfg = scenes.CONFIGS[args.config]
    n, W, H, D = cfg["n"], cfg["W"], cfg["H"], (cfg["D"] if args.feat < 0 else args.feat)
    # config 2 is a fixed 64-view batch split over the ranks (strong scaling); the others fix the work per GPU
    strong = args.config == 2
    V = args.views if args.views else (1 if args.config == 1 else (max(1, cfg["views"] // world) if strong else cfg["views"]))
    chunk = min(V, args.chunk)
    C = 7 + D
    CP = (C + 3) // 4 * 4
    sc = scenes.random_scene(n, feature_dim=D, seed=1234 + args.config)
    names = ("means", "log_scales", "quats", "opacity_logit", "sh_coeffs", "features")
    P = {k: sc[k].to(dev).requires_grad_(cfg["backward"]) for k in names}
    total_views = V * world
    cams = scenes.orbit_cameras(V, W, H, first=rank * V, total=max(total_views, 8))
    views = ViewBatch.from_cameras(cams[:chunk], dev)
    g = torch.Generator().manual_seed(100 + rank)
    v_img = torch.randn((chunk, H, W, CP), generator=g).to(dev) if cfg["backward"] else None
    # views are rendered `chunk` at a time (one projection + one sort + one blend launch per chunk)
    chunks = [ViewBatch.from_cameras(cams[i:i + chunk], dev) for i in range(0, V, chunk)]
    from gaussiangrasper_b200.distributed import FactoredExchange, GradientBucket
    from gaussiangrasper_b200.training import pixel_loss
    from gaussiangrasper_b200 import ops as _ops
    sh_degree = _ops.sh_degree_from_bases(P["sh_coeffs"].shape[1])
    # gradient exchange of a multi-GPU training step: SH gradient as per-view factors (all-gather) + one
    # all-reduce of the other leaves; --exchange allreduce = one all-reduce of everything
    # (worth it while the gathered factors, 3 floats per view and Gaussian over ALL views, stay well below the
    # 3K-float product: at 64 views the all-reduce of the product is the cheaper exchange again)
    factored = (world > 1 and cfg["backward"] and len(chunks) == 1 and args.exchange == "factored"
                and world * chunks[0].n_views * 2 <= P["sh_coeffs"].shape[1])
    ex = FactoredExchange(P, chunks[0].n_views) if factored else None
    # every rank can name every rank's cameras (a shared sampler): no collective for the camera centres
    all_pos = torch.stack([c.position for r in range(world) for c in
                           scenes.orbit_cameras(V, W, H, first=r * V, total=max(total_views, 8))]).float().to(dev) \
        if factored else None
    bucket = GradientBucket(P) if (world > 1 and cfg["backward"] and not factored) else None

    def step_dropin():
        from gaussiangrasper_b200.reference_flow import get_outputs
        for p in P.values():
            p.grad = None
        for cam_i in cams:
            o = get_outputs(P["means"], P["log_scales"], P["quats"], P["opacity_logit"], P["sh_coeffs"], P["features"],
                            cam_i)
            if cfg["backward"]:
                loss = ((o["rgb"] * v_img[0, ..., 0:3]).sum() + (o["depth"] * v_img[0, ..., 3:4]).sum() +
                        (o["normal"] * v_img[0, ..., 4:7]).sum() + (o["feature"] * v_img[0, ..., 7:7 + D]).sum())
                loss.backward()
        return o["rgb"]

    def step():
        if args.path == "dropin":
            return step_dropin()
        for p in P.values():
            p.grad = None
        out = None
        # one launch group per step: the backward writes the leaf gradients straight into the flat
        # all-reduce buffer; several chunks accumulate in .grad and are packed afterwards
        direct = bucket is not None and len(chunks) == 1
        for vb in chunks:
            holder = ex.holder() if ex is not None else ({"grad_out": bucket.unpack()} if direct else None)
            with torch.set_grad_enabled(cfg["backward"]):
                out = render_views(P["means"], P["log_scales"], P["quats"], P["opacity_logit"], P["sh_coeffs"],
                                   P["features"], vb, holder=holder)
            if cfg["backward"]:
                out["image"].backward(v_img[:vb.n_views])  # leaf gradients accumulate over the chunks
        if ex is not None:
            return ex.exchange(P["means"], chunks[0].positions, sh_degree, 4, holder, all_pos)["sh_coeffs"]
        if bucket is not None:
            if not direct:
                bucket.pack({k: P[k].grad for k in names})
            bucket.all_reduce()
            return bucket.flat
        return out["image"]

    # at least 10 untimed steps: the first collectives / allocations of a fresh process settle here
    n_warm = max(args.warmup, 10 if args.path == "fused" else 3)
    for _ in range(n_warm):
        step()
    torch.cuda.synchronize()

    # --- timed region: exactly K steps, device-timed, max over ranks -------------------------
    sampler = ClockSampler(local) if rank == 0 else None
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    if sampler:
        sampler.start()
    l0 = _lib.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    # inside the timed region only the two blend entry points carry an event pair (the roofline's
    # launch duration is measured live, here); the full per-entry-point breakdown is a separate pass
    with _lib.profile(only=("gg_blend_fwd", "gg_blend_bwd")) as prof_timed:
        e0.record()
        for _ in range(args.steps):
            step()
        e1.record()
        torch.cuda.synchronize()
        blend_calls = prof_timed.ms()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    launches = _lib.launch_count() - l0
    clocks = sampler.stop() if sampler else None
    ms = e0.elapsed_time(e1)
    # per-entry-point breakdown: a separate, instrumented pass (an event pair around every C-ABI call
    # perturbs the stream, so it stays out of the timed region above)
    prof_steps = min(args.steps, 10)
    with _lib.profile() as prof:
        for _ in range(prof_steps):
            step()
        torch.cuda.synchronize()
        per_call = prof.ms()
    per_call.update(blend_calls)   # the dominant kernels: from the timed region itself
    t_ms = torch.tensor([ms], device=dev)
    if world > 1:
        dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
    ms_per_step = float(t_ms.item()) / args.steps
    mpix_per_step = total_views * W * H / 1e6
    value = mpix_per_step / (ms_per_step * 1e-3)

    # --- e2e: host buffers in, host results out, every step ---------------------------------
    target_host = torch.randn((chunk, H, W, CP), generator=g).pin_memory() if cfg["backward"] else None
    rgb_host = torch.empty((chunk, H, W, 3)).pin_memory()
    loss_host = torch.empty((1,)).pin_memory()

    h2d_stream = torch.cuda.Stream(device=dev)
    d2h_stream = torch.cuda.Stream(device=dev)
    # Supervision images are double-buffered on the device.  The copy for chunk q+1 is enqueued after chunk
    # q's backward has been launched (what a prefetching data loader does): every step still moves its full
    # H2D bytes inside the timed region, but while the GPU runs the two long blend kernels -- a 30 MB PCIe
    # read that overlaps the launch-bound front of the forward costs ~0.14 ms per step of launch latency.
    target_dev = [torch.empty((chunk, H, W, CP), dtype=torch.float32, device=dev) for _ in range(2)] if cfg["backward"] else None
    pf = {"q": 0, "ready": [None, None], "consumed": [None, None]}

    def prefetch_target(q, nv):
        b = q & 1
        with torch.cuda.stream(h2d_stream):
            if pf["consumed"][b] is not None:
                h2d_stream.wait_event(pf["consumed"][b])                     # chunk q-2 has read this buffer
            target_dev[b][:nv].copy_(target_host[:nv], non_blocking=True)   # H2D: supervision images (pinned)
            pf["ready"][b] = h2d_stream.record_event()

    def e2e_step():
        """One user-level training step from HOST buffers: cameras + supervision images go host->device,
        the rendered rgb and the loss come back device->host, all inside the timed region.  The big
        copies run on side streams so they overlap the render (as a data loader would prefetch); the
        step returns when its loss and its rgb image are in host memory."""
        for p in P.values():
            p.grad = None
        main = torch.cuda.current_stream(dev)
        loss = None
        rgb_done = None
        for ci in range(0, V, chunk):
            cc = cams[ci:ci + chunk]
            nv = len(cc)
            q = pf["q"]
            if cfg["backward"] and pf["ready"][q & 1] is None:
                prefetch_target(q, nv)                                      # very first chunk only
            # H2D: cameras (pinned).  Like the images, the next chunk's cameras are staged once this chunk's
            # kernels are enqueued (below), so the copy is not on the launch-bound front of the step.
            vb = pf.pop("vb", None)
            if vb is None:
                vb = ViewBatch.from_cameras(cc, dev)                        # very first chunk only
            holder = ex.holder() if ex is not None else None
            with torch.set_grad_enabled(cfg["backward"]):
                out = render_views(P["means"], P["log_scales"], P["quats"], P["opacity_logit"], P["sh_coeffs"],
                                   P["features"], vb, holder=holder)
            fwd_done = main.record_event()
            if cfg["backward"]:
                main.wait_event(pf["ready"][q & 1])
                img = out["image"]
                l, grad = pixel_loss(img, target_dev[q & 1][:nv], "l2")    # L2 loss against the host-fed targets:
                pf["consumed"][q & 1] = main.record_event()                # loss and gradient in one kernel
                pf["ready"][q & 1] = None
                img.backward(grad)
            else:
                l = out["alpha"].mean()
            loss = l if loss is None else loss + l
            # the copies are enqueued once the step's kernels are: they run under the backward
            with torch.cuda.stream(d2h_stream):
                d2h_stream.wait_event(fwd_done)
                rgb_host[:nv].copy_(out["rgb"].detach(), non_blocking=True)  # D2H: rendered rgb
                out["rgb"].record_stream(d2h_stream)
                rgb_done = d2h_stream.record_event()
            nxt = ci + chunk if ci + chunk < V else 0
            if cfg["backward"]:
                prefetch_target(q + 1, min(chunk, V - nxt))                 # next chunk (of this or the next step)
            pf["vb"] = ViewBatch.from_cameras(cams[nxt:nxt + chunk], dev)   # next chunk's cameras
            pf["q"] = q + 1
        if ex is not None:
            ex.exchange(P["means"], vb.positions, sh_degree, 4, holder, all_pos)
        elif bucket is not None:
            bucket.pack({k: P[k].grad for k in names})
            bucket.all_reduce()
        loss_host.copy_(loss.detach().reshape(1), non_blocking=True)    # D2H: loss
        main.synchronize()
        rgb_done.synchronize()
        return float(loss_host[0])

    for _ in range(2):
        e2e_step()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        e2e_step()
    torch.cuda.synchronize()
    t_e2e = torch.tensor([time.perf_counter() - t0], device=dev)
    if world > 1:
        dist.all_reduce(t_e2e, op=dist.ReduceOp.MAX)
    e2e_val = mpix_per_step * args.steps / float(t_e2e.item())
    if args.trace and rank == 0:
        _trace_timeline(e2e_step, step, args.trace)
    h2d = (chunk * H * W * CP * 4 if cfg["backward"] else 0) * (V // chunk) + V * 35 * 4
    d2h = V * H * W * 3 * 4 + 4

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # --- roofline of the dominant kernel ------------------------------------------------------
    hbm_peak, peak_src = load_peaks()
    avg = {k: sum(v) / len(v) for k, v in per_call.items() if v}
    share = {k: sum(v) / (args.steps if k in blend_calls else prof_steps) for k, v in per_call.items()}
    stats = torch.zeros(1, dtype=torch.int64, device=dev)
    with torch.no_grad():
        o = render_views(*(P[k].detach() for k in names), views, stats=stats)  # first chunk only
    torch.cuda.synchronize()
    pairs = int(stats.item())
    binning_m = None
    # measured FP32 FMA roof (same process, same clocks)
    lib = _lib.load()
    probe = torch.zeros(1, device=dev)
    fl = ctypes.c_double(0.0)
    for _ in range(2):
        lib.gg_bench_fma(148 * 8, 20000, probe.data_ptr(), ctypes.byref(fl), torch.cuda.current_stream().cuda_stream)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    lib.gg_bench_fma(148 * 8, 20000, probe.data_ptr(), ctypes.byref(fl), torch.cuda.current_stream().cuda_stream)
    b.record()
    torch.cuda.synchronize()
    fma_tflops = fl.value / (a.elapsed_time(b) * 1e-3) / 1e12
    dom = max(avg, key=lambda k: share[k]) if avg else None
    flops_fwd = pairs * (12 + 2 * C)
    flops_bwd = pairs * (30 + 8 * C)
    roof = None
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath) and args.config == 1 and V == 1:
        with open(tpath) as f:
            traffic = json.load(f).get(dom, {}).get("dram_bytes_per_launch")
    if dom in ("gg_blend_bwd", "gg_blend_fwd"):
        fl_k = flops_bwd if dom == "gg_blend_bwd" else flops_fwd
        ach = fl_k / (avg[dom] * 1e-3) / 1e12
        roof = {"kernel": dom, "bound": "fp32", "achieved": ach, "peak": fma_tflops, "unit": "TFLOP/s",
                "frac": ach / fma_tflops, "traffic": traffic, "traffic_source": "ncu --set full capture, profiles/traffic.json",
                "peak_source": "FMA probe kernel timed in this run",
                "algorithmic_flops": fl_k, "pairs": pairs, "avg_launch_ms": avg[dom]}
    elif dom is not None:
        roof = {"kernel": dom, "bound": "hbm", "achieved": None, "peak": hbm_peak, "unit": "GB/s", "frac": None,
                "traffic": None, "peak_source": peak_src, "avg_launch_ms": avg[dom]}
    # HBM-bound stages against the measured copy bandwidth
    # algorithmic bytes per (view, Gaussian) of the HBM-bound stages
    prep_b = (12 + 12 + 16 + 4 + 300 + 4 * D) + (32 + 4 * CP + 12) + 44      # in + out + phase-2 re-read
    prep_bwd_b = (44 + 32 + 12 + 4 + 32 + 4 * CP) + (12 + 12 + 16 + 4 + 300 + 4 * D)
    bytes_model = {"gg_prepare_views": n * V * prep_b, "gg_prepare_views_bwd": n * V * prep_bwd_b}
    hbm_stages = {}
    for k, bts in bytes_model.items():
        if k in share:
            gbs = bts / (share[k] * 1e-3) / 1e9   # all launches of the stage in one step
            hbm_stages[k] = {"GB/s": gbs, "frac": gbs / hbm_peak, "ms_per_step": share[k], "bytes_per_step": bts}

    # --- CPU baseline (rank 0, N=1 only): the reference maths on the host cores ---------------
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        sc_cpu = {k: sc[k] for k in names}
        cam0 = cams[0]
        n_sample, _ = cpu_sample_size(sc_cpu, cam0, D, threads, 12.0)   # ~12 s of CPU blend work
        tg, tb, scale, ns, nt = cpu_reference_step(sc_cpu, cam0, D, n_sample, threads)
        t_full = tg + tb * scale
        cpu = {"value": (W * H / 1e6) / t_full, "unit": UNIT, "cores": threads, "kind": "port",
               "sample": (f"torch-CPU port of the reference maths, 1 step: SH+projection+binning fwd+bwd of all {n} "
                          f"Gaussians ({tg:.2f} s) + blend fwd+bwd of {ns}/{nt} tiles at length quantiles ({tb:.2f} s, "
                          f"scaled by frame pairs / sample pairs = {scale:.1f})")}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": n_warm,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong" if strong else "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": cfg["name"], "gaussians": n, "image": [W, H], "views_per_gpu": V, "views_per_launch": chunk,
                   "channels": C, "backward": cfg["backward"], "parallelism": f"view-sharded x{world}", "path": args.path,
                   "gradient_exchange": ("sh factors all-gather + all-reduce of the other leaves" if ex is not None else
                                         "all-reduce" if bucket is not None else "none"),
                   "l2": "inputs larger than L2 (parameters+gradients 2x%.0f MB per step)" % (n * (86 + D) * 4 / 1e6)},
        "e2e": {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
        "gpu_launches": launches,
        "clocks": clocks,
        "roofline": roof,
        "hbm_stages": hbm_stages,
        "stage_ms_per_step": share,
        "fma_probe_tflops": fma_tflops,
        "cpu_baseline": cpu,
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def _trace_timeline(e2e_step, step, path):
    """Diagnostic only: kernel/memcpy intervals of 2 e2e steps and 2 async steps, with the idle gaps between them."""
    from torch.profiler import ProfilerActivity, profile
    lines = []
    for name, fn in (("e2e_step", e2e_step), ("device_step", step)):
        fn(); torch.cuda.synchronize()
        with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
            fn(); fn()
            torch.cuda.synchronize()
        ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
        ev.sort(key=lambda e: e.time_range.start)
        if not ev:
            lines.append(f"== {name}: no device events (CUPTI unavailable?)"); continue
        t0 = ev[0].time_range.start
        busy = 0.0; prev_end = t0
        lines.append(f"== {name} x2: start_us dur_us gap_before_us name")
        for e in ev:
            st, en = e.time_range.start, e.time_range.end
            gap = st - prev_end
            lines.append(f"{st - t0:10.1f} {en - st:8.1f} {gap:8.1f}  {e.name[:90]}")
            busy += en - st
            prev_end = max(prev_end, en)
        lines.append(f"== {name}: span {prev_end - t0:.1f} us, sum of intervals {busy:.1f} us")
    with open(path, "w") as f:
        f.write("\n".join(lines) + "\n")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", type=int, default=1, help="index into BASELINE.json configs (default 1)")
    ap.add_argument("--views", type=int, default=0, help="views per GPU per step (default: config's)")
    ap.add_argument("--chunk", type=int, default=8, help="views rendered per launch group")
    ap.add_argument("--feat", type=int, default=-1, help="feature channels D (default: the config's)")
    ap.add_argument("--path", default="fused", choices=["fused", "dropin"],
                    help="fused: render_views; dropin: the reference's 1 projection + SH + 4 rasterize calls")
    ap.add_argument("--exchange", default="factored", choices=["factored", "allreduce"],
                    help="N>1 training steps: SH gradient exchanged as per-view factors (default) or one all-reduce of "
                         "every leaf gradient")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--trace", default="", help="diagnostic: write a GPU timeline (kernels + idle gaps) of one e2e and one "
                                                "device-timed step to this file (torch.profiler; not part of the measurement)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
