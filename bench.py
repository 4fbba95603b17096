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
# what both arms print as `config` (same keys, same values: the reference arm runs our arm's workload)
# ---------------------------------------------------------------------------------------------
SETUP_STEPS = 3   # untimed initialisation before the warm-up: first-call capacity measurement of the binning,
                  # allocator growth, NCCL communicator set-up.  Not counted as warm-up, never timed.


def _exchange_mode(multicast: bool, world: int, bucket_bytes: int) -> str:
    """Which form of the exchange kernel csrc/exchange.cu picks for the bucket reduction (same rule as there)."""
    mode = os.environ.get("GG_NVLS_MODE", "")
    if mode not in ("0", "1", "2") or (mode in ("0", "2") and not multicast):
        mode = "0" if (multicast and world >= 4 and bucket_bytes >= (80 << 20)) else "1"
    if mode == "0":
        return "NVLS multimem.ld_reduce / multimem.st"
    if mode == "2":
        return "peer loads + multimem.st"
    return "peer loads / stores over NVLink, fixed summation order" + ("; multicast mapping available" if multicast else "")


def bench_config(args, world):
    from gaussiangrasper_b200 import scenes
    cfg = scenes.CONFIGS[args.config]
    D = cfg["D"] if args.feat < 0 else args.feat
    strong = args.config == 2
    V = args.views if args.views else (1 if args.config == 1 else (max(1, cfg["views"] // world) if strong else cfg["views"]))
    chunk = min(V, args.chunk)
    n = cfg["n"]
    n_chunks = (V + chunk - 1) // chunk
    factored = (world > 1 and cfg["backward"] and n_chunks == 1 and args.exchange == "factored" and world * chunk * 2 <= 25)
    exchange = ("sh factors all-gather + all-reduce of the other leaves" if factored else
                "all-reduce" if (world > 1 and cfg["backward"]) else "none")
    return cfg, D, V, chunk, strong, {
        "workload": cfg["name"], "gaussians": n, "image": [cfg["W"], cfg["H"]], "views_per_gpu": V,
        "views_per_launch": chunk, "channels": 7 + D, "backward": cfg["backward"],
        "parallelism": f"view-sharded x{world}", "path": args.path, "gradient_exchange": exchange,
        "l2": "inputs larger than L2 (parameters+gradients 2x%.0f MB per step)" % (n * (86 + D) * 4 / 1e6),
        "setup_steps": SETUP_STEPS}


# ---------------------------------------------------------------------------------------------
# the reference's CPU maths (oracle/torch_oracle.py is the port of gsplat's _torch_impl)
# ---------------------------------------------------------------------------------------------
def cpu_reference_step(sc, cam, D, tiles, threads, backward=True):
    """One step of the reference maths on the host: SH + projection + binning of ALL Gaussians, blend of the
    tiles in `tiles` (None = the whole frame), backward to the leaves when `backward`.
    Returns (seconds, pixels blended, tiles blended, tiles of the frame)."""
    from oracle import torch_oracle as to
    torch.set_num_threads(threads)
    P = {k: v.clone().requires_grad_(backward) for k, v in sc.items()}
    t0 = time.perf_counter()
    with torch.set_grad_enabled(backward):
        scales = torch.exp(P["log_scales"])
        q = P["quats"] / P["quats"].norm(dim=-1, keepdim=True)
        xys, depths, radii, conics, nth, _ = to.project_gaussians(P["means"], scales, 1.0, q, cam.viewmat, cam.fullmat,
                                                                  cam.fx, cam.fy, cam.cx, cam.cy, cam.H, cam.W,
                                                                  cam.tile_bounds)
        dirs = P["means"].detach() - cam.position
        rgbs = torch.clamp(to.spherical_harmonics(4, dirs, P["sh_coeffs"]) + 0.5, 0.0, 1.0)
        op = torch.sigmoid(P["opacity_logit"]).reshape(-1)
        parts = [rgbs, depths[:, None]]
        if D > 0 or backward:
            R = to.quat_to_rotmat(P["quats"])
            idx = P["log_scales"].min(dim=-1)[1][..., None, None].expand(-1, 3, -1)
            parts += [R.gather(2, idx).squeeze(dim=2), P["features"]]
        cols = torch.cat(parts, dim=1)
        _, _, ids_s, ranges = to.bin_and_sort(xys, depths, radii, nth, cam.tile_bounds)
    n_tiles = int(ranges.shape[0])
    tile_list = list(range(n_tiles)) if tiles is None else list(tiles)
    bg = torch.zeros(cols.shape[1]); bg[3] = 10.0
    if backward:
        v_out = torch.randn((cam.H, cam.W, cols.shape[1]), generator=torch.Generator().manual_seed(0))
        out, v_xys, v_con, v_op, v_col = to.rasterize_grads(xys.detach(), conics.detach(), op.detach(), cols.detach(),
                                                            ids_s, ranges, cam.H, cam.W, bg, v_out, tiles=tile_list)
        torch.autograd.backward([xys, conics, op, cols], [v_xys, v_con, v_op, v_col])
    else:
        with torch.no_grad():
            to.rasterize(xys, conics, op, cols, ids_s, ranges, cam.H, cam.W, bg, tiles=tile_list)
    dt = time.perf_counter() - t0
    tx = cam.tile_bounds[0]
    pixels = sum((min(16, cam.W - 16 * (t % tx))) * (min(16, cam.H - 16 * (t // tx))) for t in tile_list)
    return dt, pixels, len(tile_list), n_tiles


def spread_tiles(n_tiles, count):
    """`count` tile ids spread evenly over the frame (deterministic)."""
    if count >= n_tiles:
        return None
    step = n_tiles / count
    return sorted({int(i * step) for i in range(count)})


def cpu_plan(sc, cam, D, threads, budget_s, backward=True):
    """Tiles one CPU step blends so that it takes about `budget_s` seconds: the whole frame when that fits, else an
    evenly spread subset.  Sized from a warm 16-tile probe (its first call pays thread-pool start-up).  The
    throughput of a step is ALWAYS (pixels it blended) / (seconds it took): nothing is extrapolated."""
    nt = cam.tile_bounds[0] * cam.tile_bounds[1]
    cpu_reference_step(sc, cam, D, spread_tiles(nt, 16), threads, backward)
    dt16, _, _, n_tiles = cpu_reference_step(sc, cam, D, spread_tiles(nt, 16), threads, backward)
    dt32, _, _, _ = cpu_reference_step(sc, cam, D, spread_tiles(nt, 48), threads, backward)
    per_tile = max((dt32 - dt16) / 32.0, 1e-6)
    fixed = max(dt16 - 16 * per_tile, 0.0)
    full = fixed + per_tile * n_tiles
    if full <= budget_s:
        return None, n_tiles, full
    return spread_tiles(n_tiles, max(16, int((budget_s - fixed) / per_tile))), n_tiles, full


def run_reference(args):
    from gaussiangrasper_b200 import scenes
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    world = int(os.environ.get("WORLD_SIZE", "1"))
    cfg, D, V, chunk, strong, config = bench_config(args, world)
    threads = os.cpu_count() or 1
    sc = scenes.random_scene(cfg["n"], feature_dim=D, seed=1234 + args.config)
    cam = scenes.orbit_cameras(1, cfg["W"], cfg["H"], total=8)[0]
    # the whole run (warm-up + steps) stays within ~5 minutes; one step blends the full frame when that fits
    budget = min(14.0, 300.0 / max(1, args.warmup + args.steps))
    tiles, n_tiles, est_full = cpu_plan(sc, cam, D, threads, budget, cfg["backward"])
    times, pix = [], 0
    for s_i in range(args.warmup + args.steps):
        dt, pixels, nt, _ = cpu_reference_step(sc, cam, D, tiles, threads, cfg["backward"])
        if s_i >= args.warmup:
            times.append(dt)
            pix = pixels
    t = sum(times) / len(times)
    val = pix / 1e6 / t
    sample = (f"each step: SH+projection+binning of all {cfg['n']} Gaussians + blend of "
              f"{'the whole frame' if tiles is None else f'{len(tiles)} of {n_tiles} tiles spread over the frame'} "
              f"({pix} pixels){', forward and backward' if cfg['backward'] else ''}; value = pixels blended / seconds "
              f"taken, nothing extrapolated{'' if tiles is None else '; the per-Gaussian stage is not sampled, so a partial frame understates the CPU rate'}"
              " (one view; the CPU port has no multi-view path)")
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": t * 1e3, "higher_is_better": True,
            "scaling": "strong" if strong else "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": config,
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def cpu_baseline_leg(sc, cam, D, n, W, H, backward):
    """The `cpu_baseline` object of our arm's line (rank 0, N=1): ~10-30 s of CPU work in total.
    (ii) of SURVEY 8d on the benchmarked workload, plus (i) the literal-loop crop and config 0's plumbing case."""
    from gaussiangrasper_b200 import scenes
    from oracle import torch_oracle as to
    threads = os.cpu_count() or 1
    tiles, n_tiles, _ = cpu_plan(sc, cam, D, threads, 14.0, backward)
    dt, pixels, nt, _ = cpu_reference_step(sc, cam, D, tiles, threads, backward)
    cpu = {"value": pixels / 1e6 / dt, "unit": UNIT, "cores": threads, "kind": "port",
           "sample": (f"torch-CPU port of the reference maths (oracle/torch_oracle.py), 1 step of {dt:.1f} s: SH+projection+"
                      f"binning of all {n} Gaussians + blend of "
                      f"{'the whole frame' if tiles is None else f'{nt} of {n_tiles} tiles spread over the frame'} "
                      f"({pixels} pixels){', forward and backward' if backward else ''}; pixels blended / seconds, "
                      "nothing extrapolated")}
    # BASELINE configs[0]: 50k Gaussians, 640x480, RGB+depth forward on the CPU (plumbing case)
    c0 = scenes.CONFIGS[0]
    sc0 = scenes.random_scene(c0["n"], feature_dim=0, seed=1234)
    cam0 = scenes.orbit_cameras(1, c0["W"], c0["H"], total=8)[0]
    dt0, pix0, _, _ = cpu_reference_step(sc0, cam0, 0, None, threads, backward=False)
    cpu["config0"] = {"workload": c0["name"], "value": pix0 / 1e6 / dt0, "unit": UNIT, "seconds": dt0,
                      "what": "vectorised torch-CPU forward, RGB+depth, whole frame"}
    # the literal per-pixel / per-entry Python loop on a fixed 64x48 crop of config 0
    with torch.no_grad():
        q = sc0["quats"] / sc0["quats"].norm(dim=-1, keepdim=True)
        xys, depths, radii, conics, nth, _ = to.project_gaussians(sc0["means"], sc0["log_scales"].exp(), 1.0, q,
                                                                  cam0.viewmat, cam0.fullmat, cam0.fx, cam0.fy, cam0.cx,
                                                                  cam0.cy, cam0.H, cam0.W, cam0.tile_bounds)
        _, _, ids_s, ranges = to.bin_and_sort(xys, depths, radii, nth, cam0.tile_bounds)
        rgbs = torch.clamp(to.spherical_harmonics(4, sc0["means"] - cam0.position, sc0["sh_coeffs"]) + 0.5, 0.0, 1.0)
        cols = torch.cat([rgbs, depths[:, None]], dim=1)
        bg = torch.tensor([0.0, 0.0, 0.0, 10.0])
        t0 = time.perf_counter()
        _, pairs = to.rasterize_literal(xys, conics, torch.sigmoid(sc0["opacity_logit"]), cols, ids_s, ranges, cam0.H,
                                        cam0.W, bg, crop=(288, 216, 64, 48))
        dtl = time.perf_counter() - t0
    cpu["literal_loop_crop"] = {"crop": [288, 216, 64, 48], "pairs": pairs, "seconds": dtl, "pairs_per_s": pairs / dtl,
                                "Mpixels/s": 64 * 48 / 1e6 / dtl, "cores": 1,
                                "what": "per-pixel, per-entry Python loop (the shape of gsplat's _torch_impl rasterizer)"}
    return cpu


# ---------------------------------------------------------------------------------------------
def run_ours(args, sub=False):
    """One measurement; returns the JSON line as a dict on rank 0 (None elsewhere).  sub: a second workload inside
    the same process group (BASELINE configs[3] reported next to configs[1] at N > 1)."""
    import torch.distributed as dist

    from gaussiangrasper_b200 import _lib, scenes
    from gaussiangrasper_b200.render import ViewBatch, render_views

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py (our arm) needs a CUDA device; there is no CPU fallback"
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    if world > 1 and not sub:
        dist.init_process_group("nccl", device_id=dev)
    _lib.check(_lib.load().gg_check_device(), "gg_check_device")

    # config 2 is a fixed 64-view batch split over the ranks (strong scaling); the others fix the work per GPU
    cfg, D, V, chunk, strong, config = bench_config(args, world)
    n, W, H = cfg["n"], cfg["W"], cfg["H"]
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
    ex, transport = None, ("nccl" if world > 1 and cfg["backward"] else "none")
    if factored:
        if args.transport in ("auto", "nvls"):
            try:   # one kernel of this library over symmetric memory (multimem through the NVSwitch)
                from gaussiangrasper_b200.distributed import NvlsExchange
                ex = NvlsExchange(P, chunks[0].n_views)
                transport = "gg_nvls_exchange kernel over symmetric memory (%s)" % _exchange_mode(ex.multicast, world, ex.bucket.flat.numel() * 4)
            except Exception as e:   # no symmetric memory on this box: NCCL collectives
                if args.transport == "nvls":
                    raise
                ex, transport = None, f"nccl (symmetric memory unavailable: {type(e).__name__}: {str(e)[:120]})"
        if ex is None:
            ex = FactoredExchange(P, chunks[0].n_views)
    # every rank can name every rank's cameras (a shared sampler): no collective for the camera centres
    all_pos = torch.stack([c.position for r in range(world) for c in
                           scenes.orbit_cameras(V, W, H, first=r * V, total=max(total_views, 8))]).float().to(dev) \
        if factored else None
    bucket = None
    if world > 1 and cfg["backward"] and not factored:
        if args.transport in ("auto", "nvls"):
            try:   # plain all-reduce of every leaf gradient, as this library's two-shot NVLS kernel
                from gaussiangrasper_b200.distributed import SymmetricBucket
                bucket = SymmetricBucket(P)
                transport = "gg_nvls_exchange kernel over symmetric memory, reduction only (%s)" % _exchange_mode(bucket.multicast, world, bucket.flat.numel() * 4)
            except Exception as e:
                if args.transport == "nvls":
                    raise
                bucket, transport = None, f"nccl (symmetric memory unavailable: {type(e).__name__}: {str(e)[:120]})"
        if bucket is None:
            bucket = GradientBucket(P)

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

    def render_part():
        """Forward (and backward) of every launch group of the step; returns (outputs of the last group, holder).
        Reads only tensors that live across steps (parameters, camera batches, cotangents): capturable."""
        for p in P.values():
            p.grad = None
        out = holder = None
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
        return out, holder

    def comm_part(out, holder):
        """The gradient exchange of a multi-GPU training step (never captured: NCCL / barriers of their own)."""
        if ex is not None:
            return ex.exchange(P["means"], chunks[0].positions, sh_degree, 4, holder, all_pos)["sh_coeffs"]
        if bucket is not None:
            if not (bucket is not None and len(chunks) == 1):
                bucket.pack({k: P[k].grad for k in names})
            bucket.all_reduce()
            return bucket.flat
        return out["image"]

    def step_plain():
        if args.path == "dropin":
            return step_dropin()
        return comm_part(*render_part())

    # SETUP_STEPS of initialisation (first-call capacity measurement, allocator growth, communicator set-up), then
    # exactly --warmup untimed warm-up steps
    if args.warmup < 3:
        raise SystemExit("bench.py: --warmup must be >= 3")
    for _ in range(SETUP_STEPS):
        step_plain()
    torch.cuda.synchronize()
    # The render part of the step is captured into a CUDA graph once (the binning never reads the host, all
    # buffers are capacity-sized): a step is then ONE graph launch + the exchange, not ~15 launches from Python.
    cap = None
    blend_names = ("gg_blend_fwd", "gg_blend_bwd")
    if args.graph and args.path == "fused":
        from gaussiangrasper_b200.graph import CapturedStep
        try:
            # two captures of the same step sharing one memory pool: `cap` is what the warm-up and the timed region
            # replay; `cap_prof` additionally carries an external event pair around each blend launch (four event
            # nodes: ~5 us of idle time each inside the step) and is replayed right behind the timed region only
            cap = CapturedStep(render_part, dev, warmup=2)
            cap_prof = CapturedStep(render_part, dev, warmup=0, pool=cap.pool(),
                                    capture_context=lambda: _lib.profile(only=blend_names, external=True))
        except Exception as e:   # capture refused: plain launches
            print(f"[bench] CUDA-graph capture failed ({type(e).__name__}: {str(e)[:200]}); plain launches", file=sys.stderr)
            cap = None
            torch.cuda.synchronize()

    def step():
        if cap is None:
            return step_plain()
        return comm_part(*cap.replay())

    n_warm = args.warmup
    for _ in range(n_warm):
        step()
    torch.cuda.synchronize()

    # --- timed region: exactly K steps, device-timed, max over ranks -------------------------
    # the clock sampler starts BEFORE the ranks line up (NVML initialisation takes tens of ms on rank 0 only: started
    # behind the barrier it would leave every other rank waiting inside its first timed step)
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    l0 = _lib.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    # inside the timed region only the two blend entry points carry an event pair (the roofline's
    # launch duration is measured live, here); the full per-entry-point breakdown is a separate pass.
    # With a captured step the pairs are external events recorded inside the graph (see below).
    with _lib.profile(only=blend_names) as prof_timed:
        e0.record()
        for _ in range(args.steps):
            step()
        e1.record()
        torch.cuda.synchronize()
        blend_calls = prof_timed.ms()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    launches = _lib.launch_count() - l0 + (cap.launches * args.steps if cap is not None else 0)
    clocks = sampler.stop() if sampler else None
    ms = e0.elapsed_time(e1)
    blend_source = "CUDA events around the C-ABI call, every step of the timed region"
    if cap is not None:
        cap.check()   # the intersection lists of the timed steps fitted the capacity frozen into the graph
        # duration of the blend launches INSIDE the captured step: the event pairs recorded in the graph hold the
        # times of the latest replay; 10 more replays of the same graph right behind the timed region, read one by one
        try:
            acc = {}
            for _ in range(10):
                cap_prof.replay()
                torch.cuda.synchronize()
                for k, v in cap_prof.context.ms().items():
                    acc.setdefault(k, []).extend(v)
            cap_prof.check()
            blend_calls = acc
            blend_source = ("external CUDA events recorded inside a second capture of the same step, replayed 10 times "
                            "directly behind the timed region (the timed graph itself carries no event nodes)")
        except Exception as e:
            blend_calls = {}
            blend_source = f"un-captured instrumented pass (events inside the graph unavailable: {type(e).__name__})"
    # per-entry-point breakdown: a separate, instrumented pass of plain launches (an event pair around every
    # C-ABI call perturbs the stream, so it stays out of the timed region above)
    prof_steps = min(args.steps, 10)
    with _lib.profile() as prof:
        for _ in range(prof_steps):
            step_plain()
        torch.cuda.synchronize()
        per_call = prof.ms()
    blend_steps = {k: (10 if cap is not None else args.steps) for k in blend_calls if blend_calls[k]}
    per_call.update({k: v for k, v in blend_calls.items() if v})   # the dominant kernels: from the timed path itself
    if args.diag and cap is not None:
        # diagnostic: device time and host enqueue time of the two halves of a step, alone and together
        def timed(fn, n=10):
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
            a_, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            h0 = time.perf_counter()
            a_.record()
            for _ in range(n):
                fn()
            b_.record()
            h1 = time.perf_counter()
            torch.cuda.synchronize()
            return a_.elapsed_time(b_) / n, (h1 - h0) * 1e3 / n
        outs = cap.outputs
        print(f"[diag rank {rank}] replay only      : device %.3f ms, host enqueue %.3f ms" % timed(lambda: cap.replay()), file=sys.stderr)
        print(f"[diag rank {rank}] exchange only    : device %.3f ms, host enqueue %.3f ms" % timed(lambda: comm_part(*outs)), file=sys.stderr)
        print(f"[diag rank {rank}] replay + exchange: device %.3f ms, host enqueue %.3f ms" % timed(step), file=sys.stderr)
        print(f"[diag rank {rank}] plain launches   : device %.3f ms, host enqueue %.3f ms" % timed(step_plain), file=sys.stderr)
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

    # --- e2e through captured graphs (single launch group): forward + loss in one graph, backward in a second one
    # that shares its memory pool, two such pairs ping-ponging on the two target buffers.  Per step the host does:
    # camera H2D (in place), replay F, enqueue the D2H of rgb + loss behind F, replay B, prefetch the next target.
    e2e_mode = "plain launches"
    if cap is not None and len(chunks) == 1:
        try:
            from gaussiangrasper_b200.graph import CapturedStep
            vb_e2e = ViewBatch.from_cameras(cams[:chunk], dev)
            pairs_fb = []
            for b_i in range(2):
                state = {}

                def fwd(b_i=b_i, state=state):
                    for p in P.values():
                        p.grad = None
                    holder = ex.holder() if ex is not None else None
                    with torch.set_grad_enabled(cfg["backward"]):
                        out = render_views(P["means"], P["log_scales"], P["quats"], P["opacity_logit"], P["sh_coeffs"],
                                           P["features"], vb_e2e, holder=holder)
                    if cfg["backward"]:
                        l, grad = pixel_loss(out["image"], target_dev[b_i], "l2")
                    else:
                        l, grad = out["alpha"].mean(), None
                    state.update(out=out, grad=grad, holder=holder)
                    return out["rgb"].detach(), l.detach().reshape(1)

                def bwd(state=state):
                    state["out"]["image"].backward(state["grad"])
                    return state["holder"]

                def both(fwd=fwd, bwd=bwd):        # warm-up of the pair outside any capture
                    r = fwd()
                    if cfg["backward"]:
                        bwd()
                    return r
                for _ in range(2):
                    both()
                torch.cuda.synchronize()
                F = CapturedStep(fwd, dev, warmup=0)
                B = CapturedStep(bwd, dev, warmup=0, pool=F.pool()) if cfg["backward"] else None
                pairs_fb.append((F, B))
            e2e_mode = "captured graphs (forward+loss | backward), double-buffered targets"
        except Exception as e:
            print(f"[bench] e2e graph capture failed ({type(e).__name__}: {str(e)[:300]}); plain launches", file=sys.stderr)
            pairs_fb = None
            torch.cuda.synchronize()
    else:
        pairs_fb = None

    rgb_hosts = [rgb_host, torch.empty_like(rgb_host).pin_memory()]
    loss_hosts = [loss_host, torch.empty_like(loss_host).pin_memory()]

    def e2e_step_graph(defer=False):
        """defer=False: the host waits for this step's loss before it returns.  defer=True: it returns the PREVIOUS
        step's loss (read from pinned memory once that step's copies have landed) and leaves this step in flight, the
        way a training loop that logs its loss keeps the device fed; e2e_drain() collects the last one."""
        main = torch.cuda.current_stream(dev)
        q = pf["q"]
        b_i = q & 1
        F, B = pairs_fb[b_i]
        if cfg["backward"]:
            if pf["ready"][b_i] is None:
                prefetch_target(q, chunk)                                # very first step only
            main.wait_event(pf["ready"][b_i])
            pf["ready"][b_i] = None
        if not pf.get("cams_staged"):
            vb_e2e.update_(cams[:chunk])                                 # H2D: cameras (pinned), in place; first step only
        if B is not None:
            # H2D: the NEXT step's supervision images go to the other buffer while this step computes (issued first:
            # the 29.5 MB copy takes longer than the forward, and the next forward + loss graph waits for it)
            prefetch_target(q + 1, chunk)
        rgb, l = F.replay()
        fwd_done = main.record_event()
        with torch.cuda.stream(d2h_stream):                              # D2H of the results runs under the backward
            d2h_stream.wait_event(fwd_done)
            rgb_hosts[b_i][:chunk].copy_(rgb, non_blocking=True)
            loss_hosts[b_i].copy_(l, non_blocking=True)
            results_done = d2h_stream.record_event()
        holder = None
        if B is not None:
            holder = B.replay()
            pf["consumed"][b_i] = main.record_event()                    # this target buffer may be refilled
        pf["q"] = q + 1
        if ex is not None:
            ex.exchange(P["means"], vb_e2e.positions, sh_degree, 4, holder, all_pos)
        elif bucket is not None:
            bucket.pack({k: P[k].grad for k in names})
            bucket.all_reduce()
        # H2D: the NEXT step's cameras, staged behind this step's kernels like its images (stream order keeps them
        # from overwriting the batch this step's backward still reads)
        vb_e2e.update_(cams[:chunk])
        pf["cams_staged"] = True
        if defer:
            prev, pf["inflight"] = pf.get("inflight"), (results_done, b_i)
            if prev is None:
                return None
            prev[0].synchronize()
            return float(loss_hosts[prev[1]][0])
        main.synchronize()
        results_done.synchronize()
        return float(loss_hosts[b_i][0])

    def e2e_drain():
        prev = pf.pop("inflight", None)
        v = None
        if prev is not None:
            prev[0].synchronize()
            v = float(loss_hosts[prev[1]][0])
        torch.cuda.current_stream(dev).synchronize()
        return v

    if pairs_fb is not None:
        plain_e2e_step = e2e_step
        pf.update(q=0, ready=[None, None], consumed=[None, None])
        pf.pop("vb", None)
        e2e_step = e2e_step_graph

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
    e2e_sync_val = e2e_val
    if pairs_fb is not None:
        # the same steps with one step in flight: the loss / rgb of step k are read back while step k+1 runs (every
        # step's inputs still go host->device and every step's results device->host inside the timed region; the last
        # step is drained before the clock stops)
        losses = []
        for _ in range(2):
            e2e_step(defer=True)
        e2e_drain()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            losses.append(e2e_step(defer=True))
        losses.append(e2e_drain())
        torch.cuda.synchronize()
        t_pipe = torch.tensor([time.perf_counter() - t0], device=dev)
        if world > 1:
            dist.all_reduce(t_pipe, op=dist.ReduceOp.MAX)
        assert sum(v is not None for v in losses) == args.steps and all(v == v for v in losses if v is not None)
        e2e_val = mpix_per_step * args.steps / float(t_pipe.item())
        e2e_mode += "; results of step k read back while step k+1 runs (value_sync_every_step: host waits for every step)"
    if args.trace and rank == 0 and world == 1:   # (single process only: the traced steps would leave the other ranks behind)
        _trace_timeline(e2e_step, step, args.trace)
    h2d = (chunk * H * W * CP * 4 if cfg["backward"] else 0) * (V // chunk) + V * 35 * 4
    d2h = V * H * W * 3 * 4 + 4

    if rank != 0:
        return None

    # --- roofline of the dominant kernel ------------------------------------------------------
    hbm_peak, peak_src = load_peaks()
    hbm_nominal = 8000.0   # B200 HBM3e nominal GB/s (the north star's "~8 TB/s"); the measured copy peak is the bar
    avg = {k: sum(v) / len(v) for k, v in per_call.items() if v}
    share = {k: sum(v) / blend_steps.get(k, prof_steps) for k, v in per_call.items()}
    # workload counters of the first chunk: pairs visited (K of SURVEY 8d) and pairs blended, intersections M
    hold = {}
    with torch.no_grad():
        render_views(*(P[k].detach() for k in names), views, holder=hold)
    pairs, pairs_contrib = _ops.blend_pair_stats(hold["binning"], hold["geo"], H, W)
    m_intersects = int(hold["binning"].num_intersects)
    n_tiles_total = int(hold["binning"].tile_ranges.shape[0])
    n_launch_groups = len(chunks)
    del hold
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
    # per launch (one launch group = `chunk` views; `pairs` was counted on the first chunk)
    flops_model = {"gg_blend_fwd": (12 + 2 * C), "gg_blend_bwd": (30 + 8 * C)}
    # static ncu evidence of the same kernels on config 1 (profiles/traffic.json, with its provenance)
    ncu = {}
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath) and args.config == 1 and V == 1 and args.feat < 0:
        with open(tpath) as f:
            ncu = json.load(f)

    def fp32_roof(kernel):
        """SURVEY 8d's fraction (every VISITED pair charged the full per-pair flops) next to what is executed:
        the same model over the pairs that are actually blended, and ncu's executed-instruction count."""
        if kernel not in avg:
            return None
        t = avg[kernel] * 1e-3
        fl_k = pairs * flops_model[kernel]
        ach = fl_k / t / 1e12
        r = {"kernel": kernel, "bound": "fp32", "achieved": ach, "peak": fma_tflops, "unit": "TFLOP/s",
             "frac": ach / fma_tflops, "peak_source": "FMA probe kernel timed in this run (nominal 74.4)",
             "frac_note": ("SURVEY 8d's definition: every VISITED pair charged the full per-pair flops; the kernels touch "
                           "only the blended pairs (frac_contributing_pairs) and part of the backward runs on the tensor "
                           "cores, so this fraction can exceed 1 -- executed_frac is the FP32 pipe's real load"),
             "algorithmic_flops": fl_k, "pairs": pairs, "pairs_contributing": pairs_contrib,
             "frac_contributing_pairs": pairs_contrib * flops_model[kernel] / t / 1e12 / fma_tflops,
             "avg_launch_ms": avg[kernel]}
        e = ncu.get(kernel, {})
        r["traffic"] = e.get("dram_bytes_per_launch")
        if e:
            r["traffic_source"] = "ncu capture of this kernel on this workload: " + e.get("source", "profiles/traffic.json")
            if e.get("executed_fp32_flops_per_launch"):
                r["executed_frac"] = e["executed_fp32_flops_per_launch"] / t / 1e12 / fma_tflops
                r["executed_frac_source"] = ("(2 x FFMA + FADD + FMUL thread instructions, ncu smsp__sass_thread_inst_executed_"
                                             "op_*_pred_on.sum) / this run's launch time / FMA probe")
            if e.get("pipe_fma_cycles_active_pct") is not None:
                r["ncu_pipe_fma_cycles_active_pct"] = e["pipe_fma_cycles_active_pct"]
            if e.get("tensor_flops_per_launch"):
                r["executed_tensor_tflops"] = e["tensor_flops_per_launch"] / t / 1e12
                r["executed_tensor_note"] = "mma.sync m16n8k8 TF32, 3 MMAs per product (3xTF32 split): phase B of the backward"
        return r

    roof = fp32_roof(dom) if dom in flops_model else None
    if roof is None and dom is not None:
        roof = {"kernel": dom, "bound": "hbm", "achieved": None, "peak": hbm_peak, "unit": "GB/s", "frac": None,
                "traffic": None, "peak_source": peak_src, "avg_launch_ms": avg[dom]}
    roofline_other = {k: fp32_roof(k) for k in flops_model if k != dom and k in avg}
    # HBM-bound stages: algorithmic bytes per STEP against the measured copy bandwidth and
    # the nominal 8 TB/s.  Per launch group of v views over n Gaussians:
    #   prepare fwd : parameters once (44 geometry + 300 SH + 4D features; +40 re-read by the channel kernel when
    #                 v > 1) + per view geo 32 + depth/radius/count 12 + channel row 4CP (+8 re-read when v > 1)
    #   prepare bwd : per view geo 32 + chan 4CP + radius 4 + v_geo 32 + v_chan 4CP, parameters 44 once, gradients
    #                 44 + 4D + (300, or 12 per view when the SH gradient is left as its factor)
    #   binning     : SURVEY 8d: scan 8 + key emission 20 per (view, Gaussian); 12 + 24 (sort, compulsory) + 8
    #                 (ranges) per intersection; 8 per tile
    def prep_bytes(v):
        return n * (344 + 4 * D + (40 if v > 1 else 0)) + v * n * (44 + 4 * CP + (8 if v > 1 else 0))

    def prep_bwd_bytes(v):
        sh_out = 12 * v if ex is not None else 300
        return v * n * (68 + 8 * CP) + n * 44 + n * (44 + 4 * D + sh_out)

    groups = [c.n_views for c in chunks]
    m_per_view = m_intersects / max(1, chunks[0].n_views)
    bytes_model = {"gg_prepare_views": sum(prep_bytes(v) for v in groups),
                   "gg_prepare_views_bwd": sum(prep_bwd_bytes(v) for v in groups),
                   "gg_bin_tiles": sum(28 * n * v + 44 * m_per_view * v + 8 * (n_tiles_total / chunks[0].n_views) * v
                                       for v in groups)}
    hbm_stages = {}
    for k, bts in bytes_model.items():
        if k in share and share[k] > 0:
            gbs = bts / (share[k] * 1e-3) / 1e9   # all launches of the stage in one step
            hbm_stages[k] = {"GB/s": gbs, "frac": gbs / hbm_peak, "frac_of_nominal_8000": gbs / hbm_nominal,
                             "ms_per_step": share[k], "bytes_per_step": int(bts)}
    if "gg_bin_tiles" in hbm_stages:
        hbm_stages["gg_bin_tiles"]["intersections_first_group"] = m_intersects

    # --- CPU baseline (rank 0, N=1 only): the reference maths on the host cores ---------------
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        cpu = cpu_baseline_leg({k: sc[k] for k in names}, cams[0], D, n, W, H, cfg["backward"])

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": n_warm,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong" if strong else "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": config,
        "e2e": {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "mode": e2e_mode,
                "value_sync_every_step": e2e_sync_val},
        "gpu_launches": launches,
        "clocks": clocks,
        "roofline": roof,
        "roofline_other_kernels": roofline_other,
        "hbm_stages": hbm_stages,
        "hbm_peak_GBs": {"measured": hbm_peak, "source": peak_src, "nominal": hbm_nominal},
        "stage_ms_per_step": share,
        "exchange_transport": transport,
        "cuda_graph": (f"render part of the step captured once ({cap.launches} kernel launches of this library per replay)"
                       if cap is not None else "off"),
        "blend_launch_timing": blend_source,
        "fma_probe_tflops": fma_tflops,
        "cpu_baseline": cpu,
    }
    return line


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
    ap.add_argument("--transport", default="auto", choices=["auto", "nvls", "nccl"],
                    help="factored exchange: this library's symmetric-memory kernel (nvls; auto = when available) or NCCL")
    ap.add_argument("--no-graph", dest="graph", action="store_false",
                    help="enqueue every launch from Python instead of replaying the captured step")
    ap.add_argument("--diag", action="store_true", help="diagnostic timings of the step's halves on stderr")
    ap.add_argument("--no-config3", action="store_true", help="N > 1: do not add the configs[3] measurement to the line")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--trace", default="", help="diagnostic: write a GPU timeline (kernels + idle gaps) of one e2e and one "
                                                "device-timed step to this file (torch.profiler; not part of the measurement)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
        return
    import copy
    import torch.distributed as dist
    line = run_ours(args)
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world > 1 and args.config == 1 and not args.no_config3:
        # BASELINE configs[3] (1 M Gaussians, 8 views per GPU, fwd+bwd + all-reduce) measured in the same run
        a3 = copy.copy(args)
        a3.config, a3.steps, a3.warmup, a3.no_cpu_baseline, a3.trace, a3.diag, a3.views, a3.feat = 3, 10, 3, True, "", False, 0, -1
        try:
            l3 = run_ours(a3, sub=True)
            if line is not None and l3 is not None:
                line["baseline_config3"] = {k: l3[k] for k in ("value", "unit", "ms_per_step", "steps", "warmup", "config", "e2e",
                                                               "exchange_transport", "cuda_graph", "stage_ms_per_step", "scaling")}
        except Exception as e:
            if line is not None:
                line["baseline_config3"] = {"error": f"{type(e).__name__}: {str(e)[:300]}"}
    if line is not None:
        print(json.dumps(line), flush=True)
    if world > 1 and dist.is_initialized():
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
