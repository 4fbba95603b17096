"""The two per-Gaussian streaming steps directly behind the rasterizer's backward in a training
iteration ("next rows"): a fused Adam over the flat all-reduced gradient buffer and the
densification statistics of GaussianSplattingModel.after_train."""
from __future__ import annotations

import ctypes as C
from typing import Dict, Optional

import torch

from . import _lib
from .distributed import GradientBucket

# learning rates of the reference's optimizer groups (nerfstudio/configs/method_configs.py:618-664)
REFERENCE_LRS = dict(means=1.6e-4, log_scales=5e-3, quats=1e-3, opacity_logit=5e-2, sh_coeffs=5e-4, features=5e-4)


# gradient accumulation of the reference trainer (method_configs.py:611, engine/trainer.py:466-481): the
# "xyz", "color" (= SH coefficients) and "feature" groups step every 10th iteration on the sum of 10 gradients
REFERENCE_ACCUMULATION = dict(means=10, sh_coeffs=10, features=10)
# ExponentialDecayScheduler of each group (method_configs.py:621-650): (lr_final, max_steps); opacity / rotation
# have no scheduler
REFERENCE_SCHEDULES = dict(means=(1.6e-6, 30000), sh_coeffs=(1e-4, 30000), features=(1e-4, 30000),
                           log_scales=(1e-3, 30000))
_MODE = dict(step=0, acc_first=1, acc=2, acc_step=3, skip=4)   # GG_ADAM_* of include/gg_b200.h


def sh_degrees_to_use(step: int, sh_degree: int = 4, sh_degree_interval: int = 1000) -> int:
    """How many SH bands the model evaluates at training step `step` (gaussian_splatting.py:729: one more band every
    sh_degree_interval steps) -- the `degrees_to_use` argument of render_views / SphericalHarmonics."""
    return min(step // sh_degree_interval, sh_degree)


def exponential_decay_lr(lr_init: float, lr_final: float, max_steps: int, step: int) -> float:
    """engine/schedulers.py:122-138 without warm-up: log-linear interpolation lr_init -> lr_final over max_steps."""
    import math
    t = min(max(step / max_steps, 0.0), 1.0)
    return math.exp(math.log(lr_init) * (1.0 - t) + math.log(lr_final) * t)


class FusedAdam:
    """torch.optim.Adam semantics (betas (0.9, 0.999), eps 1e-15 as the reference configures it, no weight
    decay) for all Gaussian parameters in ONE kernel over the flat gradient bucket.  Every parameter keeps its
    own update count, as the reference's per-group optimizers do.

    step()            one plain update of every parameter from the bucket.
    train_step(it)    what the reference trainer does at iteration `it` (0-based): per-group gradient
                      accumulation (`accumulation`, name -> k) and the per-group exponential learning-rate decay
                      (`schedules`, name -> (lr_final, max_steps)), still one kernel."""

    def __init__(self, params: Dict[str, torch.Tensor], bucket: Optional[GradientBucket] = None,
                 lrs: Optional[Dict[str, float]] = None, betas=(0.9, 0.999), eps: float = 1e-15,
                 accumulation: Optional[Dict[str, int]] = None, schedules: Optional[Dict[str, tuple]] = None):
        self.params = params
        self.bucket = bucket if bucket is not None else GradientBucket(params)
        self.lrs = dict(REFERENCE_LRS if lrs is None else lrs)
        self.lr_init = dict(self.lrs)
        self.betas, self.eps = betas, eps
        self.steps = {k: 0 for k in self.bucket.names}     # per-parameter update count (torch.optim.Adam's `step`)
        self.accumulation = dict(accumulation or {})
        self.schedules = dict(schedules or {})
        self.exp_avg = torch.zeros_like(self.bucket.flat)
        self.exp_avg_sq = torch.zeros_like(self.bucket.flat)
        self.accum = None
        for k in self.bucket.names:
            p = params[k]
            if not (p.is_cuda and p.is_contiguous() and p.dtype == torch.float32):
                raise _lib.GGError(f"parameter {k} must be a contiguous fp32 CUDA tensor")

    @classmethod
    def reference(cls, params, bucket=None):
        """Configured like the reference's `gaussian-splatting` method (method_configs.py:603-664)."""
        return cls(params, bucket, accumulation=REFERENCE_ACCUMULATION, schedules=REFERENCE_SCHEDULES)

    @property
    def t(self) -> int:
        return max(self.steps.values()) if self.steps else 0

    def _launch(self, modes: Dict[str, str]) -> None:
        b = self.bucket
        names = b.names
        for k in names:   # a bucket left over from before a refinement must not be applied to the new parameters
            if tuple(self.params[k].shape) != b.shapes[k]:
                raise _lib.GGError(f"FusedAdam: parameter {k} is {tuple(self.params[k].shape)} but the gradient "
                                   f"bucket was built for {b.shapes[k]}; call rebuild() after refine_gaussians")
        if any(m in ("acc_first", "acc", "acc_step") for m in modes.values()) and self.accum is None:
            self.accum = torch.zeros_like(b.flat)
        for k in names:
            if modes[k] in ("step", "acc_step"):
                self.steps[k] += 1
        n = len(names)
        ptrs = (C.c_void_p * n)(*[self.params[k].data_ptr() for k in names])
        offs = (C.c_longlong * n)(*[b.offsets[k] for k in names])
        cnts = (C.c_longlong * n)(*[b.sizes[k] for k in names])
        lr = (C.c_float * n)(*[float(self.lrs[k]) for k in names])
        st = (C.c_int * n)(*[max(1, self.steps[k]) for k in names])
        md = (C.c_int * n)(*[_MODE[modes[k]] for k in names])
        dev = b.flat.device
        with _lib.device_guard(dev):
            _lib.call("gg_adam_step", n, ptrs, offs, cnts, lr, st, md, b.flat.data_ptr(), _lib.ptr(self.accum),
                      self.exp_avg.data_ptr(), self.exp_avg_sq.data_ptr(), float(self.betas[0]), float(self.betas[1]),
                      float(self.eps), _lib.stream_ptr(dev))

    @torch.no_grad()
    def step(self, lrs: Optional[Dict[str, float]] = None) -> None:
        """Apply one update from the gradients currently in the bucket (already summed over ranks)."""
        if lrs:
            self.lrs.update(lrs)
        self._launch({k: "step" for k in self.bucket.names})

    def plan(self, iteration: int) -> Dict[str, str]:
        """What each group does at trainer iteration `iteration` (0-based), as engine/trainer.py:466-481:
        zero_grad at it % k == 0, optimizer step at it % k == k - 1, gradients summed in between."""
        modes = {}
        for k in self.bucket.names:
            a = int(self.accumulation.get(k, 1))
            if a <= 1:
                modes[k] = "step"
            elif iteration % a == a - 1:
                modes[k] = "acc_step"
            elif iteration % a == 0:
                modes[k] = "acc_first"
            else:
                modes[k] = "acc"
        return modes

    @torch.no_grad()
    def train_step(self, iteration: int) -> Dict[str, str]:
        """One trainer iteration: learning rates of `iteration` (the LambdaLR schedulers have been stepped once
        per past iteration, engine/trainer.py:497), accumulation per group, one kernel."""
        for k, (lr_final, max_steps) in self.schedules.items():
            if k in self.lr_init:
                self.lrs[k] = exponential_decay_lr(self.lr_init[k], lr_final, max_steps, iteration)
        modes = self.plan(iteration)
        self._launch(modes)
        return modes

    def moments(self) -> Dict[str, tuple]:
        """name -> (exp_avg, exp_avg_sq) views shaped like the parameters (what refine_gaussians takes)."""
        b = self.bucket
        out = {}
        for k in b.names:
            o, n = b.offsets[k], b.sizes[k]
            out[k] = (self.exp_avg[o:o + n].view(b.shapes[k]), self.exp_avg_sq[o:o + n].view(b.shapes[k]))
        return out

    @torch.no_grad()
    def rebuild(self, params: Dict[str, torch.Tensor], moments: Optional[Dict[str, tuple]] = None) -> None:
        """Adopt a refined parameter set (refine_gaussians): new bucket, the given moments (zeros if None).
        The update counts are kept, as the reference's surgery keeps each optimizer's `step`; gradients under
        accumulation are dropped, as the reference's new Parameters start without a .grad."""
        self.params = params
        self.bucket = GradientBucket(params, self.bucket.group)
        self.exp_avg = torch.zeros_like(self.bucket.flat)
        self.exp_avg_sq = torch.zeros_like(self.bucket.flat)
        self.accum = None
        if moments is not None:
            for k, (ea, es) in self.moments().items():
                ea.copy_(moments[k][0].reshape(ea.shape))
                es.copy_(moments[k][1].reshape(es.shape))

    @torch.no_grad()
    def reset_opacity(self, cull_alpha_thresh: float = 0.1) -> None:
        """gaussian_splatting.py:458-464: opacities = logit(0.8 * cull_alpha_thresh), opacity moments cleared."""
        value = torch.logit(torch.tensor(cull_alpha_thresh * 0.8)).item()
        self.params["opacity_logit"].fill_(value)
        ea, es = self.moments()["opacity_logit"]
        ea.zero_()
        es.zero_()

    # ---- checkpoint (engine/trainer.py:427-456 saves {"step", "pipeline", "optimizers", "schedulers"}) ----------
    def state_dict(self) -> dict:
        """Per parameter: torch.optim.Adam's state entries {step, exp_avg, exp_avg_sq} + the current lr."""
        m = self.moments()
        return {k: dict(step=int(self.steps[k]), lr=float(self.lrs[k]), lr_init=float(self.lr_init[k]),
                        exp_avg=m[k][0].detach().clone(), exp_avg_sq=m[k][1].detach().clone())
                for k in self.bucket.names}

    @torch.no_grad()
    def load_state_dict(self, state: dict) -> None:
        m = self.moments()
        for k in self.bucket.names:
            e = state[k]
            if tuple(e["exp_avg"].shape) != tuple(m[k][0].shape):
                raise _lib.GGError(f"optimizer state of {k} is {tuple(e['exp_avg'].shape)}, parameter is {tuple(m[k][0].shape)}")
            m[k][0].copy_(e["exp_avg"])
            m[k][1].copy_(e["exp_avg_sq"])
            self.steps[k] = int(e["step"])
            self.lrs[k] = float(e.get("lr", self.lrs[k]))
            self.lr_init[k] = float(e.get("lr_init", self.lr_init[k]))
        self.accum = None


class DensifyStats:
    """xys_grad_norm / vis_counts / max_2Dsize of gaussian_splatting.py:373-393, fed from the holder of
    render_views after backward (v_geo columns 0..1 are d loss / d xys).

    View-sharded training: every rank accumulates the statistics of ITS views; `all_reduce()` (called by
    refine_gaussians) combines them so that every rank takes identical decisions -- SUM of the gradient norms and
    visibility counts, MAX of the screen sizes.  The reference's first step initialises the counts to one for every
    Gaussian (:381); only rank 0 does that here, so the reduced statistics equal a single process that saw rank 0's
    first view first and then every other view."""

    def __init__(self, n: int, device, group=None):
        import torch.distributed as dist
        self.n = n
        self.group = group
        self.xys_grad_norm = torch.zeros(n, dtype=torch.float32, device=device)
        self.vis_counts = torch.zeros(n, dtype=torch.float32, device=device)
        self.max_2Dsize = torch.zeros(n, dtype=torch.float32, device=device)
        on = dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1
        self.first = (not on) or dist.get_rank(group) == 0
        self.reduced = False

    @torch.no_grad()
    def update(self, v_geo: torch.Tensor, radii: torch.Tensor, img_height: int, img_width: int) -> None:
        if self.reduced:
            raise _lib.GGError("DensifyStats.update after all_reduce(): the reduced statistics belong to the refinement "
                               "that follows; start a new DensifyStats for the next interval")
        n_views = radii.numel() // self.n
        dev = v_geo.device
        radii = radii.contiguous()
        with _lib.device_guard(dev):
            _lib.call("gg_densify_stats", self.n, n_views, v_geo.data_ptr(), radii.data_ptr(), int(img_height),
                      int(img_width), 1 if self.first else 0, self.xys_grad_norm.data_ptr(), self.vis_counts.data_ptr(),
                      self.max_2Dsize.data_ptr(), _lib.stream_ptr(dev))
        self.first = False

    @torch.no_grad()
    def all_reduce(self) -> "DensifyStats":
        """Combine the ranks' statistics in place (idempotent; a no-op without a process group)."""
        import torch.distributed as dist
        if self.reduced or not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(self.group) == 1:
            self.reduced = True
            return self
        both = torch.stack([self.xys_grad_norm, self.vis_counts])
        dist.all_reduce(both, op=dist.ReduceOp.SUM, group=self.group)
        self.xys_grad_norm.copy_(both[0])
        self.vis_counts.copy_(both[1])
        dist.all_reduce(self.max_2Dsize, op=dist.ReduceOp.MAX, group=self.group)
        self.reduced = True
        return self


class RefineConfig(C.Structure):
    """Mirror of gg_refine_config (include/gg_b200.h)."""
    _fields_ = [("max_dim", C.c_float), ("densify_grad_thresh", C.c_float), ("densify_size_thresh", C.c_float),
                ("split_screen_size", C.c_float), ("cull_alpha_thresh", C.c_float), ("cull_scale_thresh", C.c_float),
                ("cull_screen_size", C.c_float), ("do_densify", C.c_int), ("split_by_screen", C.c_int),
                ("do_cull", C.c_int), ("cull_by_scale", C.c_int), ("cull_by_screen", C.c_int)]


# GaussianSplattingModelConfig defaults (nerfstudio/models/gaussian_splatting.py:113-170)
REFERENCE_REFINE = dict(warmup_length=500, refine_every=100, cull_alpha_thresh=0.1, cull_scale_thresh=0.5,
                        reset_alpha_every=30, densify_grad_thresh=0.0002, densify_size_thresh=0.01, n_split_samples=2,
                        cull_screen_size=0.15, split_screen_size=0.05, stop_screen_size_at=4000, stop_split_at=15000)


def refine_schedule(step: int, num_train_data: int, max_dim: int, cfg: Optional[dict] = None) -> Optional[dict]:
    """Which rules refinement_after applies at `step` (gaussian_splatting.py:396-410, 456, 471-475); None when
    nothing happens.  The returned dict holds the fields of gg_refine_config plus `reset_opacity`."""
    c = dict(REFERENCE_REFINE)
    c.update(cfg or {})
    if step < c["warmup_length"]:
        return None
    reset_interval = c["reset_alpha_every"] * c["refine_every"]
    past_reset = step % reset_interval > num_train_data + c["refine_every"]
    return dict(max_dim=float(max_dim), densify_grad_thresh=c["densify_grad_thresh"],
                densify_size_thresh=c["densify_size_thresh"], split_screen_size=c["split_screen_size"],
                cull_alpha_thresh=c["cull_alpha_thresh"], cull_scale_thresh=c["cull_scale_thresh"],
                cull_screen_size=c["cull_screen_size"],
                do_densify=int(step < c["stop_split_at"] and past_reset),
                split_by_screen=int(step < c["stop_screen_size_at"]),
                do_cull=int(past_reset),
                cull_by_scale=int(step > c["refine_every"] * c["reset_alpha_every"]),
                cull_by_screen=int(step < c["stop_screen_size_at"]),
                reset_opacity=bool(step % reset_interval == c["refine_every"]))


_ROW_KIND = dict(means=2, log_scales=3)  # GG_REFINE_MEANS / GG_REFINE_LOG_SCALES; everything else copies

_refine_cache = {}


def _refine_buffers(dev, need: int):
    """Grow-only plan workspace (256-byte aligned) and the pinned totals record, per device: a refinement every
    100 iterations should not allocate device memory or pin host memory each time."""
    key = (dev.type, dev.index)
    ent = _refine_cache.get(key)
    if ent is None or ent[0].numel() < need + 256:
        raw = torch.empty(int(need * 1.25) + 512, dtype=torch.uint8, device=dev)
        ent = _refine_cache[key] = (raw, torch.zeros(4, dtype=torch.int32).pin_memory())
    raw, totals = ent
    return raw[(-raw.data_ptr()) % 256:], totals


def philox_normals(count: int, n_samples: int, seed: int, step: int, device, parents: Optional[torch.Tensor] = None):
    """[n_samples, count, 3] standard normals of the in-kernel generator (gg_philox_normals): row (s, r) is the
    sample s of parent row parents[r] (or r)."""
    out = torch.empty((n_samples, count, 3), dtype=torch.float32, device=device)
    if parents is not None:
        parents = parents.to(torch.int32).contiguous()
    with _lib.device_guard(out.device):
        _lib.call("gg_philox_normals", int(count), _lib.ptr(parents), int(n_samples), int(seed) & 0xFFFFFFFFFFFFFFFF,
                  int(step) & 0xFFFFFFFF, out.data_ptr(), _lib.stream_ptr(out.device))
    return out


@torch.no_grad()
def refine_gaussians(params: Dict[str, torch.Tensor], moments: Optional[Dict[str, tuple]], stats: Optional[DensifyStats],
                     rules: dict, n_split_samples: int = 2, samples_fn=None, seed: Optional[int] = None, step: int = 0):
    """Densify + cull on the device, in the reference's output order, parameters and Adam moments together.

    params: name -> [N, ...] fp32 CUDA tensors; moments: name -> (exp_avg, exp_avg_sq) like params, or None.
    rules: fields of gg_refine_config (see refine_schedule).
    The split children's offsets: with `seed` given they come from the counter-based Philox generator inside the
    kernel, keyed on (seed, step, parent row, sample index) -- identical on every rank of a view-sharded run, no
    broadcast, nothing allocated (REQUIRED when a process group with more than one rank is initialised).
    Otherwise samples_fn(k) -> [k,3] standard-normal CUDA tensor (default torch.randn, like the reference :491)
    is called with n_split_samples * (number of split parents).
    `stats` are all-reduced over their process group first (DensifyStats.all_reduce), so the ranks decide alike.
    Returns (new params, new moments, info).  One host read (the four totals), as the reference has several."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1 and seed is None and samples_fn is None:
        raise _lib.GGError("refine_gaussians under a process group needs `seed` (counter-based split samples): "
                           "per-rank torch.randn would make the replicas diverge")
    if stats is not None:
        stats.all_reduce()
    names = [k for k in ("means", "log_scales", "quats", "opacity_logit", "sh_coeffs", "features") if k in params]
    for k in ("means", "log_scales", "quats", "opacity_logit"):
        if k not in params:
            raise _lib.GGError(f"refine_gaussians needs parameter {k}")
    dev = _lib.require_cuda(*[params[k] for k in names])
    P = {k: _lib.f32c(params[k].detach()) for k in names}
    n = P["means"].shape[0]
    cfg = RefineConfig(**{f: rules[f] for f, _ in RefineConfig._fields_})
    lib = _lib.load()
    ws, totals = _refine_buffers(dev, int(lib.gg_refine_workspace_bytes(n)))
    need_stats = bool(rules["do_densify"])
    if need_stats and stats is None:
        raise _lib.GGError("densification needs the DensifyStats of the last refine interval")
    m2d = stats.max_2Dsize if stats is not None else None
    with _lib.device_guard(dev):
        st = _lib.stream_ptr(dev)
        _lib.call("gg_refine_plan", n, _lib.ptr(stats.xys_grad_norm) if stats is not None else None,
                  _lib.ptr(stats.vis_counts) if stats is not None else None, _lib.ptr(m2d), _lib.ptr(P["log_scales"]),
                  _lib.ptr(P["opacity_logit"]), C.byref(cfg), ws.data_ptr(), ws.numel(), totals.data_ptr(), st)
        torch.cuda.current_stream(dev).synchronize()
        n_keep, n_sk, n_dk, n_sa = (int(x) for x in totals)
        n_out = n_keep + n_split_samples * n_sk + n_dk
        samples = None
        if n_sa > 0 and (seed is None or samples_fn is not None):
            samples = (samples_fn or (lambda k: torch.randn((k, 3), device=dev)))(n_split_samples * n_sa)
            samples = _lib.f32c(samples.to(dev))
        new_p = {k: torch.empty((n_out,) + tuple(P[k].shape[1:]), dtype=torch.float32, device=dev) for k in names}
        new_m = None
        src, dst, rows, kinds = [], [], [], []
        for k in names:
            src.append(P[k]); dst.append(new_p[k]); rows.append(max(1, P[k][0].numel()) if n else 1)
            kinds.append(_ROW_KIND.get(k, 0))
        if moments is not None:
            new_m = {}
            for k in names:
                ea, es = (_lib.f32c(t) for t in moments[k])
                na, ns = torch.empty_like(new_p[k]), torch.empty_like(new_p[k])
                new_m[k] = (na, ns)
                for s_, d_ in ((ea, na), (es, ns)):
                    src.append(s_); dst.append(d_); rows.append(rows[names.index(k)]); kinds.append(1)
        if n_out > 0:
            na_ = len(src)
            scratch = torch.empty(n_out * 9 + 512, dtype=torch.uint8, device=dev)
            _lib.call("gg_refine_apply", n, int(n_split_samples), (C.c_int32 * 4)(n_keep, n_sk, n_dk, n_sa), ws.data_ptr(),
                      na_, (C.c_void_p * na_)(*[t.data_ptr() for t in src]), (C.c_void_p * na_)(*[t.data_ptr() for t in dst]),
                      (C.c_int * na_)(*rows), (C.c_int * na_)(*kinds), _lib.ptr(P["means"]), _lib.ptr(P["log_scales"]),
                      _lib.ptr(P["quats"]), _lib.ptr(samples), int(seed or 0) & 0xFFFFFFFFFFFFFFFF, int(step) & 0xFFFFFFFF,
                      scratch.data_ptr(), scratch.numel(), st)
    info = dict(n_in=n, n_out=n_out, n_kept=n_keep, n_split=n_sa, n_split_kept=n_sk, n_dup_kept=n_dk)
    return new_p, new_m, info


@torch.no_grad()
def knn_scale_init(means: torch.Tensor):
    """The model's scale initialisation (gaussian_splatting.py:259-263): log of the mean distance to the three
    nearest other points, on all three axes -- sklearn's NearestNeighbors on the host in the reference, one exact
    brute-force kernel here.  Returns (log_scales [N,3], distances [N,3] ascending)."""
    dev = _lib.require_cuda(means)
    m = _lib.f32c(means.detach())
    n = m.shape[0]
    dist = torch.empty((n, 3), dtype=torch.float32, device=dev)
    ls = torch.empty((n, 3), dtype=torch.float32, device=dev)
    with _lib.device_guard(dev):
        _lib.call("gg_knn3_scales", n, m.data_ptr(), dist.data_ptr(), ls.data_ptr(), _lib.stream_ptr(dev))
    return ls, dist


_loss_ws = {}


@torch.no_grad()
def pixel_loss(pred: torch.Tensor, target: torch.Tensor, kind: str = "l1", mask: Optional[torch.Tensor] = None,
               weight: float = 1.0, mean_over: str = "valid"):
    """weight * mean(|pred - target|) (kind "l1") or weight * mean((pred - target)^2) ("l2") and its gradient
    w.r.t. pred, in one kernel.  mask: optional [pixels] bool/uint8, False = ignored pixel; the mean then runs
    over the valid pixels (mean_over="valid": `abs(gt[valid] - rgb[valid]).mean()`, gaussian_splatting.py:882)
    or over all pixels with the ignored ones contributing zero (mean_over="all").
    Returns (loss [1], grad like pred); continue with `pred.backward(grad)`."""
    dev = _lib.require_cuda(pred, target, mask)
    p, t = _lib.f32c(pred.detach()), _lib.f32c(target.detach())
    if p.shape != t.shape:
        raise ValueError(f"pred {tuple(p.shape)} and target {tuple(t.shape)} differ")
    channels = int(p.shape[-1])
    m = None
    if mask is not None:
        m = mask.to(torch.uint8).contiguous()
        if m.numel() * channels != p.numel():
            raise ValueError("mask must have one entry per pixel")
    key = (dev.type, dev.index)
    ws = _loss_ws.get(key)
    if ws is None:
        ws = _loss_ws[key] = torch.zeros(int(_lib.load().gg_pixel_loss_workspace_bytes()), dtype=torch.uint8, device=dev)
    grad = torch.empty_like(p)
    loss = torch.empty(1, dtype=torch.float32, device=dev)
    if mean_over not in ("valid", "all"):
        raise ValueError("mean_over is 'valid' or 'all'")
    count = m.sum(dtype=torch.int32).reshape(1) if (m is not None and mean_over == "valid") else None
    with _lib.device_guard(dev):
        _lib.call("gg_pixel_loss", p.numel(), channels, p.data_ptr(), t.data_ptr(), _lib.ptr(m), _lib.ptr(count),
                  {"l1": 1, "l2": 2}[kind], float(weight), grad.data_ptr(), loss.data_ptr(), ws.data_ptr(), ws.numel(),
                  _lib.stream_ptr(dev))
    return loss, grad


_ssim_ws = {}


@torch.no_grad()
def ssim_loss(pred: torch.Tensor, target: torch.Tensor, channels: int = 3, weight: float = 1.0,
              grad: Optional[torch.Tensor] = None, loss: Optional[torch.Tensor] = None):
    """weight * (1 - SSIM(pred[..., :channels], target[..., :channels])) with the reference's SSIM settings
    (gaussian_splatting.py:284, :885) and its gradient w.r.t. pred.  pred [V,H,W,Cp] or [H,W,Cp], target
    [...,Ct] (channel-last, contiguous).  If `grad` (like pred) / `loss` ([1]) are given, the results are ADDED
    to them -- e.g. the outputs of pixel_loss, to form main_loss = (1-l) L1 + l (1-SSIM) (:931).
    Returns (loss [1], grad like pred)."""
    dev = _lib.require_cuda(pred, target, grad, loss)
    p, t = _lib.f32c(pred.detach()), _lib.f32c(target.detach())
    if p.dim() == 3:
        p, t = p[None], t[None]
    if p.dim() != 4 or t.dim() != 4 or p.shape[:3] != t.shape[:3] or min(p.shape[3], t.shape[3]) < channels:
        raise ValueError(f"pred {tuple(pred.shape)} / target {tuple(target.shape)} must be [V,H,W,C>=channels]")
    V, H, W, ps = p.shape
    if H < 11 or W < 11:
        raise ValueError("SSIM needs images of at least 11x11 pixels")
    accumulate = grad is not None
    if (grad is None) != (loss is None):
        raise ValueError("pass both grad and loss to accumulate, or neither")
    if accumulate:
        if grad.numel() != p.numel() or not grad.is_contiguous() or grad.dtype != torch.float32:
            raise ValueError("grad must be a contiguous fp32 tensor shaped like pred")
    else:
        grad = torch.zeros_like(p)   # channels beyond `channels` get no SSIM gradient
        loss = torch.empty(1, dtype=torch.float32, device=dev)
    need = int(_lib.load().gg_ssim_workspace_bytes(V, H, W, channels))
    key = (dev.type, dev.index)
    ws = _ssim_ws.get(key)
    if ws is None or ws.numel() < need:
        ws = _ssim_ws[key] = torch.empty(need, dtype=torch.uint8, device=dev)
    with _lib.device_guard(dev):
        _lib.call("gg_ssim_loss", V, H, W, int(channels), p.data_ptr(), int(ps), t.data_ptr(), int(t.shape[3]),
                  float(weight), grad.data_ptr(), int(ps), 1 if accumulate else 0, loss.data_ptr(), ws.data_ptr(),
                  ws.numel(), _lib.stream_ptr(dev))
    return loss, grad.view(pred.shape)
