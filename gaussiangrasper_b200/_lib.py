"""ctypes binding of libgg_b200.so (the C ABI declared in include/gg_b200.h).

There is no CPU fallback: if the shared library is missing and cannot be built, or a call
fails, this module raises.  PyTorch is used only to own device memory and name the stream.
"""
from __future__ import annotations

import ctypes as C
import os
import threading

import torch

from . import build as _build

_lock = threading.Lock()
_lib = None
ABI_VERSION = 200   # == GG_ABI_VERSION in include/gg_b200.h; bumped whenever a signature changes

_i, _ll, _f, _p, _sz = C.c_int, C.c_longlong, C.c_float, C.c_void_p, C.c_size_t

_SIGNATURES = {
    "gg_version": (C.c_int, []),
    "gg_last_error_string": (C.c_char_p, []),
    "gg_launch_count": (C.c_ulonglong, []),
    "gg_check_device": (C.c_int, []),
    "gg_bench_fma": (C.c_int, [_i, _i, _p, _p, _p]),
    "gg_project_fwd": (C.c_int, [_i, _p, _p, _f, _p, _p, _p, _f, _f, _f, _f, _i, _i, _i, _i, _f, _p, _p, _p, _p, _p, _p, _p]),
    "gg_project_bwd": (C.c_int, [_i, _p, _p, _f, _p, _p, _p, _f, _f, _f, _f, _i, _i, _p, _p, _p, _p, _p, _p, _p, _p, _p]),
    "gg_project_fwd_views": (C.c_int, [_i, _i, _p, _p, _f, _p, _p, _p, _p, _f, _f, _f, _f, _i, _i, _i, _i, _f, _p, _p, _p, _p, _p, _p, _p]),
    "gg_project_bwd_views": (C.c_int, [_i, _i, _p, _p, _f, _p, _p, _p, _p, _f, _f, _f, _f, _i, _i, _p, _p, _p, _p, _p, _i, _p, _p, _p, _p]),
    "gg_sh_fwd": (C.c_int, [_i, _i, _i, _p, _p, _p, _p]),
    "gg_sh_bwd": (C.c_int, [_i, _i, _i, _p, _p, _p, _p]),
    "gg_cumsum_workspace_bytes": (C.c_size_t, [_ll]),
    "gg_cumsum": (C.c_int, [_ll, _p, _p, _p, _p, _sz, _p]),
    "gg_map_to_intersects": (C.c_int, [_i, _i, _p, _p, _p, _p, _i, _i, _p, _p, _p]),
    "gg_map_to_intersects_geo": (C.c_int, [_i, _i, _p, _p, _p, _p, _i, _i, _p, _p, _p]),
    "gg_sort_workspace_bytes": (C.c_size_t, [_ll]),
    "gg_sort_pairs": (C.c_int, [_ll, _i, _p, _p, _p, _p, _p, _sz, _p]),
    "gg_tile_ranges": (C.c_int, [_ll, _p, _ll, _p, _p]),
    "gg_blend_max_channels": (C.c_int, []),
    "gg_pack_geo": (C.c_int, [_ll, _i, _p, _p, _p, _i, _p, _p]),
    "gg_blend_fwd": (C.c_int, [_i, _ll, _i, _i, _i, _i, _i, _i, _i, _i] + [_p] * 12),
    "gg_blend_bwd": (C.c_int, [_i, _ll, _i, _i, _i, _i, _i, _i, _i, _i] + [_p] * 13),
    "gg_blend_hit_words": (C.c_size_t, [_ll, _ll, _i]),
    "gg_blend_pair_stats": (C.c_int, [_i, _ll, _i, _i, _i, _i, _p, _p, _p, _p, _p]),
    "gg_depth_keys": (C.c_int, [_ll, _i, _p, _p, _p, _p]),
    "gg_gather_counts": (C.c_int, [_ll, _p, _p, _p, _p]),
    "gg_emit_tiles_sorted": (C.c_int, [_i, _i, _p, _p, _i, _p, _p, _i, _i, _p, _p, _p]),
    "gg_tile_ranges_lowkey": (C.c_int, [_ll, _p, _ll, _p, _p]),
    "gg_bin_begin_scratch_bytes": (C.c_size_t, [_ll]),
    "gg_bin_finish_scratch_bytes": (C.c_size_t, [_ll]),
    "gg_bin_begin": (C.c_int, [_i, _i, _p, _p, _p, _sz, _p, _p]),
    "gg_bin_wait": (C.c_int, []),
    "gg_bin_finish": (C.c_int, [_i, _i, _ll, _p, _i, _p, _i, _i, _p, _p, _sz, _p, _p, _p, _p]),
    "gg_bin_tiles_scratch_bytes": (C.c_size_t, [_i, _ll, _ll]),
    "gg_bin_tiles": (C.c_int, [_i, _i, _p, _i, _p, _p, _i, _i, _ll, _p, _sz, _p, _p, _p, _p, _p, _i, _p]),
    "gg_tile_order_workspace_bytes": (C.c_size_t, []),
    "gg_tile_order": (C.c_int, [_ll, _p, _p, _p, _sz, _p]),
    "gg_unpack_vgeo": (C.c_int, [_ll, _i, _p, _p, _p, _p, _i, _p]),
    "gg_adam_step": (C.c_int, [_i, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _f, _f, _f, _p]),
    "gg_refine_workspace_bytes": (C.c_size_t, [_i]),
    "gg_refine_plan": (C.c_int, [_i, _p, _p, _p, _p, _p, _p, _p, _sz, _p, _p]),
    "gg_refine_apply": (C.c_int, [_i, _i, _p, _p, _i, _p, _p, _p, _p, _p, _p, _p, _p, C.c_ulonglong, C.c_uint, _p, _sz, _p]),
    "gg_philox_normals": (C.c_int, [_ll, _p, _i, C.c_ulonglong, C.c_uint, _p, _p]),
    "gg_philox4x32_10_host": (None, [_p, _p, _p]),
    "gg_knn3_scales": (C.c_int, [_i, _p, _p, _p, _p]),
    "gg_pixel_loss_workspace_bytes": (C.c_size_t, []),
    "gg_pixel_loss": (C.c_int, [_ll, _i, _p, _p, _p, _p, _i, _f, _p, _p, _p, _sz, _p]),
    "gg_loss_workspace_bytes": (C.c_size_t, []),
    "gg_geom_loss": (C.c_int, [_ll, _i, _p, _p, _p, _p, _p, _f, _f, _p, _i, _i, _p, _p, _sz, _p]),
    "gg_cosine_rows_loss": (C.c_int, [_ll, _i, _p, _ll, _p, _p, _ll, _p, _p, _p, _ll, _p, _ll, _p, _i, _p, _sz, _p]),
    "gg_param_regs": (C.c_int, [_ll, _i, _p, _p, _f, _f, _f, _p, _p, _p, _p, _sz, _p]),
    "gg_mlp_packed_floats": (C.c_size_t, []),
    "gg_mlp_pack_weights": (C.c_int, [_p, _p, _p, _p]),
    "gg_mlp_up": (C.c_int, [_ll, _p, _ll, _p, _p, _p, _p, _p]),
    "gg_ssim_workspace_bytes": (C.c_size_t, [_i, _i, _i, _i]),
    "gg_ssim_loss": (C.c_int, [_i, _i, _i, _i, _p, _i, _p, _i, _f, _p, _i, _i, _p, _p, _sz, _p]),
    "gg_densify_stats": (C.c_int, [_ll, _i, _p, _p, _i, _i, _i, _p, _p, _p, _p]),
    "gg_prepare_views": (C.c_int, [_i] * 6 + [_p] * 10 + [_i] * 4 + [_f] + [_p] * 7 + [_i, _p]),
    "gg_prepare_views_bwd": (C.c_int, [_i] * 6 + [_p] * 9 + [_i] * 2 + [_p] * 13),
    "gg_sh_grad_from_views": (C.c_int, [_i] * 4 + [_p] * 5),
    "gg_nvls_exchange": (C.c_int, [_i, _i, _p, _p, _ll, _p, _p, _ll, _i, _p]),
}

# optional symbols of later translation units (bound when present)
_OPTIONAL = {}


class GGError(RuntimeError):
    pass


def lib_path() -> str:
    return _build.LIB


def load():
    """Load (building in-tree first if needed) libgg_b200.so; raises if that is impossible."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        if _build.needs_build():
            try:
                _build.build()
            except Exception as e:
                # A library older than its sources must not be bound to these (newer) signatures: argument
                # drift would corrupt memory instead of raising.  Missing or stale, the answer is the same.
                state = "stale (older than csrc/ or the header)" if os.path.exists(_build.LIB) else "missing"
                raise ImportError(
                    f"gaussiangrasper_b200: libgg_b200.so is {state} and could not be rebuilt ({e}); "
                    "there is no CPU fallback") from e
        lib = C.CDLL(_build.LIB)
        lib.gg_version.restype = C.c_int
        if int(lib.gg_version()) != ABI_VERSION:
            raise ImportError(f"gaussiangrasper_b200: libgg_b200.so reports ABI version {int(lib.gg_version())}, the "
                              f"bindings expect {ABI_VERSION} (include/gg_b200.h GG_ABI_VERSION); rebuild with "
                              "python -m gaussiangrasper_b200.build --force")
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(lib, name)  # AttributeError if the library does not export the header's symbol
            fn.restype = res
            fn.argtypes = args
        for name, (res, args) in _OPTIONAL.items():
            if hasattr(lib, name):
                fn = getattr(lib, name)
                fn.restype = res
                fn.argtypes = args
        _lib = lib
    return _lib


def exported_symbols():
    return list(_SIGNATURES)


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        msg = load().gg_last_error_string().decode("utf-8", "replace")
        raise GGError(f"{what or 'gg_b200'} failed (code {rc}): {msg}")


# optional per-call CUDA-event timing (bench.py: "measured live inside the timed region")
_profile = None
_profile_only = None
_profile_external = False


class profile:
    """with _lib.profile() as prof: ...  ->  prof.ms() = {entry point: [ms per call]}
    only: restrict the event pairs to these entry points (every pair perturbs the stream a little)."""

    def __init__(self, only=None, external=False):
        """external: events that may be recorded inside a CUDA-graph capture (cudaEventRecordExternal); after a
        replay they hold the times of that replay."""
        self.only = frozenset(only) if only else None
        self.external = bool(external)

    def __enter__(self):
        global _profile, _profile_only, _profile_external
        self.records = {}
        _profile, _profile_only, _profile_external = self.records, self.only, self.external
        return self

    def __exit__(self, *exc):
        global _profile, _profile_only, _profile_external
        _profile = _profile_only = None
        _profile_external = False
        return False

    def ms(self):
        torch.cuda.synchronize()
        return {k: [a.elapsed_time(b) for a, b in v] for k, v in self.records.items()}


def call(name: str, *args) -> None:
    """Invoke a C-ABI entry point on torch's current stream and raise on a non-zero status."""
    fn = getattr(load(), name)
    if _profile is None or (_profile_only is not None and name not in _profile_only):
        rc = fn(*args)
    else:
        a = torch.cuda.Event(enable_timing=True, external=_profile_external)
        b = torch.cuda.Event(enable_timing=True, external=_profile_external)
        a.record()
        rc = fn(*args)
        b.record()
        _profile.setdefault(name, []).append((a, b))
    if rc != 0:
        check(rc, name)


def launch_count() -> int:
    return int(load().gg_launch_count())


def ptr(t):
    """Device pointer of a tensor (None -> NULL)."""
    if t is None:
        return None
    return t.data_ptr()


class _NoGuard:
    def __enter__(self):
        return None

    def __exit__(self, *exc):
        return False


_NO_GUARD = _NoGuard()


def device_guard(device):
    """`with torch.cuda.device(device)` only when `device` is not current already (the context
    manager costs microseconds per use, which adds up over the calls of one render)."""
    idx = device.index
    if idx is None or idx == torch.cuda.current_device():
        return _NO_GUARD
    return torch.cuda.device(device)


def stream_ptr(device) -> int:
    return torch.cuda.current_stream(device).cuda_stream


def require_cuda(*tensors) -> torch.device:
    dev = None
    for t in tensors:
        if t is None:
            continue
        if not t.is_cuda:
            raise GGError("gaussiangrasper_b200 runs on CUDA tensors only (no CPU fallback); got a "
                          f"{t.device} tensor")
        if dev is None:
            dev = t.device
        elif t.device != dev:
            raise GGError(f"tensors on different devices: {dev} vs {t.device}")
    return dev


def f32c(t: torch.Tensor) -> torch.Tensor:
    """float32, contiguous view/copy of t."""
    if t.dtype != torch.float32:
        t = t.float()
    return t.contiguous()
