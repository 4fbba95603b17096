"""In-tree build of libgg_b200.so with nvcc for sm_100a (no torch, no JIT cache).

python -m gaussiangrasper_b200.build [--force]
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
# GG_LIB_PATH: load this prebuilt library instead (kernel experiments: variants built with extra -D flags, see
# tools/build_variant.py); it is never rebuilt implicitly
LIB = os.environ.get("GG_LIB_PATH") or os.path.join(HERE, "libgg_b200.so")

ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "-I", os.path.join(ROOT, "include"), "-I", CSRC]
# (source, extra flags).  -fmad=false pins the fp32 operation order of every kernel whose
# results feed integer outputs (radii, tile counts, sort keys).
UNITS = [
    ("capi.cu", []),
    ("project.cu", ["-fmad=false"]),
    ("binning.cu", ["-fmad=false"]),
    ("binning2.cu", ["-fmad=false"]),
    ("blend.cu", []),
    ("fused.cu", ["-fmad=false"]),
    ("train.cu", ["-fmad=false"]),
    ("exchange.cu", []),
    ("loss.cu", ["-fmad=false"]),
    ("mlp.cu", []),
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found; cannot build libgg_b200.so")


def _host_cxx_flags():
    # nvcc's default host compiler must be a real g++ (the image's $CC/$CXX wrapper lacks some specs)
    for cand in ("/usr/bin/g++",):
        if os.path.exists(cand):
            return ["-ccbin", cand]
    return []


def needs_build() -> bool:
    if os.environ.get("GG_LIB_PATH"):
        if not os.path.exists(LIB):
            raise RuntimeError(f"GG_LIB_PATH={LIB} does not exist")
        return False
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(ROOT, "include", "gg_b200.h"), __file__]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile and link under an exclusive file lock (N ranks started by torchrun may all find the library stale
    at once); objects go to a per-process directory and the library is linked under a temporary name and moved
    into place atomically, so no process can ever load a half-written file."""
    if not force and not needs_build():
        return LIB
    import fcntl
    os.makedirs(os.path.join(HERE, "build"), exist_ok=True)
    with open(os.path.join(HERE, "build", ".lock"), "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            if not force and not needs_build():   # another process built it while we waited
                return LIB
            return _build_locked(verbose)
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)


def _build_locked(verbose: bool) -> str:
    nvcc = _nvcc()
    objdir = os.path.join(HERE, "build", f"obj.{os.getpid()}")
    os.makedirs(objdir, exist_ok=True)
    objs = []
    procs = []
    for src, extra in UNITS:
        path = os.path.join(CSRC, src)
        if not os.path.exists(path):
            continue
        obj = os.path.join(objdir, src.replace(".cu", ".o"))
        cmd = [nvcc, *ARCH, *COMMON, *_host_cxx_flags(), *extra, "-c", path, "-o", obj]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(obj)
    failed = False
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            failed = True
            sys.stderr.write(f"--- nvcc failed on {src} ---\n{out}\n")
        elif verbose or out.strip():
            sys.stderr.write(f"--- {src} ---\n{out}\n")
    if failed:
        raise RuntimeError("nvcc compilation failed")
    tmp = f"{LIB}.tmp.{os.getpid()}"
    link = [nvcc, *ARCH, *_host_cxx_flags(), "-shared", "-o", tmp, *objs, "-lcudart"]
    r = subprocess.run(link, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError("link failed:\n" + r.stdout)
    os.replace(tmp, LIB)
    shutil.rmtree(objdir, ignore_errors=True)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
