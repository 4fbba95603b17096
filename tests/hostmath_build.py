"""Builds tests/_hostmath.so (host compile of csrc/gg_math.cuh) for the CPU-side math tests."""
import ctypes as C
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
SO = os.path.join(HERE, "_hostmath.so")


def load():
    src = os.path.join(HERE, "hostmath.cpp")
    hdr = os.path.join(ROOT, "gaussiangrasper_b200", "csrc", "gg_math.cuh")
    if not os.path.exists(SO) or os.path.getmtime(SO) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
        cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
        subprocess.run([cxx, "-O2", "-ffp-contract=off", "-fno-fast-math", "-fPIC", "-shared", "-std=c++17",
                        "-x", "c++", "-I", os.path.dirname(hdr), src, "-o", SO], check=True)
    return C.CDLL(SO)
