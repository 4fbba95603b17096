"""Build a variant of libgg_b200.so with extra -D flags for kernel experiments:
    python tools/build_variant.py NAME -DGG_BWD_DOT4=1 ...   ->  gaussiangrasper_b200/variants/libgg_NAME.so
Run it with GG_LIB_PATH=<that file> python bench.py ...  (objects of untouched units are reused)."""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gaussiangrasper_b200 import build as B

name, flags = sys.argv[1], sys.argv[2:]
out_dir = os.path.join(B.HERE, "variants")   # (build/ is not shipped to the GPU box)
obj_dir = os.path.join(out_dir, "obj_base")
os.makedirs(obj_dir, exist_ok=True)
nvcc = B._nvcc()
objs, procs = [], []
for src, extra in B.UNITS:
    path = os.path.join(B.CSRC, src)
    touched = src in os.environ.get("GG_VARIANT_UNITS", "blend.cu").split(",")   # units the -D flags apply to
    obj = os.path.join(out_dir if touched else obj_dir, (f"{name}_" if touched else "") + src.replace(".cu", ".o"))
    objs.append(obj)
    if touched or not os.path.exists(obj) or os.path.getmtime(obj) < os.path.getmtime(path):
        cmd = [nvcc, *B.ARCH, *B.COMMON, *B._host_cxx_flags(), *extra, *(flags if touched else []), "-c", path, "-o", obj]
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
for src, p in procs:
    out, _ = p.communicate()
    if p.returncode:
        sys.exit(f"nvcc failed on {src}:\n{out}")
lib = os.path.join(out_dir, f"libgg_{name}.so")
r = subprocess.run([nvcc, *B.ARCH, *B._host_cxx_flags(), "-shared", "-o", lib, *objs, "-lcudart"], capture_output=True, text=True)
if r.returncode:
    sys.exit(r.stdout + r.stderr)
print(lib)
