#!/bin/bash
# time the blend stages of every kernel variant under gaussiangrasper_b200/variants/
mkdir -p gpurun_out
for lib in gaussiangrasper_b200/variants/libgg_*.so; do
  name=$(basename $lib .so)
  GG_LIB_PATH=$PWD/$lib timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/var_$name.log 2>gpurun_out/var_$name.err
  python - "$name" <<'PY'
import json, sys
name=sys.argv[1]
try:
    d=json.loads(open(f"gpurun_out/var_{name}.log").read().strip().splitlines()[-1])
    st=d["stage_ms_per_step"]
    print(f"{name:14s} step {d['ms_per_step']:.4f} ms  fwd {st.get('gg_blend_fwd',0):.4f}  bwd {st.get('gg_blend_bwd',0):.4f}  e2e {d['e2e']['value']:.1f}")
except Exception as e:
    print(name, "failed", e)
PY
done
timeout 300 python -m pytest tests/test_gpu_training.py -q -m gpu -p no:cacheprovider -k knn > gpurun_out/t_knn.log 2>&1; echo "knn rc=$?"; grep -E "^E  |passed|failed" gpurun_out/t_knn.log | head -5
