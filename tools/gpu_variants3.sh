#!/bin/bash
mkdir -p gpurun_out
GG_LIB_PATH=$PWD/gaussiangrasper_b200/variants/libgg_fmma.so timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_golden.py tests/test_gpu_parity_at_size.py tests/test_gpu_edge_cases.py -q -m gpu -p no:cacheprovider -k "not config2 and not config3" > gpurun_out/t_fmma.log 2>&1
echo "fmma parity rc=$?"; grep -E "^E  |passed|failed|^FAILED" gpurun_out/t_fmma.log | cut -c1-300 | head -12
for lib in gaussiangrasper_b200/variants/libgg_*.so; do
  name=$(basename $lib .so)
  GG_LIB_PATH=$PWD/$lib timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/var_$name.log 2>gpurun_out/var_$name.err
  python - "$name" <<'PY'
import json, sys
name=sys.argv[1]
try:
    d=json.loads(open(f"gpurun_out/var_{name}.log").read().strip().splitlines()[-1])
    st=d["stage_ms_per_step"]
    print(f"{name:16s} step {d['ms_per_step']:.4f} ms  fwd {st.get('gg_blend_fwd',0):.4f}  bwd {st.get('gg_blend_bwd',0):.4f}  e2e {d['e2e']['value']:.1f}")
except Exception as e:
    print(name, "failed", e)
PY
done
