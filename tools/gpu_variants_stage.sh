#!/bin/bash
# per-stage times of the default library and of every variant under gaussiangrasper_b200/variants/ (config 1)
mkdir -p gpurun_out
one() { name=$1; timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/var_$name.log 2>gpurun_out/var_$name.err
  python - "$name" <<'PY'
import json, sys
name=sys.argv[1]
try:
    d=json.loads(open(f"gpurun_out/var_{name}.log").read().strip().splitlines()[-1])
    print(f"{name:14s} step {d['ms_per_step']:.4f} ms  e2e {d['e2e']['value']:.1f}  ", {k: round(v, 4) for k, v in d["stage_ms_per_step"].items()})
except Exception as e:
    print(name, "failed", e)
PY
}
one base
for lib in gaussiangrasper_b200/variants/libgg_*.so; do
  GG_LIB_PATH=$PWD/$lib one $(basename $lib .so)
done
