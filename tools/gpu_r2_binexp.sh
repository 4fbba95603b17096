#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; timeout 300 python bench.py --config 2 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/$name.json 2>gpurun_out/$name.err;
  python - "$name" <<'PY'
import json, sys
n=sys.argv[1]
try:
    d=json.loads(open(f"gpurun_out/{n}.json").read().strip().splitlines()[-1])
    print(n, round(d["value"],1), round(d["ms_per_step"],3), {k: round(v,3) for k,v in d["stage_ms_per_step"].items()})
except Exception as e: print(n, "failed", e)
PY
}
run binexp_base
GG_BIN_BLOCKS=148 run binexp_b148
GG_BIN_BLOCKS=224 run binexp_b224
GG_LIB_PATH=$PWD/gaussiangrasper_b200/variants/libgg_ldcs.so run binexp_ldcs
GG_BIN_BLOCKS=148 GG_LIB_PATH=$PWD/gaussiangrasper_b200/variants/libgg_ldcs.so run binexp_ldcs_b148
