#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -p no:cacheprovider -x > gpurun_out/t_all.log 2>&1; echo "gpu tests rc=$?"; tail -3 gpurun_out/t_all.log
run() { name=$1; shift; timeout 400 python bench.py "$@" --no-cpu-baseline > gpurun_out/$name.json 2>gpurun_out/$name.err; echo "$name rc=$?";
  python - "$name" <<'PY'
import json, sys
n=sys.argv[1]
d=json.loads(open(f"gpurun_out/{n}.json").read().strip().splitlines()[-1])
print(" ", d['config']['workload'], round(d["value"],1), round(d["ms_per_step"],4), "e2e", round(d["e2e"]["value"],1), {k: round(v,4) for k,v in d["stage_ms_per_step"].items()})
PY
}
GG_PREP_BWD_ODD=0 run cfg1_dense --steps 20 --warmup 5
GG_PREP_BWD_ODD=1 run cfg1_odd --steps 20 --warmup 5
run cfg2 --config 2 --steps 3 --warmup 3
run cfg3 --config 3 --steps 5 --warmup 3
