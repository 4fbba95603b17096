#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -q -m gpu -p no:cacheprovider -x -k "bin or tile or overflow or fused or golden" > gpurun_out/t_bin.log 2>&1; echo "bin tests rc=$?"; tail -2 gpurun_out/t_bin.log
for c in 1 2; do
  st=20; [ $c = 2 ] && st=3
  timeout 400 python bench.py --config $c --steps $st --warmup 3 --no-cpu-baseline > gpurun_out/bench_cfg${c}_f.json 2> gpurun_out/bench_cfg${c}_f.err; echo "cfg$c rc=$?"
  python - $c <<'PY'
import json, sys
c=sys.argv[1]
d=json.loads(open(f"gpurun_out/bench_cfg{c}_f.json").read().strip().splitlines()[-1])
print("cfg", c, round(d["value"],1), round(d["ms_per_step"],4), "e2e", round(d["e2e"]["value"],1), {k: round(v,4) for k,v in d["stage_ms_per_step"].items()})
PY
done
