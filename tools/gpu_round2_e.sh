#!/bin/bash
# round 2, call E: bench (captured step), up-projection timing + ncu capture, launch list
mkdir -p gpurun_out
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench.log 2>gpurun_out/bench.err
echo "bench rc=$?"; tail -3 gpurun_out/bench.err | cut -c1-300
python - <<'PY'
import json
d=json.loads(open("gpurun_out/bench.log").read().strip().splitlines()[-1])
print("value", round(d["value"],1), "ms", round(d["ms_per_step"],4), "e2e", round(d["e2e"]["value"],1), "ratio", round(d["e2e"]["value"]/d["value"],3))
print("cpu", d["cpu_baseline"]["value"], d["cpu_baseline"]["sample"][:120])
PY
timeout 200 python tools/bench_mlp.py > gpurun_out/mlp_bench.json 2>gpurun_out/mlp_bench.err; echo "mlp bench rc=$?"; cat gpurun_out/mlp_bench.json; tail -2 gpurun_out/mlp_bench.err
ncu --set full --clock-control none --import-source on -k regex:'mlp_up' -s 3 -c 1 -o gpurun_out/prof_mlp -f python tools/bench_mlp.py > gpurun_out/ncu_mlp.log 2>&1; echo "ncu mlp rc=$?"
CMD="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-graph"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_list.log 2>&1
echo "list rc=$?"
