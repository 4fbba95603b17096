#!/bin/bash
# round 2, closing single-GPU pass after the late prepare / binning / e2e changes: GPU suite, smoke, bench, launch list,
# full ncu capture of the step's kernels, the other BASELINE configs
mkdir -p gpurun_out; rm -f gpurun_out/at_size.jsonl
GG_AT_SIZE_REPORT=gpurun_out/at_size.jsonl timeout 1500 python -m pytest tests -q -m gpu -p no:cacheprovider > gpurun_out/t_all.log 2>&1
echo "tests rc=$?"; grep -E "^E  |passed|failed|^FAILED" gpurun_out/t_all.log | cut -c1-300 | head -12
timeout 200 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench.log 2>gpurun_out/bench.err
echo "bench rc=$?"; tail -2 gpurun_out/bench.err | cut -c1-200
CMD="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-graph"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_list.log 2>&1
echo "list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:'blend_|tile_|prepare_views|pixel_loss' -s 150 -c 14 -o gpurun_out/prof_r2 -f $CMD > gpurun_out/ncu_full.log 2>&1
echo "full rc=$?"
run() { name=$1; shift; timeout 400 python bench.py "$@" --no-cpu-baseline > gpurun_out/$name.json 2>gpurun_out/$name.err; echo "$name rc=$?";
  python - "$name" <<'PY'
import json, sys
n=sys.argv[1]
try:
    d=json.loads(open(f"gpurun_out/{n}.json").read().strip().splitlines()[-1])
    print(" ", d['config']['workload'], "C", d['config']['channels'], round(d["value"],1), round(d["ms_per_step"],4), "e2e", round(d["e2e"]["value"],1), {k: round(v,4) for k,v in d["stage_ms_per_step"].items()})
except Exception as e: print("  parse failed", e)
PY
}
run cfg2 --config 2 --steps 3 --warmup 3
run cfg3 --config 3 --steps 5 --warmup 3
for D in 3 16 32 64; do run cfg4_D$D --config 4 --feat $D --steps 5 --warmup 3; done
python - <<'PY'
import json
d=json.loads(open("gpurun_out/bench.log").read().strip().splitlines()[-1])
print("cfg1", round(d["value"],1), round(d["ms_per_step"],4), d["e2e"], {k: round(v,4) for k,v in d["stage_ms_per_step"].items()}, d["cpu_baseline"])
PY
