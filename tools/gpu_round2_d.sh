#!/bin/bash
# round 2, call D: new tests (losses, captured step, two-rank refine), bench with the captured step, timeline
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_losses.py tests/test_gpu_training.py -q -m gpu -p no:cacheprovider > gpurun_out/t_new.log 2>&1
echo "tests rc=$?"; grep -E "^E  |passed|failed|^FAILED" gpurun_out/t_new.log | cut -c1-300 | head -20
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --trace gpurun_out/timeline_r2.txt > gpurun_out/bench_graph.log 2>gpurun_out/bench_graph.err
echo "bench graph rc=$?"; tail -1 gpurun_out/bench_graph.log | cut -c1-1200; tail -5 gpurun_out/bench_graph.err
python - <<'PY'
import json
try:
    d=json.loads(open("gpurun_out/bench_graph.log").read().strip().splitlines()[-1])
    print("value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"], "graph", d["cuda_graph"], "| timing:", d["blend_launch_timing"])
    print("stages", d["stage_ms_per_step"]); print("launches", d["gpu_launches"])
except Exception as e: print("parse failed", e)
PY
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-graph > gpurun_out/bench_nograph.log 2>&1
echo "bench nograph rc=$?"; tail -1 gpurun_out/bench_nograph.log | cut -c1-400
