#!/bin/bash
mkdir -p gpurun_out
timeout 100 python -m pytest tests/test_gpu_parity.py tests/test_gpu_edge_cases.py -q -m gpu -p no:cacheprovider -x -k "fused or factored_sh or model_render" > gpurun_out/t_last.log 2>&1; echo "tests rc=$?"; tail -2 gpurun_out/t_last.log
timeout 60 python bench.py --config 3 --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('cfg3', round(d['value'],1), round(d['ms_per_step'],4), {k: round(v,4) for k,v in d['stage_ms_per_step'].items()})"
