#!/bin/bash
N=${GG_N_GPUS:-2}
mkdir -p gpurun_out
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533"
timeout 300 $RUN bench.py --gpus $N --steps 20 --warmup 5 --diag > gpurun_out/bench_n${N}_diag.log 2>gpurun_out/bench_n${N}_diag.err
echo "bench rc=$?"; grep diag gpurun_out/bench_n${N}_diag.err
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_n${N}_diag.log").read().strip().splitlines()[-1])
    print("N", d["n_gpus"], "value", round(d["value"],1), "ms", round(d["ms_per_step"],4), "e2e", round(d["e2e"]["value"],1))
except Exception as e: print("parse failed", e)
PY
