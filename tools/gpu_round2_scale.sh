#!/bin/bash
N=${GG_N_GPUS:-8}
mkdir -p gpurun_out
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533"
timeout 400 $RUN bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/bench_n${N}.json 2>gpurun_out/bench_n${N}.err
echo "bench rc=$?"; tail -2 gpurun_out/bench_n${N}.err | cut -c1-300
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_n${N}.json").read().strip().splitlines()[-1])
    print("N", d["n_gpus"], "value", round(d["value"],1), "ms", round(d["ms_per_step"],4), "e2e", round(d["e2e"]["value"],1), "|", d["exchange_transport"][:60])
    print("stages", {k: round(v,4) for k,v in d["stage_ms_per_step"].items()})
    c3=d.get("baseline_config3"); 
    if c3 and "value" in c3: print("config3:", round(c3["value"],1), "Mpix/s", round(c3["ms_per_step"],3), "ms; e2e", round(c3["e2e"]["value"],1), "|", c3["exchange_transport"][:70], "| stages", {k: round(v,3) for k,v in c3["stage_ms_per_step"].items()})
    else: print("config3:", c3)
except Exception as e: print("parse failed", e)
PY
