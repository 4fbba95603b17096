#!/bin/bash
N=${GG_N_GPUS:-2}
mkdir -p gpurun_out
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533"
timeout 300 $RUN tools/check_nvls_exchange.py > gpurun_out/nvls_check_n$N.log 2>&1
echo "check rc=$?"; grep -E "multicast|exchange|NVLS_|Error|error" gpurun_out/nvls_check_n$N.log | head -30
timeout 300 $RUN bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/bench_n${N}_nvls.log 2>gpurun_out/bench_n${N}_nvls.err
echo "bench nvls rc=$?"; tail -3 gpurun_out/bench_n${N}_nvls.err | cut -c1-300
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_n${N}_nvls.log").read().strip().splitlines()[-1])
    print("N", d["n_gpus"], "value", round(d["value"],1), "ms", round(d["ms_per_step"],4), "e2e", round(d["e2e"]["value"],1), d["e2e"].get("mode"), "|", d["exchange_transport"], "|", d["cuda_graph"])
    print("stages", {k: round(v,4) for k,v in d["stage_ms_per_step"].items()})
except Exception as e: print("parse failed", e)
PY
