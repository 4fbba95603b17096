#!/bin/bash
# N-GPU call: exchange check (timings), bench with a timeline trace of the async and the e2e step
N=${GG_N_GPUS:-2}
mkdir -p gpurun_out
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533"
timeout 300 $RUN tools/check_nvls_exchange.py > gpurun_out/nvls_check_n$N.log 2>&1
echo "check rc=$?"; grep -E "multicast|exchange|NVLS_|Error|error" gpurun_out/nvls_check_n$N.log | head -30
timeout 400 $RUN bench.py --gpus $N --steps 20 --warmup 5 --trace gpurun_out/timeline_n$N.txt > gpurun_out/bench_n${N}_nvls.log 2>gpurun_out/bench_n${N}_nvls.err
echo "bench nvls rc=$?"; tail -1 gpurun_out/bench_n${N}_nvls.log | cut -c1-400; tail -3 gpurun_out/bench_n${N}_nvls.err
