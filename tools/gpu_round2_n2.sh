#!/bin/bash
# N-GPU call: symmetric-memory exchange kernel vs NCCL, then the bench at N ranks with both transports
N=${GG_N_GPUS:-2}
mkdir -p gpurun_out
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533"
timeout 300 $RUN tools/check_nvls_exchange.py > gpurun_out/nvls_check_n$N.log 2>&1
echo "check rc=$?"; grep -E "multicast|exchange|max \||NVLS_|Error|error" gpurun_out/nvls_check_n$N.log | head -30
timeout 400 $RUN bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/bench_n${N}_nvls.log 2>gpurun_out/bench_n${N}_nvls.err
echo "bench nvls rc=$?"; tail -1 gpurun_out/bench_n${N}_nvls.log | cut -c1-1800; tail -3 gpurun_out/bench_n${N}_nvls.err
timeout 400 $RUN bench.py --gpus $N --steps 20 --warmup 5 --transport nccl > gpurun_out/bench_n${N}_nccl.log 2>gpurun_out/bench_n${N}_nccl.err
echo "bench nccl rc=$?"; tail -1 gpurun_out/bench_n${N}_nccl.log | cut -c1-600; tail -3 gpurun_out/bench_n${N}_nccl.err
