#!/bin/bash
N=${GG_N_GPUS:-2}
mkdir -p gpurun_out
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533"
for mode in 2 0 1; do
  GG_NVLS_MODE=$mode timeout 150 $RUN tools/check_nvls_exchange.py > gpurun_out/tune_n${N}_mode$mode.log 2>&1
  echo "N=$N mode=$mode: $(grep -E 'nvls: exchange|NVLS_' gpurun_out/tune_n${N}_mode$mode.log | tr '\n' ' ')"
done
