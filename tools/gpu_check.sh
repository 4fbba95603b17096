#!/bin/bash
# One gpurun call: sort/scan kernels first (look-back spin is the riskiest code), then the whole
# GPU test-suite, smoke, and a short bench.  Everything is wrapped in `timeout`.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
(ls -la baseline/_ref 2>&1; python -m pip list 2>/dev/null | grep -i -E "gsplat|nerfstudio") > gpurun_out/probe.txt 2>&1
echo "== sort/scan ==" 
timeout 300 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "radix or cumsum" -x > gpurun_out/t_sort.log 2>&1
echo "sort rc=$?"; tail -5 gpurun_out/t_sort.log
echo "== all gpu tests =="
timeout 1200 python -m pytest tests -q -m gpu --maxfail=40 -p no:cacheprovider > gpurun_out/t_all.log 2>&1
echo "tests rc=$?"; tail -40 gpurun_out/t_all.log
echo "== smoke =="
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1
echo "smoke rc=$?"; tail -5 gpurun_out/smoke.log
echo "== bench =="
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.log 2>&1
echo "bench rc=$?"; tail -3 gpurun_out/bench.log
