#!/bin/bash
# quick loop: selected tests + smoke + short bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -p no:cacheprovider ${GG_TEST_ARGS:-} > gpurun_out/t_all.log 2>&1
echo "tests rc=$?"; tail -15 gpurun_out/t_all.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1
echo "smoke rc=$?"; tail -3 gpurun_out/smoke.log
timeout 600 python bench.py --steps 10 --warmup 3 ${GG_BENCH_ARGS:-} > gpurun_out/bench.log 2>&1
echo "bench rc=$?"; tail -2 gpurun_out/bench.log
