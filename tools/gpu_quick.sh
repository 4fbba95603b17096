#!/bin/bash
# quick loop: selected tests + smoke + short bench.  GG_TEST_K = pytest -k expression (optional)
mkdir -p gpurun_out
if [ -n "$GG_TEST_K" ]; then
  timeout 900 python -m pytest tests -q -m gpu -p no:cacheprovider -k "$GG_TEST_K" > gpurun_out/t_all.log 2>&1
else
  timeout 900 python -m pytest tests -q -m gpu -p no:cacheprovider > gpurun_out/t_all.log 2>&1
fi
echo "tests rc=$?"; tail -15 gpurun_out/t_all.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1
echo "smoke rc=$?"; tail -3 gpurun_out/smoke.log
timeout 600 python bench.py --steps 10 --warmup 3 ${GG_BENCH_ARGS:-} > gpurun_out/bench.log 2>&1
echo "bench rc=$?"; tail -2 gpurun_out/bench.log
