#!/bin/bash
# round 2, call A: full GPU suite, smoke, bench (config 1), launch list
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu -p no:cacheprovider --maxfail=30 > gpurun_out/t_all.log 2>&1
echo "tests rc=$?"; tail -30 gpurun_out/t_all.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1
echo "smoke rc=$?"; tail -3 gpurun_out/smoke.log
timeout 600 python bench.py --steps 20 --warmup 5 ${GG_BENCH_ARGS:-} > gpurun_out/bench.log 2>gpurun_out/bench.err
echo "bench rc=$?"; tail -2 gpurun_out/bench.log | cut -c1-3000; tail -5 gpurun_out/bench.err
CMD="python bench.py --steps 3 --warmup 3 --no-cpu-baseline"
$CMD > gpurun_out/ncu_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_list.log 2>&1
echo "list rc=$?"
