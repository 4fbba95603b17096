#!/bin/bash
# Round profile capture: plain bench, ncu launch list, one full capture of the hot kernels.
mkdir -p gpurun_out
CMD="python bench.py --steps 3 --warmup 3 --no-cpu-baseline"
python bench.py --steps 20 --warmup 5 > gpurun_out/bench_full.log 2>&1
echo "bench rc=$?"; tail -1 gpurun_out/bench_full.log | cut -c1-400
$CMD > gpurun_out/ncu_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_list.log 2>&1
echo "list rc=$?"
$CMD > gpurun_out/ncu_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'blend_|radix_pass|prepare_views' -s 18 -c 12 -o gpurun_out/prof_round -f $CMD > gpurun_out/ncu_full.log 2>&1
echo "full rc=$?"; tail -2 gpurun_out/ncu_full.log
