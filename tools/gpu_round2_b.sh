#!/bin/bash
# round 2, call B: GPU suite, bench, launch list, one full ncu capture of the hot kernels
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu -p no:cacheprovider --maxfail=30 > gpurun_out/t_all.log 2>&1
echo "tests rc=$?"; tail -8 gpurun_out/t_all.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench.log 2>gpurun_out/bench.err
echo "bench rc=$?"; tail -2 gpurun_out/bench.log | cut -c1-3000; tail -5 gpurun_out/bench.err
CMD="python bench.py --steps 3 --warmup 3 --no-cpu-baseline"
$CMD > gpurun_out/ncu_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_list.log 2>&1
echo "list rc=$?"
$CMD > gpurun_out/ncu_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'blend_|tile_|prepare_views' -s ${GG_NCU_S:-110} -c ${GG_NCU_C:-12} -o gpurun_out/prof_r2 -f $CMD > gpurun_out/ncu_full.log 2>&1
echo "full rc=$?"; tail -2 gpurun_out/ncu_full.log
