#!/bin/bash
# plain run first (must exit 0), then the ncu launch list and one full capture of the blend kernels
mkdir -p gpurun_out
CMD="python bench.py --steps 3 --warmup 3 --no-cpu-baseline"
$CMD > gpurun_out/ncu_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_list.log 2>&1
echo "list rc=$?"
$CMD > gpurun_out/ncu_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:blend_ -s 6 -c 2 -o gpurun_out/prof_blend -f $CMD > gpurun_out/ncu_full.log 2>&1
echo "full rc=$?"
tail -3 gpurun_out/ncu_full.log
timeout 600 python -m pytest tests -q -m gpu -k "sh_forward or end_to_end or rasterize_rgb" -p no:cacheprovider > gpurun_out/t_fix.log 2>&1
echo "tests rc=$?"; tail -5 gpurun_out/t_fix.log
