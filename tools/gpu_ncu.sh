#!/bin/bash
# plain run first (must exit 0), then one full ncu capture of the blend kernels (+ optional extra regex)
mkdir -p gpurun_out
CMD="python bench.py --steps 3 --warmup 3 --no-cpu-baseline"
$CMD > gpurun_out/ncu_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:${GG_NCU_K:-blend_} -s ${GG_NCU_S:-6} -c ${GG_NCU_C:-2} -o gpurun_out/prof_${GG_NCU_NAME:-blend} -f $CMD > gpurun_out/ncu_full.log 2>&1
echo "full rc=$?"
tail -3 gpurun_out/ncu_full.log
