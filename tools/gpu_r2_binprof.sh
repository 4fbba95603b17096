#!/bin/bash
# after the mapped-info change: binning tests + config 1 bench; then ncu --set full of the config-2 binning / prepare kernels
mkdir -p gpurun_out
timeout 600 python -m pytest tests -q -m gpu -p no:cacheprovider -x -k "bin or tile or overflow or fused or golden" > gpurun_out/t_bin.log 2>&1; echo "bin tests rc=$?"; tail -2 gpurun_out/t_bin.log
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_n1_d.json 2> gpurun_out/bench_n1_d.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open("gpurun_out/bench_n1_d.json").read().strip().splitlines()[-1])
print("cfg1", round(d["value"],1), round(d["ms_per_step"],4), "e2e", round(d["e2e"]["value"],1), {k: round(v,4) for k,v in d["stage_ms_per_step"].items()})
PY
timeout 500 ncu --set full --clock-control none --import-source on -k regex:"tile_|prepare_views_kernel" --launch-skip 140 -c 9 -f -o gpurun_out/cfg2_bin \
  python bench.py --config 2 --steps 1 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/ncu_cfg2.log 2>&1; echo "ncu rc=$?"; tail -3 gpurun_out/ncu_cfg2.log | cut -c1-300
ls -la gpurun_out/cfg2_bin.ncu-rep
