#!/bin/bash
# round 2: config 1 again (after the prepare / loss changes), then BASELINE configs 2, 3 and the channel sweep of config 4
mkdir -p gpurun_out
run() { name=$1; shift; timeout 400 python bench.py "$@" --no-cpu-baseline > gpurun_out/$name.json 2>gpurun_out/$name.err; echo "$name rc=$?";
  python - "$name" <<'PY'
import json, sys
n=sys.argv[1]
try:
    d=json.loads(open(f"gpurun_out/{n}.json").read().strip().splitlines()[-1])
    st={k: round(v,4) for k,v in d["stage_ms_per_step"].items()}
    rf=d["roofline"] or {}
    print(f"  {d['config']['workload']} C={d['config']['channels']}: {d['value']:.1f} Mpix/s, {d['ms_per_step']:.4f} ms, e2e {d['e2e']['value']:.1f}; roofline {rf.get('kernel')} frac {rf.get('frac')}; stages {st}")
    print("  hbm", {k:(round(v['GB/s']), round(v['frac'],3)) for k,v in d['hbm_stages'].items()})
except Exception as e: print("  parse failed", e)
PY
}
run cfg1 --steps 20 --warmup 5
timeout 600 python -m pytest tests/test_gpu_parity_at_size.py tests/test_golden.py tests/test_gpu_training.py -q -m gpu -p no:cacheprovider -k "config1 or small_case or golden or pixel_loss or main_loss" > gpurun_out/t_prep.log 2>&1; echo "parity rc=$?"; grep -E "^E  |passed|failed" gpurun_out/t_prep.log | head -5
run cfg2 --config 2 --steps 3 --warmup 3
run cfg3 --config 3 --steps 5 --warmup 3
for D in 3 16 32 64; do run cfg4_D$D --config 4 --feat $D --steps 5 --warmup 3; done
