#!/bin/bash
# round 2, final single-GPU pass: whole GPU suite, smoke, bench, reference arm, launch list, full ncu capture, FLOP counters
mkdir -p gpurun_out; rm -f gpurun_out/at_size.jsonl
GG_AT_SIZE_REPORT=gpurun_out/at_size.jsonl timeout 1500 python -m pytest tests -q -m gpu -p no:cacheprovider --durations=6 > gpurun_out/t_all.log 2>&1
echo "tests rc=$?"; grep -E "^E  |passed|failed|^FAILED" gpurun_out/t_all.log | cut -c1-300 | head -12
timeout 200 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench.log 2>gpurun_out/bench.err
echo "bench rc=$?"; tail -2 gpurun_out/bench.err | cut -c1-200
timeout 900 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/bench_reference.log 2>gpurun_out/bench_reference.err
echo "reference rc=$?"; tail -1 gpurun_out/bench_reference.log | cut -c1-700
CMD="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-graph"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_list.log 2>&1
echo "list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:'blend_|tile_|prepare_views|pixel_loss' -s 150 -c 14 -o gpurun_out/prof_r2 -f $CMD > gpurun_out/ncu_full.log 2>&1
echo "full rc=$?"
ncu --metrics smsp__sass_thread_inst_executed_op_ffma_pred_on.sum,smsp__sass_thread_inst_executed_op_fadd_pred_on.sum,smsp__sass_thread_inst_executed_op_fmul_pred_on.sum,smsp__thread_inst_executed.sum,sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_tensor.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:'blend_' -s 20 -c 4 --csv --log-file gpurun_out/flops.csv $CMD > gpurun_out/ncu_flops.log 2>&1
echo "flops rc=$?"
ncu --set full --clock-control none --import-source on -k regex:'mlp_up' -s 3 -c 1 -o gpurun_out/prof_mlp -f python tools/bench_mlp.py > gpurun_out/ncu_mlp.log 2>&1; echo "ncu mlp rc=$?"
timeout 100 python tools/bench_mlp.py > gpurun_out/mlp_bench.json 2>/dev/null; cat gpurun_out/mlp_bench.json | cut -c1-400
