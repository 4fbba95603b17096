#!/bin/bash
# round 2, call C: two-rank divergence diagnosis, at-size parity, training tests, bench line, executed-FLOP counters
mkdir -p gpurun_out; rm -f gpurun_out/at_size.jsonl
timeout 300 python tools/debug_two_rank.py > gpurun_out/dbg2.log 2>&1; echo "dbg reset rc=$?"; grep -v "same$" gpurun_out/dbg2.log | tail -15
timeout 300 python tools/debug_two_rank.py --no-reset > gpurun_out/dbg2n.log 2>&1; echo "dbg noreset rc=$?"; grep -v "same$" gpurun_out/dbg2n.log | tail -15
GG_AT_SIZE_REPORT=gpurun_out/at_size.jsonl timeout 1400 python -m pytest tests/test_gpu_parity_at_size.py -q -m gpu -p no:cacheprovider > gpurun_out/t_at_size.log 2>&1
echo "at-size rc=$?"; grep -E "^E  |passed|failed" gpurun_out/t_at_size.log | head -30
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench.log 2>gpurun_out/bench.err
echo "bench rc=$?"; tail -1 gpurun_out/bench.log | cut -c1-6000; tail -5 gpurun_out/bench.err
CMD="python bench.py --steps 3 --warmup 3 --no-cpu-baseline"
ncu --metrics smsp__sass_thread_inst_executed_op_ffma_pred_on.sum,smsp__sass_thread_inst_executed_op_fadd_pred_on.sum,smsp__sass_thread_inst_executed_op_fmul_pred_on.sum,smsp__thread_inst_executed.sum,sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:'blend_' -s 20 -c 4 --csv --log-file gpurun_out/flops.csv $CMD > gpurun_out/ncu_flops.log 2>&1
echo "flops rc=$?"
