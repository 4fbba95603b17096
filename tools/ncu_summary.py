#!/usr/bin/env python
"""Key metrics of every kernel in an .ncu-rep (raw page): usage ncu_summary.py file.ncu-rep"""
import csv, subprocess, sys, io

KEYS = ['gpu__time_duration.sum', 'launch__registers_per_thread', 'launch__occupancy_limit_registers',
        'launch__occupancy_limit_shared_mem', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__inst_executed.sum', 'sm__inst_executed.avg.per_cycle_elapsed',
        'smsp__thread_inst_executed_per_inst_executed.ratio', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'sm__cycles_active.avg', 'sm__cycles_elapsed.max', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'lts__t_sectors_op_red.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'smsp__sass_thread_inst_executed_op_ffma_pred_on.sum', 'smsp__sass_thread_inst_executed_op_fadd_pred_on.sum',
        'smsp__sass_thread_inst_executed_op_fmul_pred_on.sum', 'smsp__thread_inst_executed.sum',
        'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
        'sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.avg.pct_of_peak_sustained_elapsed']

def main(path):
    out = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        print('=====', d['Kernel Name'][:70])
        for k in KEYS:
            if k in d:
                print(f"  {k:62s} {d[k]} {units[hdr.index(k)]}")
        st = {}
        for k, v in d.items():
            if 'pcsamp_warps_issue_stalled' in k and 'not_issued' not in k:
                try:
                    st[k.replace('smsp__pcsamp_warps_issue_stalled_', '')] = float(v.replace(',', ''))
                except ValueError:
                    pass
        tot = sum(st.values()) or 1
        print('  stalls: ' + ', '.join(f"{k} {v/tot*100:.0f}%" for k, v in sorted(st.items(), key=lambda x: -x[1])[:8]))

if __name__ == '__main__':
    main(sys.argv[1])
