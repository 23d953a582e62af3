"""Extract the metrics quoted in profiles/README.md from an ncu report: python tools/ncu_select.py report.ncu-rep out.csv"""
import csv, subprocess, sys
rep, out = sys.argv[1], sys.argv[2]
raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.split('\n')))
h, units, v = rows[0], rows[1], rows[2]
want = ['Kernel Name', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__time_duration.sum', 'launch__registers_per_thread', 'launch__block_size', 'launch__grid_size',
        'launch__shared_mem_per_block_dynamic', 'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active', 'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_tensor.avg.pct_of_peak_sustained_active', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
        'smsp__thread_inst_executed_per_inst_executed.ratio', 'sm__warps_active.avg.pct_of_peak_sustained_active', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'dram__bytes_write.sum.per_second', 'lts__t_sector_hit_rate.pct', 'sm__cycles_elapsed.max',
        'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active', 'sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active', 'sm__warps_active.avg.per_cycle_active',
        'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_sleeping_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
        'smsp__sass_inst_executed_op_tmem_ldt.sum', 'smsp__sass_inst_executed_op_tmem_stt.sum']
with open(out, 'w') as f:
    f.write('metric,unit,value\n')
    for k in want:
        for i, n in enumerate(h):
            if n == k:
                f.write(f'{k},{units[i]},{v[i]}\n')
                break
print(open(out).read())
