"""Print the handful of ncu raw metrics we look at after every capture: python tools/ncu_key.py <rep>"""
import csv, subprocess, sys
raw = subprocess.run(f"ncu -i {sys.argv[1]} --page raw --csv", shell=True, capture_output=True, text=True).stdout
rows = list(csv.reader(raw.split('\n')))
h, u = rows[0], rows[1]
want = ['gpu__time_duration.sum', 'smsp__inst_executed.sum', 'sm__inst_executed.avg.per_cycle_active', 'sm__inst_executed.avg.per_cycle_elapsed',
        'smsp__thread_inst_executed_per_inst_executed.ratio', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum', 'l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum', 'l1tex__t_sector_hit_rate.pct',
        'lts__t_sector_hit_rate.pct', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'l1tex__data_pipe_lsu_wavefronts.sum',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'l1tex__lsu_writeback_active.avg.pct_of_peak_sustained_elapsed',
        'l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'launch__registers_per_thread', 'launch__occupancy_limit',
        'smsp__average_warps_issue_stalled', 'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed_op_global',
        'smsp__inst_executed_op_shared', 'l1tex__t_set_accesses', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed_pipe_fp64', 'sm__inst_executed_pipe_lsu',
        'smsp__inst_executed_pipe_lsu', 'smsp__inst_executed_pipe_alu', 'smsp__inst_executed_pipe_fma', 'smsp__inst_executed_pipe_xu',
        'sm__cycles_active.avg', 'l1tex__t_requests_pipe_lsu_mem_global_op_red', 'l1tex__t_sectors_pipe_lsu_mem_global_op_red', 'l1tex__t_sectors_pipe_lsu_mem_global_op_st']
for r in rows[2:]:
    if len(r) < len(h): continue
    print('==', r[h.index('Kernel Name')] if 'Kernel Name' in h else '')
    for i, n in enumerate(h):
        if any(n.startswith(w) for w in want) and not ('.pct_of_peak' in n and not any(n == w for w in want)) and '.per_second' not in n and '.peak_sustained' not in n:
            try:
                v = float(r[i].replace(',', ''))
            except ValueError:
                continue
            if v != 0: print('  %-90s %-12s %s' % (n, u[i], r[i]))
