"""Turn the ncu captures of tools/collect_profiles.sh into the committed summaries:
python tools/summarize_profiles.py <tag>   ->  profiles/<tag>_ncu_full_summary.csv, profiles/ncu_traffic.json,
profiles/<tag>_k_route_source_hotspots.txt, and copies of the bench lines / launch list."""
import csv, json, os, shutil, subprocess, sys
tag = sys.argv[1]
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
os.chdir(root)
want = ['gpu__time_duration.sum', 'smsp__inst_executed.sum', 'sm__inst_executed.avg.per_cycle_active', 'smsp__thread_inst_executed_per_inst_executed.ratio',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem',
        'launch__grid_size', 'launch__block_size', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'dram__throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__t_sector_hit_rate.pct', 'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__t_sector_hit_rate.pct',
        'l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed',
        'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum', 'l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'smsp__issue_active.avg.pct_of_peak_sustained_active'] + ['smsp__average_warps_issue_stalled_%s_per_issue_active.ratio' % x for x in (
            'long_scoreboard', 'short_scoreboard', 'wait', 'barrier', 'branch_resolving', 'not_selected', 'math_pipe_throttle', 'lg_throttle', 'mio_throttle')]
caps = [('k_route', 2), ('k_scan', 2), ('k_emit', 2), ('k_rank_scatter', 2), ('k_emit', 3), ('k_route', 3), ('k_route', 4), ('k_runs', 4)]
out, traffic = [['capture', 'kernel', 'metric', 'unit', 'value']], []
for k, c in caps:
    rep = 'gpurun_out/%s_%s_c%d.ncu-rep' % (tag, k, c)
    if not os.path.exists(rep): continue
    rows = list(csv.reader(subprocess.run('ncu -i %s --page raw --csv' % rep, shell=True, capture_output=True, text=True).stdout.split('\n')))
    h, u = rows[0], rows[1]
    cand = [r for r in rows[2:] if len(r) == len(h)]
    ti = h.index('gpu__time_duration.sum')
    d = max(cand, key=lambda r: float(r[ti].replace(',', '')))  # (k_emit launches twice per step: blocks, then the segments of long blocks)
    vals = {}
    for i, n in enumerate(h):
        if n in want:
            out.append(['%s_c%d' % (k, c), d[h.index('Kernel Name')], n, u[i], d[i]])
            vals[n] = (u[i], d[i])
    tob = lambda n: float(vals[n][1].replace(',', '')) * {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}[vals[n][0]]
    nb = json.loads([l for l in open('gpurun_out/%s_bench_config%d.json' % (tag, c)) if l.startswith('{')][-1])['config']['bytes_per_gpu_per_step']
    traffic.append({'kernel': k, 'config': c, 'input_bytes': nb, 'dram_read_bytes': tob('dram__bytes_read.sum'), 'dram_write_bytes': tob('dram__bytes_write.sum'),
                    'l2_hit_rate_pct': float(vals['lts__t_sector_hit_rate.pct'][1]), 'l1_hit_rate_pct': float(vals['l1tex__t_sector_hit_rate.pct'][1]),
                    'lts_throughput_pct': float(vals['lts__throughput.avg.pct_of_peak_sustained_elapsed'][1]),
                    'capture': 'profiles/%s_ncu_full_summary.csv (%s_c%d)' % (tag, k, c)})
csv.writer(open('profiles/%s_ncu_full_summary.csv' % tag, 'w')).writerows(out)
json.dump(traffic, open('profiles/ncu_traffic.json', 'w'), indent=1)
shutil.copy('gpurun_out/%s_ncu_launches_config2.csv' % tag, 'profiles/')
if os.path.exists('gpurun_out/%s_ncu_launches_config4.csv' % tag): shutil.copy('gpurun_out/%s_ncu_launches_config4.csv' % tag, 'profiles/')
for c in ('config2', 'config3', 'config4', 'reference_arm'):
    open('profiles/%s_bench_%s.json' % (tag, c), 'w').write([l for l in open('gpurun_out/%s_bench_%s.json' % (tag, c)) if l.startswith('{')][-1])
for extra in ('latency.json', 'pcie_1gpu.json'):
    if os.path.exists('gpurun_out/%s_%s' % (tag, extra)): shutil.copy('gpurun_out/%s_%s' % (tag, extra), 'profiles/%s_%s' % (tag, extra))
hot = subprocess.run('python tools/ncu_lines.py gpurun_out/%s_k_route_c2.ncu-rep k_routeILi16 jb_stream.cu 45' % tag, shell=True, capture_output=True, text=True).stdout
open('profiles/%s_k_route_source_hotspots.txt' % tag, 'w').write('k_route<16,4>, config 2 (1e9 B), ncu --set full, per CUDA source line (tools/ncu_lines.py)\n' + hot)
for t in traffic: print(t['kernel'], t['config'], round(t['dram_read_bytes'] / 1e6), round(t['dram_write_bytes'] / 1e6))
