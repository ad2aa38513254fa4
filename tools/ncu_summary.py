#!/usr/bin/env python
"""Key counters of one ncu report (raw page) as text + a traffic JSON.   python tools/ncu_summary.py REP MEMBERS OUT_PREFIX"""
import csv, io, json, subprocess, sys
rep, members, prefix = sys.argv[1], int(sys.argv[2]), sys.argv[3]
txt = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
hdr, units, vals = rows[0], rows[1], rows[2]
d = {h: (v, u) for h, v, u in zip(hdr, vals, units)}
keys = ['Kernel Name', 'gpu__time_duration.sum', 'launch__registers_per_thread', 'launch__block_size', 'launch__grid_size',
        'launch__shared_mem_per_block_dynamic', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__warps_eligible.avg.per_cycle_active',
        'inst_executed', 'sass__inst_executed_register_spilling', 'sass__inst_executed_shared_loads', 'sass__inst_executed_shared_stores', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_lsu.sum.pct_of_peak_sustained_active',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'lts__t_sector_hit_rate.pct', 'smsp__inst_executed_op_shared_ld.sum',
        'smsp__inst_executed_op_shared_st.sum', 'smsp__inst_executed_pipe_fp64.sum', 'sm__inst_executed_pipe_tensor_subpipe_dmma.sum']
def num(x):
    try: return float(x.replace(',', ''))
    except ValueError: return None
with open(prefix + '_ncu_full.txt', 'w') as fh:
    fh.write('# ncu --set full --clock-control none --import-source on -k regex:mpc_kernel -s 1 -c 1 python tools/prof_run.py transmon_h16 %d\n' % members)
    for k in keys:
        if k in d:
            fh.write('%-90s %s %s\n' % (k, d[k][0], d[k][1]))
    w = num(d['l1tex__data_pipe_lsu_wavefronts_mem_shared.sum'][0])
    fh.write('%-90s %.0f\n' % ('shared-memory wavefronts per member', w / members))
    fh.write('%-90s %.0f\n' % ('warp instructions per member', num(d['inst_executed'][0]) / members))
    stall = {h: num(v[0]) for h, v in d.items() if h.startswith('smsp__average_warps_issue_stalled') and h.endswith('_per_issue_active.ratio') and 'not_issued' not in h}
    tot = sum(x for x in stall.values() if x)
    fh.write('stall reasons (share of stalled warp-cycles per issue):\n')
    for h, x in sorted(stall.items(), key=lambda kv: -(kv[1] or 0))[:8]:
        fh.write('   %-60s %.1f%%\n' % (h.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', ''), 100 * x / tot))
def by(u, v):
    return v * {'Gbyte': 1e9, 'Mbyte': 1e6, 'Kbyte': 1e3, 'byte': 1}[u]
tr = {'members': members, 'kernel': 'mpc_kernel<Cfg<9,2>,false>', 'workload': 'transmon_h16',
      'dram_bytes_read': by(d['dram__bytes_read.sum'][1], num(d['dram__bytes_read.sum'][0])),
      'dram_bytes_write': by(d['dram__bytes_write.sum'][1], num(d['dram__bytes_write.sum'][0])),
      'gpu_time_ms': num(d['gpu__time_duration.sum'][0]) * {'ms': 1, 'us': 1e-3, 'ns': 1e-6, 's': 1e3}[d['gpu__time_duration.sum'][1]],
      'source': 'ncu --set full capture %s' % rep.split('/')[-1]}
json.dump(tr, open(prefix + '_ncu_traffic.json', 'w'), indent=1)
print(open(prefix + '_ncu_full.txt').read())
