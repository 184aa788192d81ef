#!/usr/bin/env python
"""Print the metrics that matter from an .ncu-rep: python scripts/ncu_keys.py file.ncu-rep [kernel-substring]"""
import csv, subprocess, sys
rep = sys.argv[1]
filt = sys.argv[2] if len(sys.argv) > 2 else ""
txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
hdr = rows[0]
EXACT = ['gpu__time_duration.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
         'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum', 'sm__cycles_elapsed.max',
         'smsp__warps_eligible.avg.per_cycle_active', 'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active',
         'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
         'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed',
         'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
         'dram__bytes_read.sum', 'dram__bytes_write.sum', 'lts__t_bytes.sum', 'launch__registers_per_thread',
         'launch__shared_mem_per_block_dynamic', 'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_registers',
         'launch__grid_size', 'launch__block_size', 'l1tex__t_sectors_pipe_lsu_mem_local_op_ld.sum',
         'l1tex__t_sectors_pipe_lsu_mem_local_op_st.sum', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
         'gpu__compute_memory_throughput.avg.pct_of_peak_sustained_elapsed', 'dram__throughput.avg.pct_of_peak_sustained_elapsed',
         'lts__t_sector_hit_rate.pct']
for r in rows[2:]:
    d = dict(zip(hdr, r))
    if filt and filt not in d.get('Kernel Name', ''):
        continue
    print('==', d.get('Kernel Name', '')[:120])
    for k in EXACT:
        if k in d: print('  %-75s %s' % (k, d[k]))
    st = []
    for k in hdr:
        if 'issue_stalled' in k and (k.endswith('_per_warp_active.pct') or k.endswith('_per_issue_active.ratio')) \
                and 'not_issued' not in k:
            try: st.append((float(d[k].replace(',', '')), k))
            except ValueError: pass
    tot = sum(v for v, _ in st) or 1.0
    for v, k in sorted(st, reverse=True)[:6]:
        print('  stall %5.1f%% of the stalled warp-cycles  %s' % (100 * v / tot, k.split('issue_stalled_')[1].split('_per_')[0]))
