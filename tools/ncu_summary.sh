#!/bin/bash
# usage: tools/ncu_summary.sh report.ncu-rep   -- key metrics + stall breakdown of every kernel in the report
ncu -i "$1" --page raw --csv 2>/dev/null | python -c "
import csv,sys
rows=list(csv.reader(sys.stdin))
hdr=rows[0]
for r in rows[2:]:
    d=dict(zip(hdr,r))
    print(d['Kernel Name'][:60], 'grid', d.get('launch__grid_size'), 'block', d.get('launch__block_size'))
    keys=['gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed','sm__throughput.avg.pct_of_peak_sustained_elapsed','smsp__issue_active.avg.pct_of_peak_sustained_active','sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active','smsp__inst_executed.sum','launch__registers_per_thread','launch__occupancy_limit_registers','launch__occupancy_limit_shared_mem','sm__warps_active.avg.pct_of_peak_sustained_active','l1tex__data_pipe_lsu_wavefronts_mem_shared.sum','l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum','smsp__cycles_active.avg','sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active','sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active','sm__pipe_xu_cycles_active.avg.pct_of_peak_sustained_active','lts__t_bytes.sum','lts__t_sectors_srcunit_tex_op_read.sum','sm__cycles_elapsed.max']
    for k in keys: print('   %-75s %s %s' % (k, d.get(k), ''))
    st=[(float(d[k]),k) for k in hdr if k.startswith('smsp__average_warp') and 'issue_stalled' in k and k.endswith('_per_warp_active.pct') and d[k] not in ('','n/a')]
    for v,k in sorted(st, reverse=True)[:12]: print('   stall %-60s %.1f' % (k.replace('smsp__average_warps_issue_stalled_','').replace('_per_warp_active.pct',''), v))
"
