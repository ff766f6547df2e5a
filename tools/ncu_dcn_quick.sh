#!/bin/bash
# quick ncu metric pass over dcn3d_kernel (one launch): wavefronts, hit rates, issue utilisation
M=gpu__time_duration.sum,l1tex__data_pipe_lsu_wavefronts.sum,l1tex__data_pipe_lsu_wavefronts_mem_lgds.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,l1tex__data_pipe_lsu_wavefronts.sum.pct_of_peak_sustained_elapsed,l1tex__t_sector_hit_rate.pct,l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum,l1tex__t_sectors_pipe_lsu_mem_global_op_ld_lookup_miss.sum,lts__t_sector_hit_rate.pct,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__cycles_elapsed.max,l1tex__m_xbar2l1tex_read_bytes.sum,l1tex__lsu_writeback_active.avg.pct_of_peak_sustained_elapsed
for env in "$@"; do
  echo "== $env"
  env $env ncu --clock-control none --metrics $M -k regex:dcn3d_kernel -s 2 -c 1 --csv python tools/bench_dcn.py --iters 2 2>/dev/null | grep -v "^==" | python -c "
import csv,sys
for r in csv.reader(sys.stdin):
    if len(r)>10 and r[0]!='ID': print('   ', r[-3], r[-2], r[-1])
"
done
