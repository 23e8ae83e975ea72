#!/bin/bash
# usage: tools/ncu_summary.sh <file.ncu-rep>  -- the metrics DESIGN.md / profiles/ quote, one block per captured launch
ncu -i "$1" --page raw --csv 2>/dev/null | python3 -c '
import csv,sys
rows=list(csv.reader(sys.stdin))
h=rows[0]; units=rows[1]
want=["Kernel Name","gpu__time_duration.sum","dram__bytes_read.sum","dram__bytes_write.sum","gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
"sm__throughput.avg.pct_of_peak_sustained_elapsed","launch__registers_per_thread","launch__occupancy_limit_shared_mem","launch__occupancy_limit_registers",
"sm__warps_active.avg.pct_of_peak_sustained_active","smsp__inst_executed.sum","smsp__issue_active.avg.pct_of_peak_sustained_active",
"l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum","l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum","l1tex__t_sector_hit_rate.pct","lts__t_sectors_op_read.sum","lts__t_sector_hit_rate.pct",
"lts__t_sectors_srcunit_tex_op_read.sum","l1tex__data_pipe_lsu_wavefronts.sum","l1tex__data_pipe_lsu_wavefronts_mem_shared.sum","l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
"l1tex__throughput.avg.pct_of_peak_sustained_elapsed","lts__throughput.avg.pct_of_peak_sustained_elapsed","l1tex__lsu_writeback_active.avg.pct_of_peak_sustained_elapsed","l1tex__f_wavefronts.sum","l1tex__t_output_wavefronts_pipe_lsu_mem_global_op_ld.sum"]
stall=[c for c in h if "issue_stalled" in c and "per_issue_active" in c and "not_issued" not in c]
for r in rows[2:]:
    print("-----")
    for w in want+stall:
        if w in h:
            i=h.index(w); print(f"{w:86s}{r[i]} {units[i]}")
'
