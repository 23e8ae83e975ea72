"""short per-kernel digest of an ncu --set full report:  python tools/ncu_brief.py file.ncu-rep"""
import csv, subprocess, sys
out = subprocess.run(['ncu', '-i', sys.argv[1], '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines())); h = rows[0]
want = ["gpu__time_duration.sum", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
        "sm__inst_executed_pipe_lsu.sum", "sm__inst_executed_pipe_fma.sum", "sm__inst_executed_pipe_alu.sum", "sm__inst_executed_pipe_uniform.sum",
        "sm__inst_executed_pipe_xu.sum", "sm__inst_executed_pipe_cbu.sum", "sm__inst_executed_pipe_adu.sum",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "smsp__warps_eligible.avg.per_cycle_active", "smsp__warps_active.avg.per_cycle_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed"]
st = [c for c in h if "issue_stalled" in c and "per_issue_active" in c and "not_issued" not in c]
for r in rows[2:]:
    print("--", r[h.index("Kernel Name")][:60])
    for w in want:
        if w in h:
            print(f"  {w:75s}{r[h.index(w)]}")
    print("  stalls:", ", ".join("%s=%.2f" % (c.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""), float(r[h.index(c)]))
                                 for c in st if float(r[h.index(c)]) > 0.1))
