"""Turn an .ncu-rep (ncu --set full) into the short text summary kept under profiles/ (selected raw metrics per kernel).
usage: python tools/ncu_summary.py report.ncu-rep > profiles/xxx.txt"""
import csv, subprocess, sys

KEEP = ("gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "launch__grid_size", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "sm__cycles_active.avg", "sm__cycles_active.max",
        "sm__cycles_elapsed.avg", "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum")
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
for r in rows[2:]:
    d = dict(zip(hdr, r))
    u = dict(zip(hdr, units))
    print(d.get("Kernel Name"))
    for k in hdr:
        if k in KEEP or (k.startswith("smsp__average_warps_issue_stalled") and k.endswith("per_issue_active.ratio")):
            print(f"  {k:100s} {d[k]:>16s} {u[k]}")
    print()
