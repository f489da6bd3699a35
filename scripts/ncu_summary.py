"""Summarise an .ncu-rep (raw page) into the handful of metrics DESIGN.md / profiles/ cite."""
import csv, subprocess, sys, json
rep = sys.argv[1]
pats = sys.argv[2:] or ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct", "launch__registers_per_thread",
    "launch__occupancy", "sm__warps_active.avg.pct", "sm__throughput.avg.pct", "smsp__inst_executed.sum ", "sm__inst_executed_pipe", "sm__pipe",
    "lts__t_bytes.sum ", "lts__t_sector_hit_rate", "l1tex__data_bank_conflicts_pipe_lsu", "smsp__issue_active.avg.pct", "smsp__average_warp", "smsp__warp_issue_stalled",
    "smsp__inst_executed_op", "sm__cycles_elapsed.avg ", "smsp__cycles_active.avg ", "l1tex__data_pipe_lsu_wavefronts_mem_shared", "smsp__sass_inst_executed_op_shared",
    "launch__shared_mem", "launch__grid_size", "launch__waves"]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[0]; units = rows[1]
for r in rows[2:]:
    print("##", r[hdr.index("Kernel Name")], "grid", r[hdr.index("Grid Size")], "block", r[hdr.index("Block Size")])
    for h, u, v in zip(hdr, units, r):
        if any(p in h + " " for p in pats):
            print(f"{h:100s} {u:14s} {v}")
