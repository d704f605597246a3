#!/usr/bin/env python
"""Text summary of one ncu report for profiles/: key raw metrics + SASS-region sample shares.
    python tools/ncu_summary.py gpurun_out/prof.ncu-rep > profiles/rNN_ncu_<kernel>_<workload>.txt"""
import csv, io, subprocess, sys
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
KEEP = ("dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct", "gpu__time_duration.sum",
        "launch__block_size", "launch__grid_size", "launch__occupancy_limit", "launch__registers_per_thread",
        "launch__shared_mem_per_block", "sm__cycles_elapsed.avg", "sm__inst_executed.sum.per_cycle", "sm__inst_executed_pipe_alu.sum.pct",
        "sm__inst_executed_pipe_fma.sum.pct", "sm__inst_executed_pipe_fp64.sum.pct", "sm__inst_executed_pipe_lsu.sum.pct",
        "sm__inst_executed_pipe_xu.sum.pct", "sm__pipe_tensor_cycles_active.avg.pct", "sm__pipe_fma_cycles_active.avg.pct",
        "sm__throughput.avg.pct", "sm__warps_active.avg.pct", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct",
        "smsp__average_warps_issue_stalled", "lts__throughput.avg.pct", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum")
name = vals[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "?"
print(f"--- {name} grid {vals[hdr.index('Grid Size')] if 'Grid Size' in hdr else ''}  (ncu --set full --clock-control none; {rep})")
for h, u, v in sorted(zip(hdr, units, vals)):
    if any(h.startswith(k) for k in KEEP) and v not in ("", "0"):
        print(f"  {h} [{u}] = {v}")
