"""Print the handful of ncu metrics that decide a kernel's next step: python tools/ncu_brief.py report.ncu-rep"""
import csv
import subprocess
import sys

WANT = ["gpu__time_duration.sum", "sm__inst_executed_pipe_fp64.avg.pct", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct", "sm__inst_executed_pipe_fma.avg.pct", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct", "sm__warps_active.avg.pct", "smsp__pcsamp_warps_issue_stalled", "launch__occupancy_limit",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "smsp__inst_executed.sum", "dram__bytes_read.sum",
        "dram__bytes_write.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared", "smsp__warps_eligible.avg",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__t_sector_hit_rate",
        "lts__t_sector_hit_rate", "sass__inst_executed_shared", "l1tex__data_pipe_lsu_wavefronts_mem_shared_op",
        "SM_A.TriageCompute.l1tex__data_pipe_lsu_wavefronts", "dram__throughput.avg.pct", "lts__t_bytes.sum.per_second",
        "sm__throughput.avg.pct", "gpu__dram_throughput.avg.pct"]

out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[0]
for r in rows[2:]:
    print("==", r[hdr.index("Kernel Name")])
    for h, u, v in zip(hdr, rows[1], r):
        if any(h.startswith(w) for w in WANT) and "not_issued" not in h and v not in ("0", ""):
            print(f"  {h:84s} {u:10s} {v}")
