"""Turn the ncu captures of tools/profile_r1.sh (gpurun_out/) into the tracked summaries under profiles/."""
import csv
import json
import shutil
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
OUT = ROOT / "profiles"
GP = ROOT / "gpurun_out"
sys.path.insert(0, str(ROOT / "tools"))

KEEP = ["dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "gpu__time_duration.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "launch__block_size", "launch__grid_size",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "sm__cycles_elapsed.max.per_second", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__warps_eligible.avg.per_cycle_active"]


def raw(rep):
    out = subprocess.run(["ncu", "-i", str(rep), "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    return rows[0], rows[1], rows[2:]


def main():
    global OUT
    tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
    if len(sys.argv) > 2:     # on the GPU box: summaries next to the captures (gpurun_out/ travels back, at most 64 MiB)
        OUT = Path(sys.argv[2])
    OUT.mkdir(exist_ok=True)
    lines = [f"# ncu --set full --clock-control none summaries, round {tag[1:]} (B200, sm_100a). Source: tools/profile_{tag[0]}{int(tag[1:])}.sh + tools/profile_summarise.py", ""]
    traffic = None
    for name in sorted(p.stem for p in GP.glob(f"{tag}_*.ncu-rep")):
        rep = GP / f"{name}.ncu-rep"
        if not rep.exists():
            continue
        hdr, units, rows = raw(rep)
        for r in rows:
            kn = r[hdr.index("Kernel Name")]
            lines.append(f"== {kn}   [{name}.ncu-rep]")
            vals = dict(zip(hdr, r))
            un = dict(zip(hdr, units))
            for k in KEEP:
                if k in vals and vals[k] != "":
                    lines.append(f"{k:86s}{un[k]:16s}{vals[k]}")
            for k in sorted(vals):
                if k.startswith("smsp__pcsamp_warps_issue_stalled_") and "not_issued" not in k and vals[k] not in ("", "0"):
                    lines.append(f"{k:86s}{un[k]:16s}{vals[k]}")
            lines.append("")
            if "pair_tile_kernel" in kn and traffic is None:
                def to_bytes(key):
                    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}[un[key]]
                    return float(vals[key]) * scale
                traffic = {"kernel": kn, "dram_bytes_read": to_bytes("dram__bytes_read.sum"), "dram_bytes_write": to_bytes("dram__bytes_write.sum")}
                traffic["dram_bytes_per_launch"] = traffic["dram_bytes_read"] + traffic["dram_bytes_write"]
                traffic["source"] = "ncu --set full, one launch of the bench.py workload (65536 x 1024), tools/profile_" + tag[0] + str(int(tag[1:])) + ".sh"
    (OUT / f"{tag}_ncu_summary.txt").write_text("\n".join(lines) + "\n")
    if traffic:
        (OUT / f"{tag}_pair_tile_traffic.json").write_text(json.dumps(traffic, indent=1) + "\n")
    if (GP / f"{tag}_launches.csv").exists():
        shutil.copy(GP / f"{tag}_launches.csv", OUT / f"{tag}_launches.csv")
    print("wrote", sorted(p.name for p in OUT.glob(f"{tag}_*")))


if __name__ == "__main__":
    main()
