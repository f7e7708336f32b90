"""Aggregate the ncu source page (SASS view) by opcode class and by address range:
    python tools/ncu_regions.py report.ncu-rep [kernel-index]
Shows, per opcode, executed warp instructions, stall samples and shared-memory wavefronts."""
import csv
import subprocess
import sys
from collections import defaultdict

out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
# the page holds one block per kernel: a "Kernel Name" line, a header line, then instructions
blocks, cur = [], None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "hdr": None, "rows": []}
        blocks.append(cur)
    elif cur is not None and cur["hdr"] is None:
        cur["hdr"] = r
    elif cur is not None:
        cur["rows"].append(r)
which = int(sys.argv[2]) if len(sys.argv) > 2 else 0
blk = blocks[which]
h = blk["hdr"]
iS, iE, iN, iW = h.index("Source"), h.index("Instructions Executed"), h.index("# Samples"), h.index("L1 Wavefronts Shared")
print(blk["name"])
by = defaultdict(lambda: [0, 0, 0])
tot = [0, 0, 0]
lines = []
for k, r in enumerate(blk["rows"]):
    op = r[iS].split()
    op = op[1] if op and op[0].startswith("@") else (op[0] if op else "?")
    op = op.rstrip(";")
    e, n, w = int(r[iE] or 0), int(r[iN] or 0), int(r[iW] or 0)
    for acc in (by[op], tot):
        acc[0] += e; acc[1] += n; acc[2] += w
    lines.append((k, r[iS].strip(), e, n, w))
print(f"total: executed {tot[0]}  samples {tot[1]}  smem wavefronts {tot[2]}")
for op, (e, n, w) in sorted(by.items(), key=lambda kv: -kv[1][0])[:40]:
    print(f"  {op:28s} exec {e:12d} ({100 * e / tot[0]:5.1f}%)  samples {n:7d} ({100 * n / max(1, tot[1]):5.1f}%)  wavefronts {w:10d}")
if "--lines" in sys.argv:
    for k, s, e, n, w in lines:
        print(f"{k:5d} {e:10d} {n:6d} {w:9d}  {s}")
