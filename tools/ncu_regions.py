#!/usr/bin/env python
"""Executed warp instructions / stall samples of one kernel from an .ncu-rep, summed over
windows of W consecutive SASS instructions.  usage: ncu_regions.py rep kernel-regex [W]"""
import csv, io, subprocess, sys
rep, kern = sys.argv[1], sys.argv[2]
W = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{kern}"],
                     capture_output=True, text=True).stdout
lines = out.splitlines()
start = next(i for i, l in enumerate(lines) if l.startswith('"Address"'))
end = next((i for i in range(start + 1, len(lines)) if lines[i].startswith('"Kernel Name"') or lines[i].startswith('"Address"')), len(lines))
rows = list(csv.DictReader(io.StringIO("\n".join(lines[start:end]))))
tot = sum(int(r["Instructions Executed"] or 0) for r in rows)
samp = sum(int(r["# Samples"] or 0) for r in rows)
print(f"total warp instructions {tot:,}  samples {samp}  sass instructions {len(rows)}")
for a in range(0, len(rows), W):
    blk = rows[a:a + W]
    n = sum(int(r["Instructions Executed"] or 0) for r in blk)
    s = sum(int(r["# Samples"] or 0) for r in blk)
    ops = {}
    for r in blk:
        op = r["Source"].strip().split()[0] if not r["Source"].strip().startswith("@") else r["Source"].strip().split()[1]
        ops[op.split(".")[0]] = ops.get(op.split(".")[0], 0) + int(r["Instructions Executed"] or 0)
    top = " ".join(f"{k}:{100*v/max(n,1):.0f}%" for k, v in sorted(ops.items(), key=lambda kv: -kv[1])[:5])
    print(f"{a:5d}-{a+len(blk)-1:5d} inst {100*n/tot:5.1f}%  samples {100*s/max(samp,1):5.1f}%  {top}")
