#!/usr/bin/env python
"""Print the hot SASS of one kernel from an .ncu-rep (source page): executed warp
instructions and stall samples per instruction.  usage: ncu_hot.py rep kernel-regex [min_frac]"""
import csv, io, subprocess, sys
rep, kern = sys.argv[1], sys.argv[2]
min_frac = float(sys.argv[3]) if len(sys.argv) > 3 else 0.004
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{kern}"],
                     capture_output=True, text=True).stdout
lines = out.splitlines()
start = next(i for i, l in enumerate(lines) if l.startswith('"Address"'))
end = next((i for i in range(start + 1, len(lines)) if lines[i].startswith('"Kernel Name"') or lines[i].startswith('"Address"')), len(lines))
rows = list(csv.DictReader(io.StringIO("\n".join(lines[start:end]))))
tot = sum(int(r["Instructions Executed"] or 0) for r in rows)
samp = sum(int(r["# Samples"] or 0) for r in rows)
print(f"total warp instructions {tot:,}  samples {samp}")
stalls = [k for k in rows[0] if k.startswith("stall_") and "Not Issued" not in k]
for i, r in enumerate(rows):
    n = int(r["Instructions Executed"] or 0)
    s = int(r["# Samples"] or 0)
    if n >= min_frac * tot or s >= min_frac * samp * 2:
        top = sorted(((int(r[k] or 0), k[6:]) for k in stalls), reverse=True)[:2]
        st = " ".join(f"{k}:{v}" for v, k in top if v)
        print(f"{i:5d} {n:>11,} {100*n/tot:5.1f}% thr={r['Avg. Threads Executed']:>5} smp={s:>5} {100*s/max(samp,1):4.1f}% | {r['Source'].strip()[:70]:70s} | {st}")
