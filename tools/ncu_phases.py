#!/usr/bin/env python
"""Executed warp instructions / stall samples of one kernel per PHASE of its body: every SASS
instruction is attributed to the outermost source line inside the kernel (nvdisasm -gi inline
chains), and the lines are summed over the ranges given.
usage: ncu_phases.py rep kernel-regex lib.so mangled-regex file:first-last[:name] ...
e.g.   ncu_phases.py p.ncu-rep k_doublets lib.so k_doubletsILi0 seed_kernels.cuh:674-733:setup ..."""
import csv, io, os, re, subprocess, sys, tempfile
rep, kern, lib, mangled = sys.argv[1:5]
phases = []
for spec in sys.argv[5:]:
    f, rng, *name = spec.split(":")
    a, b = (int(x) for x in rng.split("-"))
    phases.append((f, a, b, name[0] if name else f"{a}-{b}"))
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{kern}"],
                     capture_output=True, text=True).stdout
lines = out.splitlines()
start = [i for i, l in enumerate(lines) if l.startswith('"Address"')][int(os.environ.get("TABLE", 0))]
end = next((i for i in range(start + 1, len(lines)) if lines[i].startswith('"Kernel Name"') or lines[i].startswith('"Address"')), len(lines))
rows = list(csv.DictReader(io.StringIO("\n".join(lines[start:end]))))
base = int(rows[0]["Address"], 16)
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=tmp, capture_output=True)
cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "-gi", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout
addr2chain, chain, infunc, fresh = {}, [], False, True
for l in dis.splitlines():
    if l.startswith(".text."):
        infunc = re.search(mangled, l) is not None
        continue
    if not infunc:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)(?: inlined at "([^"]+)", line (\d+))?', l)
    if m:
        if fresh:
            chain, fresh = [], False
        chain.append((os.path.basename(m.group(1)), int(m.group(2))))
        if m.group(3):
            chain.append((os.path.basename(m.group(3)), int(m.group(4))))
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/", l)
    if m:
        addr2chain[int(m.group(1), 16)] = list(chain)
        fresh = True
agg = {p[3]: [0, 0, 0] for p in phases}
agg["(other)"] = [0, 0, 0]
tot = samp = 0
stall_cols = [c for c in rows[0] if c.startswith("stall_")]
for r in rows:
    a = int(r["Address"], 16) - base
    n, s = int(r["Instructions Executed"] or 0), int(r["# Samples"] or 0)
    tn = int(r.get("Thread Instructions Executed") or 0)
    tot += n
    samp += s
    ch = addr2chain.get(a, [])
    name = "(other)"
    for f, ln in reversed(ch):      # outermost first
        hit = next((p[3] for p in phases if p[0] == f and p[1] <= ln <= p[2]), None)
        if hit:
            name = hit
            break
    agg[name][0] += n
    agg[name][1] += s
    agg[name][2] += tn
print(f"total warp instructions {tot:,}  samples {samp}")
for k, (n, s, tn) in agg.items():
    print(f"{k:28s} inst {100*n/max(tot,1):5.1f}%  samples {100*s/max(samp,1):5.1f}%  lanes {tn/max(n,1):5.1f}")
