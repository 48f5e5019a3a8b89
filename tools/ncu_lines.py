#!/usr/bin/env python
"""Executed warp instructions and stall samples of one kernel per SOURCE LINE: joins the SASS
page of an .ncu-rep with `nvdisasm -g` line info of the same build (by instruction address).
usage: ncu_lines.py rep kernel-regex lib.so [top_n [mangled-regex [table-index]]]
(template instantiations share a demangled name: give the mangled section regex, e.g.
k_doubletsILb0, and the index of the launch among those the kernel regex matches)"""
import csv, io, os, re, subprocess, sys, tempfile
rep, kern, lib = sys.argv[1], sys.argv[2], sys.argv[3]
top_n = int(sys.argv[4]) if len(sys.argv) > 4 else 45
mangled = sys.argv[5] if len(sys.argv) > 5 else kern
table = int(sys.argv[6]) if len(sys.argv) > 6 else 0
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{kern}"],
                     capture_output=True, text=True).stdout
lines = out.splitlines()
start = [i for i, l in enumerate(lines) if l.startswith('"Address"')][table]
end = next((i for i in range(start + 1, len(lines)) if lines[i].startswith('"Kernel Name"') or lines[i].startswith('"Address"')), len(lines))
rows = list(csv.DictReader(io.StringIO("\n".join(lines[start:end]))))
base = int(rows[0]["Address"], 16)
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=tmp, capture_output=True)
cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout
addr2line, cur, infunc = {}, None, False
for l in dis.splitlines():
    if l.startswith(".text."):
        infunc = re.search(mangled, l) is not None
        continue
    if not infunc:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/", l)
    if m:
        addr2line[int(m.group(1), 16)] = cur
agg, tot, samp = {}, 0, 0
for r in rows:
    a = int(r["Address"], 16) - base
    n, s = int(r["Instructions Executed"] or 0), int(r["# Samples"] or 0)
    key = addr2line.get(a, ("?", 0))
    e = agg.setdefault(key, [0, 0, 0])
    e[0] += n; e[1] += s; e[2] += 1
    tot += n; samp += s
src = {}
print(f"total warp instructions {tot:,} samples {samp}")
for key, (n, s, k) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top_n]:
    f, ln = key
    text = ""
    path = os.path.join(os.path.dirname(os.path.abspath(lib)), "csrc", f)
    if os.path.exists(path):
        if path not in src:
            src[path] = open(path).read().splitlines()
        if 0 < ln <= len(src[path]):
            text = src[path][ln - 1].strip()[:80]
    print(f"{f}:{ln:<5d} inst {100*n/tot:5.1f}%  samples {100*s/max(samp,1):5.1f}%  sass {k:4d} | {text}")
