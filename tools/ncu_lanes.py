#!/usr/bin/env python
"""Where a kernel's lanes are lost: executed warp instructions, average active threads and
idle lane-slots (32 * inst - thread inst) per region of the KERNEL source. Inlined helper code
(seed_math.cuh, intrinsics) is attributed to the closest preceding seed_kernels.cuh line in
SASS address order. usage: ncu_lanes.py rep kernel-regex lib.so [bucket_lines]"""
import csv, io, os, re, subprocess, sys, tempfile
rep, kern, lib = sys.argv[1], sys.argv[2], sys.argv[3]
bucket = int(sys.argv[4]) if len(sys.argv) > 4 else 1
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{kern}"],
                     capture_output=True, text=True).stdout
lines = out.splitlines()
start = next(i for i, l in enumerate(lines) if l.startswith('"Address"'))
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
        infunc = re.search(kern, l) is not None
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
agg, ctx = {}, 0
tot = tth = 0
for r in rows:
    a = int(r["Address"], 16) - base
    f, ln = addr2line.get(a) or ("?", 0)
    if f == "seed_kernels.cuh":
        ctx = ln
    n = int(r["Instructions Executed"] or 0)
    t = int(r["Thread Instructions Executed"] or 0)
    s = int(r["# Samples"] or 0)
    e = agg.setdefault(ctx // bucket * bucket, [0, 0, 0])
    e[0] += n; e[1] += t; e[2] += s
    tot += n; tth += t
src = open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "traccc_b200", "csrc", "seed_kernels.cuh")).read().splitlines()
print(f"total warp inst {tot:,}  avg threads {tth/max(tot,1):.2f}  idle lane-slots {32*tot-tth:,}")
idle_tot = 32 * tot - tth
for k in sorted(agg):
    n, t, s = agg[k]
    if n < tot * 0.004:
        continue
    text = src[k - 1].strip()[:70] if 0 < k <= len(src) else ""
    print(f"{k:5d} inst {100*n/tot:5.1f}%  thr/inst {t/max(n,1):5.1f}  idle {100*(32*n-t)/max(idle_tot,1):5.1f}%  samp {s:5d} | {text}")
