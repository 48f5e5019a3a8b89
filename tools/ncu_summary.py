#!/usr/bin/env python
"""Markdown summary of an ncu capture for profiles/: launch list (gpu__time_duration of every
kernel of one event) + the --set full counters of the search kernels + per-source-line tables.
usage: ncu_summary.py launches.csv full.ncu-rep lib.so out.md traffic.json"""
import csv, io, json, subprocess, sys
launches, rep, lib, out_md, out_json = sys.argv[1:6]
rows = [r for r in csv.reader(open(launches)) if len(r) > 10 and r[0].isdigit() and "b200seed::" in r[4]]
ev = rows[len(rows) // 2:]            # the second (warm) event of --profile-one
md = ["# ncu summary (one 10k-particle event, second pass of `bench.py --profile-one`)", "",
      "## Launch list (`ncu --metrics gpu__time_duration.sum --clock-control none`; cold-cache, serialised: compare shares)", "",
      "| kernel | grid | block | device time (us) | share |", "|---|---|---|---|---|"]
tot = sum(float(r[-1]) for r in ev) / 1e3
for r in ev:
    name = r[4].split("(")[0].replace("void ", "").replace("b200seed::", "")
    md.append(f"| {name} | {r[8]} | {r[7]} | {float(r[-1])/1e3:.1f} | {100*float(r[-1])/1e3/tot:.1f}% |")
md += [f"| total | | | {tot:.1f} | |", ""]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rr = list(csv.reader(io.StringIO(raw)))
hdr, units = rr[0], rr[1]
want = [("gpu__time_duration.sum", "duration"), ("launch__registers_per_thread", "registers/thread"),
        ("launch__occupancy_limit_registers", "occupancy limit (registers), CTAs/SM"),
        ("launch__occupancy_limit_shared_mem", "occupancy limit (shared memory), CTAs/SM"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy"),
        ("smsp__inst_executed.sum", "warp instructions executed"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slot utilisation"),
        ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "FMA pipe (FP32 mul/add) utilisation"),
        ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "ALU pipe utilisation"),
        ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "XU pipe (rcp/rsqrt) utilisation"),
        ("smsp__thread_inst_executed_per_inst_executed.ratio", "active threads per warp instruction (warp execution efficiency x32)"),
        ("dram__bytes_read.sum", "DRAM bytes read"), ("dram__bytes_write.sum", "DRAM bytes written"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput"),
        ("l1tex__t_sector_hit_rate.pct", "L1 hit rate"), ("lts__t_sector_hit_rate.pct", "L2 hit rate")]
traffic = {}
md += ["## `ncu --set full` counters of the two search kernels", ""]
for r in rr[2:]:
    kname = r[hdr.index("Kernel Name")].split("(")[0]
    md += [f"### {kname}", "", "| counter | value |", "|---|---|"]
    for key, label in want:
        if key in hdr:
            i = hdr.index(key)
            md.append(f"| {label} (`{key}`) | {r[i]} {units[i]} |")
    def val(key):
        i = hdr.index(key)
        v = float(r[i].replace(",", ""))
        u = units[i].lower()
        return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(u, 1)
    # template instantiations: k_doublets<0> is the common kernel, k_doublets<1> the spill pass
    tkey = ("k_" + kname.split("k_")[-1]).replace("<0>", "")
    tkey = tkey.replace("<1>", "_spill" if "doublets" in tkey else "_dense")
    traffic[tkey] = val("dram__bytes_read.sum") + val("dram__bytes_write.sum")
    md.append("")
    if "<1>" in kname and "doublets" in kname:   # the spill pass: an empty launch for ordinary events
        continue
    base = kname.replace("void ", "").split("<")[0]
    mangled = base + ("ILb0" if "<0>" in kname else "ILb1" if "<1>" in kname else "")
    lines = subprocess.run([sys.executable, __file__.replace("ncu_summary", "ncu_lines"), rep, base, lib, "22",
                            mangled, "1" if "<1>" in kname else "0"],
                           capture_output=True, text=True).stdout
    md += ["Executed warp instructions / stall samples per source line (top 22):", "", "```", lines.rstrip(), "```", ""]
open(out_md, "w").write("\n".join(md) + "\n")
json.dump({"source": rep.split("/")[-1], "dram_bytes_per_launch": traffic}, open(out_json, "w"), indent=1)
print(open(out_md).read()[:3000])
