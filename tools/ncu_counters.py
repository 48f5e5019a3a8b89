#!/usr/bin/env python
"""Per-kernel counters of an `ncu --set full` capture as JSON (read by bench.py for
roofline.issue_frac / lanes / traffic) and as a markdown table.
usage: ncu_counters.py full.ncu-rep out.json out.md"""
import csv, io, json, subprocess, sys
rep, out_json, out_md = sys.argv[1:4]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rr = list(csv.reader(io.StringIO(raw)))
hdr, units = rr[0], rr[1]


def val(r, key):
    i = hdr.index(key)
    v = float(r[i].replace(",", ""))
    u = units[i].lower()
    return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(u, 1)


keys = [("gpu__time_duration.sum", "duration_us"), ("launch__registers_per_thread", "registers"),
        ("launch__grid_size", "grid"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "occupancy_pct"),
        ("smsp__inst_executed.sum", "warp_instructions"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue_pct"),
        ("smsp__thread_inst_executed_per_inst_executed.ratio", "active_lanes_per_instruction"),
        ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "fma_pipe_pct"),
        ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "alu_pipe_pct"),
        ("l1tex__t_sector_hit_rate.pct", "l1_hit_pct"), ("lts__t_sector_hit_rate.pct", "l2_hit_pct"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_pct")]
stalls = [h for h in hdr if "issue_stalled" in h and h.endswith("per_issue_active.ratio") and "not_issued" not in h]
out = {"source": rep.split("/")[-1]}
md = [f"# ncu --set full counters ({rep.split('/')[-1]})", ""]
for r in rr[2:]:
    name = r[hdr.index("Kernel Name")].split("(")[0].replace("void ", "")
    k = {"k_doublets<0>": "k_doublets", "k_doublets<1>": "k_doublets_spill", "k_doublets<2>": "k_doublets_fallback", "k_doublets<3>": "k_doublets_sides",
         "k_triplets<0>": "k_triplets", "k_triplets<1>": "k_triplets_dense"}.get(name, name.split("<")[0])
    d = {lab: val(r, key) for key, lab in keys if key in hdr}
    d["dram_bytes"] = val(r, "dram__bytes_read.sum") + val(r, "dram__bytes_write.sum")
    d["dram_bytes_read"] = val(r, "dram__bytes_read.sum")
    d["dram_bytes_written"] = val(r, "dram__bytes_write.sum")
    d["stalls_per_issue"] = {h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""):
                             round(val(r, h), 3) for h in stalls}
    out[k] = d
    md += [f"## {name}", "", "| counter | value |", "|---|---|"]
    md += [f"| {lab} | {v:,.3f} |" for lab, v in d.items() if not isinstance(v, dict)]
    top = sorted(d["stalls_per_issue"].items(), key=lambda kv: -kv[1])[:7]
    md += ["| stalls per issued instruction | " + ", ".join(f"{a} {b}" for a, b in top) + " |", ""]
json.dump(out, open(out_json, "w"), indent=1)
open(out_md, "w").write("\n".join(md) + "\n")
print("\n".join(md))
