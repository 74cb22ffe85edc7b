#!/usr/bin/env python
"""Summarise an ncu report: one row per profiled launch with the metrics that matter for HBM/L2-bound kernels.
usage: python tools/ncu_table.py report.ncu-rep [--json out.json]"""
import csv, json, re, subprocess, sys
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
idx = {h: i for i, h in enumerate(hdr)}
want = [("gpu__time_duration.sum", "us"), ("dram__bytes_read.sum", "dramR"), ("dram__bytes_write.sum", "dramW"),
        ("lts__t_sectors.sum", "l2sect"), ("lts__t_sector_hit_rate.pct", "l2hit%"), ("l1tex__t_sector_hit_rate.pct", "l1hit%"),
        ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram%"), ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2%"),
        ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "l1%"), ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm%"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue%"), ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ%"),
        ("launch__registers_per_thread", "regs"), ("smsp__inst_executed.sum", "inst"), ("launch__grid_size", "grid"), ("launch__block_size", "blk")]
want = [(m, n) for m, n in want if m in idx]
print("%-34s" % "kernel" + "".join("%10s" % n for _, n in want))
print("%-34s" % "" + "".join("%10s" % units[idx[m]][:9] for m, _ in want))
agg = {}
for r in rows[2:]:
    name = re.sub(r"^(?:\w+::)+", "", r[idx["Kernel Name"]].split("(")[0].replace("void ", ""))
    vals = []
    for m, n in want:
        v = r[idx[m]].replace(",", "")
        try:
            vals.append(float(v))
        except ValueError:
            vals.append(float("nan"))
    print("%-34s" % name[:34] + "".join("%10.4g" % v for v in vals))
    a = agg.setdefault(name, {"n": 0, "us": 0.0, "dram": 0.0})
    a["n"] += 1
    d = dict(zip([n for _, n in want], vals))
    def tobytes(metric, v):
        u = units[idx[metric]].lower()
        return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(u, 1)
    a["us"] += d.get("us", 0)
    a["dram"] += tobytes("dram__bytes_read.sum", d.get("dramR", 0)) + tobytes("dram__bytes_write.sum", d.get("dramW", 0))
if "--json" in sys.argv:
    p = sys.argv[sys.argv.index("--json") + 1]
    json.dump({k: {"launches_profiled": v["n"], "us_per_launch_under_ncu": v["us"] / v["n"],
                   "dram_bytes_per_launch": v["dram"] / v["n"]} for k, v in agg.items()}, open(p, "w"), indent=1)
