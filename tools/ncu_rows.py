#!/usr/bin/env python
"""One line per profiled launch from an exported `ncu --page raw --csv` file.  usage: python tools/ncu_rows.py raw.csv [--json out.json]"""
import csv, json, re, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
idx = {h: i for i, h in enumerate(hdr)}
want = [("gpu__time_duration.sum", "us"), ("dram__bytes_read.sum", "dramR"), ("dram__bytes_write.sum", "dramW"),
        ("lts__t_sectors.sum", "l2sect"), ("lts__t_sector_hit_rate.pct", "l2hit%"), ("l1tex__t_sector_hit_rate.pct", "l1hit%"),
        ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram%"), ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2%"),
        ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "l1%"), ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm%"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue%"), ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ%"),
        ("launch__registers_per_thread", "regs"), ("smsp__inst_executed.sum", "inst"), ("launch__grid_size", "grid"), ("launch__block_size", "blk")]
want = [(m, n) for m, n in want if m in idx]
print("%-36s" % "kernel" + "".join("%10s" % n for _, n in want))
print("%-36s" % "" + "".join("%10s" % units[idx[m]][:9] for m, _ in want))
agg = {}
def tobytes(metric, v):
    return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(units[idx[metric]].lower(), 1)
for r in rows[2:]:
    # "void rss::point_alone::meanfield_point_kernel<5, 4, 6, 3>(..." -> "meanfield_point_kernel<5, 4, 6, 3>"
    name = re.sub(r"^(?:\w+::)+", "", r[idx["Kernel Name"]].split("(")[0].replace("void ", ""))
    vals = []
    for m, n in want:
        try:
            vals.append(float(r[idx[m]].replace(",", "")))
        except ValueError:
            vals.append(float("nan"))
    print("%-36s" % name[:36] + "".join("%10.4g" % v for v in vals))
    d = dict(zip([n for _, n in want], vals))
    a = agg.setdefault(name, {"n": 0, "us": 0.0, "dram": 0.0})
    a["n"] += 1
    a["us"] += d.get("us", 0)
    a["dram"] += tobytes("dram__bytes_read.sum", d.get("dramR", 0)) + tobytes("dram__bytes_write.sum", d.get("dramW", 0))
if "--json" in sys.argv:
    p = sys.argv[sys.argv.index("--json") + 1]
    json.dump({k: {"launches_profiled": v["n"], "us_per_launch_under_ncu": v["us"] / v["n"],
                   "dram_bytes_per_launch": v["dram"] / v["n"]} for k, v in agg.items()}, open(p, "w"), indent=1)
