"""Summarises an `ncu --metrics gpu__time_duration.sum --csv` launch list: per kernel launches, total, mean, share.
    python profiles/summarize_launches.py gpurun_out/launches_rNN.csv > profiles/rNN_launches_summary.txt
Per-launch times under ncu are cold-cache and serialised: compare SHARES with bench.py's live numbers, not absolutes."""
import collections
import csv
import sys

lines = [l for l in open(sys.argv[1]) if not l.startswith("==")]
acc = collections.OrderedDict()
tot = 0.0
for row in csv.DictReader(lines):
    v = float(row["Metric Value"].replace(",", ""))
    v = {"ns": v / 1e3, "us": v, "ms": v * 1e3, "s": v * 1e6}.get(row["Metric Unit"], v)
    a = acc.setdefault(row["Kernel Name"], [0, 0.0])
    a[0] += 1
    a[1] += v
    tot += v
print("# %s: %d launches, %.1f us total" % (sys.argv[1], sum(a[0] for a in acc.values()), tot))
print("%-64s %6s %11s %10s %7s" % ("kernel", "n", "total_us", "avg_us", "share"))
for n, (c, t) in sorted(acc.items(), key=lambda kv: -kv[1][1]):
    print("%-64s %6d %11.1f %10.2f %6.1f%%" % (n[:64], c, t, t / c, 100 * t / tot))
