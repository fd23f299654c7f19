"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel: launches, total, mean, share.
usage: python scripts/launch_summary.py launches.csv"""
import csv, collections, re, sys
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 14 and r[0].isdigit()]
agg = collections.OrderedDict()
for r in rows:
    name = re.sub(r"\(.*", "", r[4]).replace("ipic::", "")
    name = re.sub(r"cub::CUB_\w+::", "cub::", name)
    if len(name) > 70: name = name[:67] + "..."
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1; a[1] += float(r[14])
tot = sum(a[1] for a in agg.values())
print("%-72s %8s %12s %12s %7s" % ("kernel", "launches", "total_ms", "mean_us", "share"))
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("%-72s %8d %12.3f %12.1f %6.1f%%" % (k, n, t / 1e6, t / n / 1e3, 100 * t / tot))
print("%-72s %8d %12.3f" % ("TOTAL", len(rows), tot / 1e6))
