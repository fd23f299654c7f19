"""Per-launch table of an `ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__thread_inst_executed.sum --csv` log of one implicit step.
usage: python scripts/step_launches.py launches.csv [max rows]"""
import csv, sys
from collections import OrderedDict
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 14 and r[0].isdigit()]
d = OrderedDict()
for r in rows:
    d.setdefault(r[0], {})[r[12]] = float(r[14].replace(',', ''))
mx = int(sys.argv[2]) if len(sys.argv) > 2 else 10 ** 9
tot = {0: 0.0, 1: 0.0}
for i, (k, v) in enumerate(d.items()):
    t = v['gpu__time_duration.sum'] / 1e6; inst = v['smsp__inst_executed.sum']; th = v['smsp__thread_inst_executed.sum']
    tot[i % 2] += t
    if i < mx: print(i, 'e' if i % 2 == 0 else 'i', '%.3f ms inst %.3e lanes %.1f' % (t, inst, th / inst))
print(len(d), 'launches; electrons %.1f ms, ions %.1f ms' % (tot[0], tot[1]))
