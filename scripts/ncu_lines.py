"""Aggregate an `ncu --page source --csv --print-source sass,cuda` dump per source line.
usage: ncu -i rep.ncu-rep --page source --csv --print-source sass,cuda > src.csv ; python scripts/ncu_lines.py src.csv [top]"""
import csv, collections, sys

def num(x):
    try: return float(x)
    except Exception: return 0.0

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
sections = []; cur = None
for r in rows:
    if len(r) >= 2 and r[0] == "File Path": cur = {"file": r[1], "rows": []}; sections.append(cur)
    elif len(r) >= 2 and r[0] == "Function Name": cur["func"] = r[1]
    elif r and r[0] == "Line No": cur["hdr"] = r
    elif cur is not None and "hdr" in cur and r: cur["rows"].append(r)
# group sections into kernel instances: a new instance starts when a file name repeats
inst, seen = [], set()
for s in sections:
    if s["file"] in seen or not inst: inst.append([]); seen = set()
    seen.add(s["file"]); inst[-1].append(s)
for k, secs in enumerate(inst):
    tot = 0; per = []
    for s in secs:
        h = s["hdr"]; iL = h.index("Line No"); iI = h.index("Instructions Executed"); iSamp = h.index("# Samples")
        agg = collections.OrderedDict()
        for r in s["rows"]:
            try: ln = int(r[iL])
            except Exception: continue
            a = agg.setdefault(ln, [r[1], 0.0, 0.0]); a[1] += num(r[iI]); a[2] += num(r[iSamp])
        for ln, (src, ie, sm) in agg.items():
            per.append((ie, sm, s["file"].split('/')[-1], ln, src)); tot += ie
    print("===== instance", k, secs[0].get("func", "?")[:60], "total warp-instr %.4g" % tot)
    per.sort(reverse=True)
    for ie, sm, f, ln, src in per[:top]:
        print("%6.2f%% samp %7d  %s:%d  %s" % (100 * ie / max(tot, 1), sm, f, ln, src.strip()[:120]))
