"""List the SASS of one kernel instance of an `ncu --page source --csv --print-source sass,cuda` dump in address order with
executed counts and source line.  usage: python scripts/ncu_sass.py src.csv [instance] [min_share_pct]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
inst_want = int(sys.argv[2]) if len(sys.argv) > 2 else 0
sections = []; cur = None
for r in rows:
    if len(r) >= 2 and r[0] == "File Path": cur = {"file": r[1], "rows": []}; sections.append(cur)
    elif r and r[0] == "Line No": cur["hdr"] = r
    elif cur is not None and "hdr" in cur and r: cur["rows"].append(r)
inst, seen = [], set()
for s in sections:
    if s["file"] in seen or not inst: inst.append([]); seen = set()
    seen.add(s["file"]); inst[-1].append(s)
out = []
for s in inst[inst_want]:
    h = s["hdr"]; iL = 0; iA = h.index("Address"); iS = iA + 1; iI = h.index("Instructions Executed"); iT = h.index("Thread Instructions Executed")
    for r in s["rows"]:
        try: out.append((int(r[iA], 16), r[iS], float(r[iI] or 0), float(r[iT] or 0), s["file"].split("/")[-1], r[iL]))
        except Exception: pass
out.sort()
tot = sum(o[2] for o in out)
print("total warp-instr %.4g" % tot)
for a, sass, ie, te, f, ln in out:
    if ie > 0: print("%05x %6.2f%% thr %4.1f  %-70s %s:%s" % (a & 0xfffff, 100 * ie / tot, te / max(ie, 1), sass[:70], f, ln))
