"""Aggregate an `ncu --page source --csv --print-source sass,cuda` dump of the push kernel per code REGION of particles.cu
(regions are found from marker comments in the source, so the script follows the file as it changes).
usage: python scripts/ncu_regions.py src.csv [path/to/particles.cu]"""
import csv, collections, sys, os, re

def num(x):
    try: return float(x)
    except Exception: return 0.0

src = sys.argv[2] if len(sys.argv) > 2 else os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "implicitpic1d_b200", "csrc", "particles.cu")
lines = open(src).read().split("\n")
markers = [("deposit helpers", r"^struct DepositState"), ("transit loop (PTX)", r"^__device__ __forceinline__ double lean_sqrt"),
           ("push prologue / claim / loads", r"^template <int SCHEME, bool MOVE"), ("queue pop", r"the general sub-step: runs batches"),
           ("transit set-up / bookkeeping", r"---- transit phase"), ("finish stage", r"---- finish stage"), ("general sub-step", r"^\s+bool cont = requeue;"),
           ("requeue", r"const unsigned cm = __ballot_sync\(FULL, cont\);"),
           ("fresh path", r"fresh batches of the next sub-tile"), ("epilogue", r"epilogue: accumulators"), ("other kernels", r"^__global__ void __launch_bounds__\(PUSH_THREADS, 1\) k_push\(")]
bounds = []
for name, pat in markers:
    for i, l in enumerate(lines):
        if re.search(pat, l):
            bounds.append((i + 1, name)); break
bounds.sort()
def region(f, ln):
    if not f.endswith("particles.cu"): return "inlined headers (votes, atomics, fx)"
    r = "before"
    for b, name in bounds:
        if ln >= b: r = name
    return r

rows = list(csv.reader(open(sys.argv[1])))
sections = []; cur = None
for r in rows:
    if len(r) >= 2 and r[0] == "File Path": cur = {"file": r[1], "rows": []}; sections.append(cur)
    elif len(r) >= 2 and r[0] == "Function Name": cur["func"] = r[1]
    elif r and r[0] == "Line No": cur["hdr"] = r
    elif cur is not None and "hdr" in cur and r: cur["rows"].append(r)
inst, seen = [], set()
for s in sections:
    if s["file"] in seen or not inst: inst.append([]); seen = set()
    seen.add(s["file"]); inst[-1].append(s)
for k, secs in enumerate(inst):
    agg = collections.OrderedDict(); tot = 0.0; tots = 0.0
    for s in secs:
        h = s["hdr"]; iL = h.index("Line No"); iI = h.index("Instructions Executed"); iSamp = h.index("# Samples")
        for r in s["rows"]:
            try: ln = int(r[iL])
            except Exception: continue
            a = agg.setdefault(region(s["file"], ln), [0.0, 0.0]); a[0] += num(r[iI]); a[1] += num(r[iSamp]); tot += num(r[iI]); tots += num(r[iSamp])
    print("===== instance", k, secs[0].get("func", "?")[:60], "total warp-instr %.4g" % tot)
    for name, (ie, sm) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
        print("  %-40s instr %6.2f%%   samples %6.2f%%" % (name, 100 * ie / max(tot, 1), 100 * sm / max(tots, 1)))
