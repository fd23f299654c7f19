#!/bin/bash
# usage (GPU box): scripts/opt_sweep2.sh "<bench args>" "<options 1>" "<options 2>" ...  -- the bench step per option set (in-tree build)
fmt='import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print("%-44s ms/step %.2f value %.4e e %.4f i %.4f" % (sys.argv[1], d["ms_per_step"], d["value"], d["roofline"]["per_species_ms"]["electrons"], d["roofline"]["per_species_ms"]["ions"]))'
args="$1"; shift
for o in "$@"; do
  python bench.py --no-e2e --no-cpu $args $o 2>/dev/null | python -c "$fmt" "$o"
done
