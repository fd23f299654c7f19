#!/bin/bash
# usage (GPU box): scripts/ab_bench.sh [bench args]  -- the same bench line with the round-1 library (scripts/ab/libipic1d_b200_r1.so) and the in-tree build
fmt='import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print(sys.argv[1], "ms/step %.2f value %.4e iters %.2f substeps/particle/sweep %.3f dep %s" % (d["ms_per_step"], d["value"], d["config"]["residual_evaluations_per_step"], d["config"]["substeps_per_particle_per_sweep"], d["roofline"]["per_species_ms"]))
print("   ", {k: round(v, 2) for k, v in d["breakdown_ms_per_step"].items()})'
# (a library built from an older commit lacks the newer entry points and fails to load: that arm is skipped then)
if [ -f scripts/ab/libipic1d_b200_r1.so ]; then IPIC1D_B200_LIB=$PWD/scripts/ab/libipic1d_b200_r1.so python bench.py --no-e2e --no-cpu "$@" 2>/dev/null | python -c "$fmt" r1 2>/dev/null; fi
python bench.py --no-e2e --no-cpu "$@" 2>/dev/null | python -c "$fmt" new
