#!/bin/bash
# usage (GPU box): scripts/ab2.sh [bench args] -- the bench step with build/variants/base (a build of an earlier commit, made with scripts/build_variant.sh) and the in-tree build
fmt='import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print("%-6s ms/step %.2f value %.4e iters %.2f sps %.3f e %.4f i %.4f" % (sys.argv[1], d["ms_per_step"], d["value"], d["config"]["residual_evaluations_per_step"], d["config"]["substeps_per_particle_per_sweep"], d["roofline"]["per_species_ms"]["electrons"], d["roofline"]["per_species_ms"]["ions"]), {k: round(v, 2) for k, v in d["breakdown_ms_per_step"].items()})'
for v in ${VARIANTS:-base}; do
  IPIC1D_B200_LIB=$PWD/build/variants/$v/libipic1d_b200.so python bench.py --no-e2e --no-cpu "$@" 2>/dev/null | python -c "$fmt" $v
done
python bench.py --no-e2e --no-cpu "$@" 2>/dev/null | python -c "$fmt" new
