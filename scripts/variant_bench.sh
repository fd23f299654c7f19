#!/bin/bash
# usage (GPU box): scripts/variant_bench.sh "<bench args>" "<EXTRA flags variant 1>" "<variant 2>" ...  -- rebuilds the push kernels per variant, runs the bench
mkdir -p gpurun_out
args="$1"; shift
for v in "$@"; do
  touch implicitpic1d_b200/csrc/particles.cu
  make -s -C implicitpic1d_b200/csrc EXTRA="$v" 2>&1 | grep -E " error"
  python bench.py --no-e2e --no-cpu $args 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('variant [%s]: ms/step %.2f value %.3e substeps/particle/sweep %.2f dep %s' % (sys.argv[1], d['ms_per_step'], d['value'], d['config']['substeps_per_particle_per_sweep'], d['roofline']['per_species_ms']))" "$v" | tee -a gpurun_out/variants.log
done
touch implicitpic1d_b200/csrc/particles.cu; make -s -C implicitpic1d_b200/csrc 2>&1 | grep -E " error"
