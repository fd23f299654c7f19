#!/bin/bash
# usage (GPU box): scripts/stats_bench.sh [bench args] -- rebuilds the push kernels with -DIPIC_PUSH_STATS, runs the bench, prints the lane-utilisation counters, restores the build
touch implicitpic1d_b200/csrc/particles.cu
make -s -C implicitpic1d_b200/csrc EXTRA="-DIPIC_PUSH_STATS" 2>&1 | grep -E " error"
python bench.py --no-e2e --no-cpu "$@" 2>gpurun_out/stats.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('ms/step %.2f substeps/particle/sweep %.3f dep %s' % (d['ms_per_step'], d['config']['substeps_per_particle_per_sweep'], d['roofline']['per_species_ms']))"
grep "push stats" gpurun_out/stats.err
touch implicitpic1d_b200/csrc/particles.cu; make -s -C implicitpic1d_b200/csrc 2>&1 | grep -E " error"
