#!/bin/bash
# usage: scripts/option_sweep.sh  (runs bench.py --steps 6 --warmup 3 for a few kernel-option settings; prints ms/step)
mkdir -p gpurun_out
run() {
  python bench.py --steps 6 --warmup 3 --no-e2e --no-cpu "$@" 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('%-40s ms/step %.1f value %.3e e %.3f i %.3f' % (' '.join(sys.argv[1:]), d['ms_per_step'], d['value'], d['roofline']['per_species_ms']['electrons'], d['roofline']['per_species_ms']['ions']))" "$@" | tee -a gpurun_out/option_sweep.log
}
run
run --option transit_min 6
run --option transit_min 9
run --option transit_min 16
run --option transit_min 20
run --option stay_min 14
run --option stay_min 26
run --option transit_min 8 --option stay_min 16
