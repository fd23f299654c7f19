#!/bin/bash
# usage (GPU box): scripts/pair_sweep.sh  -- paired vs per-species depositJ launches and the split of warps
mkdir -p gpurun_out
run() {
  python bench.py --steps 6 --warmup 3 --no-e2e --no-cpu "$@" 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('%-44s ms/step %.1f value %.3e dep %s frac %.3f' % (' '.join(sys.argv[1:]), d['ms_per_step'], d['value'], d['roofline']['per_species_ms'], d['roofline']['frac']))" "$@" | tee -a gpurun_out/pair_sweep.log
}
run --option pair_species 0
run
run --option pair_warps 2
run --option pair_warps 6
run --option pair_warps 8
run --option pair_warps 12
