#!/bin/bash
# usage (GPU box): scripts/sort_sweep.sh -- bench line as a function of the cell-sort period (electrons) and the ion factor
mkdir -p gpurun_out
for kk in "1 8" "2 8" "2 1" "3 8" "5 1" "5 8"; do
  set -- $kk
  python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu --sort-every $1 --sort-ion-factor $2 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('sort-every %s ion-factor %s: ms/step %.1f value %.3e dep %s sort %.2f' % (sys.argv[1], sys.argv[2], d['ms_per_step'], d['value'], d['roofline']['per_species_ms'], d['breakdown_ms_per_step']['sort']))" $1 $2 | tee -a gpurun_out/sort_sweep.log
done
