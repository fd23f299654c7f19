#!/bin/bash
# usage (GPU box): scripts/gpu_configs.sh <tag>  -- bench lines of the other BASELINE configs (C1, C2, C3) next to the C5 line
tag=${1:-rX}
mkdir -p gpurun_out
for w in c1 c2 c4; do
  timeout 900 python bench.py --workload $w --steps 5 --warmup 3 --cpu-seconds 8 > gpurun_out/${tag}_bench_${w}.json 2> gpurun_out/${tag}_bench_${w}.err
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/${tag}_bench_${w}.json").read().strip().splitlines()[-1])
    print("$w", "%.1f ms/step" % d["ms_per_step"], "%.3e" % d["value"], d["roofline"]["per_species_ms"], "frac %.3f" % d["roofline"]["frac"], d["config"]["residual_evaluations_per_step"], "cpu %.3e" % d["cpu_baseline"]["value"])
except Exception as ex:
    print("$w failed", ex)
PY
done
