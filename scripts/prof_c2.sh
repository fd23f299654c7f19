#!/bin/bash
# usage (GPU box): scripts/prof_c2.sh <tag>  -- ncu --set full (with source) of four CIC push launches of the C2 workload in steady state
tag=${1:-rX}
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_push --launch-skip 60 -c 4 -o gpurun_out/${tag}_c2_full -f \
  python bench.py --workload c2 --steps 10 --warmup 3 --no-e2e --no-cpu > gpurun_out/${tag}_c2_full.log 2>&1
python scripts/ncu_summary.py gpurun_out/${tag}_c2_full.ncu-rep > gpurun_out/${tag}_c2_full.txt 2>&1
ncu -i gpurun_out/${tag}_c2_full.ncu-rep --page source --csv --print-source sass,cuda > gpurun_out/${tag}_c2_src.csv 2>/dev/null
python scripts/ncu_lines.py gpurun_out/${tag}_c2_src.csv 40 > gpurun_out/${tag}_c2_lines.txt 2>&1
head -34 gpurun_out/${tag}_c2_full.txt; head -45 gpurun_out/${tag}_c2_lines.txt
