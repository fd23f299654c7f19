#!/bin/bash
# usage (GPU box): scripts/prof_field.sh <tag>  -- ncu --set full (with source) of six k_field launches in the middle of a C5 step
tag=${1:-rX}
mkdir -p gpurun_out
IPIC_PROF_STEP=1 timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:k_field --launch-skip 30 -c 6 \
  -o gpurun_out/${tag}_field_full -f python scripts/gpu_probe.py perf 4097 62500000 0 1.0 1 6 > gpurun_out/${tag}_field_full.log 2>&1
python scripts/ncu_summary.py gpurun_out/${tag}_field_full.ncu-rep > gpurun_out/${tag}_field_full.txt 2>&1
ncu -i gpurun_out/${tag}_field_full.ncu-rep --page source --csv --print-source sass,cuda > gpurun_out/${tag}_field_src.csv 2>/dev/null
python scripts/ncu_lines.py gpurun_out/${tag}_field_src.csv 30 > gpurun_out/${tag}_field_lines.txt 2>&1
grep -E "Kernel Name|gpu__time_duration|inst_executed.sum|issue_active|top stalls" gpurun_out/${tag}_field_full.txt | head -40
head -100 gpurun_out/${tag}_field_lines.txt
