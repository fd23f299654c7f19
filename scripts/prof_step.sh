#!/bin/bash
# usage (GPU box): scripts/prof_step.sh <tag> [warm steps]  -- per-launch time/instructions of one whole implicit step at the bench's operating
# point, then `--set full` with source of a few of its push launches
tag=${1:-rX}; warm=${2:-10}
mkdir -p gpurun_out
IPIC_PROF_STEP=1 timeout 900 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__thread_inst_executed.sum,dram__bytes_read.sum \
  --clock-control none --profile-from-start off -k regex:k_push --csv --log-file gpurun_out/${tag}_step_launches.csv \
  python scripts/gpu_probe.py perf 4097 62500000 0 1.0 1 $warm > gpurun_out/${tag}_step_light.log 2>&1
tail -3 gpurun_out/${tag}_step_light.log
IPIC_PROF_STEP=1 timeout 1200 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:k_push --launch-skip ${3:-8} -c 4 \
  -o gpurun_out/${tag}_step_full -f python scripts/gpu_probe.py perf 4097 62500000 0 1.0 1 $warm > gpurun_out/${tag}_step_full.log 2>&1
python scripts/ncu_summary.py gpurun_out/${tag}_step_full.ncu-rep > gpurun_out/${tag}_step_full.txt 2>&1
ncu -i gpurun_out/${tag}_step_full.ncu-rep --page source --csv --print-source sass,cuda > gpurun_out/${tag}_step_src.csv 2>/dev/null
python scripts/ncu_lines.py gpurun_out/${tag}_step_src.csv 45 > gpurun_out/${tag}_step_lines.txt 2>&1
head -60 gpurun_out/${tag}_step_lines.txt
