#!/bin/bash
# usage (on the GPU box): scripts/gpu_round.sh <tag>   -- GPU tests, bench line, ncu launch list, ncu --set full of the push kernel
tag=${1:-rX}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/${tag}_smi.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/${tag}_pytest_gpu.log
tail -3 gpurun_out/${tag}_pytest_gpu.log
timeout 900 python bench.py > gpurun_out/${tag}_bench_c5.json 2> gpurun_out/${tag}_bench_c5.err; tail -c 3000 gpurun_out/${tag}_bench_c5.json
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${tag}_bench_c5_reference_arm.json 2>/dev/null
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/${tag}_launches.csv \
   python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu > gpurun_out/${tag}_ncu_bench.log 2>&1
python scripts/launch_summary.py gpurun_out/${tag}_launches.csv > gpurun_out/${tag}_launches_c5.txt; cat gpurun_out/${tag}_launches_c5.txt
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:k_push -o gpurun_out/${tag}_push_full -f \
   python scripts/gpu_probe.py prof 4097 62500000 0 1.0 1 3 > gpurun_out/${tag}_ncu_full.log 2>&1
python scripts/ncu_summary.py gpurun_out/${tag}_push_full.ncu-rep > gpurun_out/${tag}_push_ncu_full.txt 2>&1; cat gpurun_out/${tag}_push_ncu_full.txt
# between-step particle operations at the C5 shard's size: CUDA-event times against the measured HBM peak
timeout 300 python scripts/aux_ops_bench.py > gpurun_out/${tag}_aux_ops_c5.json 2> gpurun_out/${tag}_aux_ops.err; tail -12 gpurun_out/${tag}_aux_ops.err
