#!/bin/bash
# usage (GPU box): scripts/gpu_quick.sh <tag>  -- GPU tests + the bench line (no profiler)
tag=${1:-rX}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/${tag}_pytest_gpu.log
tail -5 gpurun_out/${tag}_pytest_gpu.log
timeout 900 python bench.py > gpurun_out/${tag}_bench_c5.json 2> gpurun_out/${tag}_bench_c5.err; tail -c 3000 gpurun_out/${tag}_bench_c5.json
