#!/bin/bash
# usage: scripts/variant_probe.sh "<EXTRA flags variant 1>" "<variant 2>" ...   (runs on the GPU box: rebuilds the push kernels per variant)
mkdir -p gpurun_out
for v in "$@"; do
  touch implicitpic1d_b200/csrc/particles.cu
  make -s -C implicitpic1d_b200/csrc EXTRA="$v" 2>&1 | grep -E "error" 
  echo "=== variant: $v" | tee -a gpurun_out/variants.log
  timeout 300 python scripts/gpu_probe.py perf 4097 62500000 0 1.0 1 3 2>&1 | grep -E "sorted|solvePotential" | tee -a gpurun_out/variants.log
done
