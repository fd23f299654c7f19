#!/bin/bash
# usage (GPU box): scripts/sanitize.sh <tag>  -- compute-sanitizer memcheck over the GPU parity tests (without the full-size, swarm, long-loop cases) and
# racecheck over the kernels that changed in round 2 (transit loop / finish stage in depositJ + moveParticles, fused Anderson iteration, MCC, refluxing)
tag=${1:-rX}; out=gpurun_out/${tag}_sanitizer.txt; mkdir -p gpurun_out
echo "# compute-sanitizer on the GPU tests (B200), $tag" > $out
echo "## memcheck: tests/test_gpu_parity.py tests/test_gpu_periodic.py tests/test_gpu_mcc.py tests/test_gpu_injection.py tests/test_gpu_density.py tests/test_gpu_simulation.py (without the full-size, swarm, 2-GPU, 3000-call cases)" >> $out
timeout 1500 compute-sanitizer --tool memcheck --print-limit 5 python -m pytest tests/test_gpu_parity.py tests/test_gpu_periodic.py tests/test_gpu_mcc.py tests/test_gpu_injection.py tests/test_gpu_density.py tests/test_gpu_simulation.py \
  -m gpu -q -k "not full_size and not large_load and not swarm and not two_gpu and not timing_state and not collisional_time_loop" 2>&1 | grep -E "passed|failed|ERROR SUMMARY|Invalid|error" | tail -6 >> $out
echo "## racecheck: depositJ[ngp_rf], moveParticles[cic_rf], solvePotential_pcr_mode[ngp_rf], anderson chunk [ngp_rf], MCC two sets, refluxing after a domain-crossing step" >> $out
timeout 1500 compute-sanitizer --tool racecheck --print-limit 5 python -m pytest tests/test_gpu_parity.py tests/test_gpu_mcc.py tests/test_gpu_injection.py -m gpu -q \
  -k "(test_depositJ and ngp_rf) or (test_moveParticles and cic_rf) or (pcr_mode and ngp_rf) or (anderson_decision and ngp_rf) or two_sets or (domain_crossing and ngp)" 2>&1 | grep -E "passed|failed|RACECHECK SUMMARY|hazard" | tail -6 >> $out
cat $out
