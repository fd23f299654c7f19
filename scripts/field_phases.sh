#!/bin/bash
# usage (GPU box): scripts/field_phases.sh  -- rebuilds the library with -DIPIC_FIELD_PHASES and prints the phase times of k_field launches
# of a C5 step (and a C1 step); rebuilds the normal library afterwards
mkdir -p gpurun_out
touch implicitpic1d_b200/csrc/field.cu
make -s -C implicitpic1d_b200/csrc EXTRA=-DIPIC_FIELD_PHASES 2>&1 | grep -E "error"
python bench.py --workload c5 --steps 1 --warmup 3 --no-e2e --no-cpu 2> gpurun_out/field_phases_c5.txt > /dev/null
python bench.py --workload c1 --steps 1 --warmup 3 --no-e2e --no-cpu 2> gpurun_out/field_phases_c1.txt > /dev/null
touch implicitpic1d_b200/csrc/field.cu
make -s -C implicitpic1d_b200/csrc 2>&1 | grep -E "error"
for w in c5 c1; do echo "== $w"; grep "k_field" gpurun_out/field_phases_$w.txt | tail -6; done
