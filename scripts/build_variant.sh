#!/bin/bash
# usage (here, no GPU needed): scripts/build_variant.sh <name> "<EXTRA nvcc flags>" [source.cu ...]
# builds build/variants/<name>/libipic1d_b200.so: the named sources (default particles.cu) recompiled with the flags, the other objects taken from
# the in-tree build.  On the GPU box: IPIC1D_B200_LIB=$PWD/build/variants/<name>/libipic1d_b200.so python bench.py ...
set -e
name=$1; extra=$2; shift 2
srcs=${@:-particles.cu}
root=$(cd "$(dirname "$0")/.." && pwd)
out=$root/build/variants/$name
mkdir -p $out
make -s -C $root/implicitpic1d_b200/csrc >/dev/null 2>&1
objs=""
for o in abi field particles collisions injection diagnostics peer; do
  if echo " $srcs " | grep -q " $o.cu "; then
    nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -ccbin g++ --expt-relaxed-constexpr $extra -c -o $out/$o.o $root/implicitpic1d_b200/csrc/$o.cu 2>&1 | grep -E " error" || true
    objs="$objs $out/$o.o"
  else
    objs="$objs $root/implicitpic1d_b200/lib/$o.o"
  fi
done
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o $out/libipic1d_b200.so $objs -ldl
ls -la $out/libipic1d_b200.so
