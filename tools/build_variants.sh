#!/bin/bash
# Builds tuning variants of libraingun_b200.so into raingun_b200/_variants/<name>.so (git-ignored; they travel
# with gpurun).  Usage: tools/build_variants.sh name "-DFLAG=1 ..." [name flags ...]; select one with RAINGUN_B200_LIB.
set -e
cd "$(dirname "$0")/../raingun_b200/csrc"
mkdir -p ../_variants
while [ $# -ge 2 ]; do
  name=$1; flags=$2; shift 2
  d=_build/var_$name; mkdir -p $d
  for f in rg_api rg_wavefront rg_grid rg_multi; do
    [ -f $f.cu ] || continue
    nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -fmad=false --expt-relaxed-constexpr \
      -Xcompiler -fPIC,-ffp-contract=off,-Wall -Xptxas -v $flags -c $f.cu -o $d/$f.o 2> $d/$f.ptxas.log &
  done
  wait
  nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../_variants/$name.so $d/*.o -lcudart
  echo "built $name ($flags): $(grep -A3 'k_trace_gridILb0ELb0' $d/rg_wavefront.ptxas.log | grep -o 'Used [0-9]* registers' | head -1), $(grep -A1 'k_trace_gridILb0ELb0' $d/rg_wavefront.ptxas.log | grep -o '[0-9]* bytes spill stores' | head -1)"
done
