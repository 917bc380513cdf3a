#!/bin/bash
# Build variants of the phase-sum kernel with different compile-time knobs into build/var/lib_<name>.so
# (run on the build box; then `PB200_LIB=build/var/lib_<name>.so python tools/perf_skyvis.py ...` on the GPU).
# usage: tools/variants.sh name1 "-DPB_STAGGER=2" name2 "-DPB_ABLATE=1" ...
set -e
cd "$(dirname "$0")/.."
mkdir -p build/var
ARCH="-gencode arch=compute_100a,code=sm_100a"
OTHERS=$(ls build/obj/*.o | grep -v skyvis.o)
while [ $# -ge 2 ]; do
  name=$1; flags=$2; shift 2
  nvcc $ARCH -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -Xptxas -v $flags -c prisim_b200/csrc/skyvis.cu -o build/var/skyvis_$name.o 2> build/var/skyvis_$name.ptxas.log
  nvcc $ARCH -shared -o build/var/lib_$name.so build/var/skyvis_$name.o $OTHERS -lcudart
  echo "$name: $(grep -A2 'k_skyvisILi2ELb1ELb0' build/var/skyvis_$name.ptxas.log | grep -o '[0-9]* bytes spill stores' )"
done
