#!/bin/sh
# Build tuning variants of libfermat_b200.so (same host objects, kernels recompiled with -D overrides) into
# fermat_b200/variants/. Select one at run time with FERMAT_B200_LIB=<path>. Usage:
#   tools/build_variants.sh name1:"-DFB_REFILL_LANES=4" name2:"-DFB_SHADE_MIN_BLOCKS=6 -DFB_TRACE_MIN_BLOCKS=5" ...
set -e
cd "$(dirname "$0")/.."
make -s -j8 fermat_b200/libfermat_b200.so > /dev/null
mkdir -p fermat_b200/variants build/variants
NVCC=/usr/local/cuda/bin/nvcc
HOSTOBJ=$(ls build/*.o | grep -v -e 'build/main.o' -e 'build/pt_kernels.cu.o')
for spec in "$@"; do
  name=${spec%%:*}; flags=${spec#*:}
  $NVCC -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -fmad=false -Xcompiler -fPIC -Iinclude -Ifermat_b200/csrc/host \
        --expt-relaxed-constexpr -Xptxas -v $flags -c fermat_b200/csrc/kernels/pt_kernels.cu -o build/variants/$name.o 2> build/variants/$name.ptxas.log
  $NVCC -shared -o fermat_b200/variants/libfermat_b200_$name.so $HOSTOBJ build/variants/$name.o -cudart static -Xlinker --no-undefined -ldl -lpthread -lgomp
  echo "$name: $flags :: $(grep -A2 'k_traceILi0' build/variants/$name.ptxas.log | grep -o 'Used [0-9]* registers' | head -1) / shade $(grep -A2 'k_shade' build/variants/$name.ptxas.log | grep -o 'Used [0-9]* registers' | head -1), spills: $(grep -A1 'k_shade' build/variants/$name.ptxas.log | grep -o '[0-9]* bytes spill stores' | head -1)"
done
