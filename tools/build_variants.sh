#!/bin/bash
# builds measurement variants of the NTT translation unit: tools/build_variants.sh name "flags" [name "flags" ...]
#   -> openfhe-development_b200/libofhe_b200_<name>.so   (select with OFHE_B200_LIB=<path>)
set -e
cd "$(dirname "$0")/../openfhe-development_b200"
ARCH="-gencode arch=compute_100a,code=sm_100a"
while [ $# -ge 2 ]; do
  name=$1; flags=$2; shift 2
  mkdir -p build/v_$name
  /usr/local/cuda/bin/nvcc -O3 -std=c++17 $ARCH -lineinfo -Xcompiler -fPIC,-fopenmp $flags -c csrc/ntt.cu -o build/v_$name/ntt.o
  objs=$(ls build/*.o | grep -v '^build/ntt.o$')
  /usr/local/cuda/bin/nvcc -shared $ARCH -o libofhe_b200_$name.so build/v_$name/ntt.o $objs -Xcompiler -fopenmp -lgomp -lcudart
  echo "built libofhe_b200_$name.so ($flags)"
done
