#!/bin/bash
# Build a compile-time variant of the library for an A/B inside one gpurun call:
#   tools/build_variant.sh NAME "-DDZO_HYBRID_WARPS=2" batched_hybrid_tu_a.cu [more .cu files]
# recompiles the named translation units with the extra flags and links them with the regular objects into
# dzoptimization.jl_b200/csrc/variants/libdzopt_NAME.so (select it with DZOPT_B200_LIB=<path>).
set -e
cd "$(dirname "$0")/.."
name=$1; flags=$2; shift 2
C=dzoptimization.jl_b200/csrc
mkdir -p $C/variants/$name
objs=""
for f in $C/build/*.o; do
  b=$(basename $f .o)
  skip=0; for s in "$@"; do [ "$b.cu" == "$s" ] && skip=1; done
  [ $skip == 0 ] && objs="$objs $f"
done
for s in "$@"; do
  /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -fmad=false -Xcompiler -fPIC -Xcompiler -O2 $flags -c $C/$s -o $C/variants/$name/$(basename $s .cu).o
  objs="$objs $C/variants/$name/$(basename $s .cu).o"
done
/usr/local/cuda/bin/nvcc -shared -o $C/variants/libdzopt_$name.so $objs -gencode arch=compute_100a,code=sm_100a -ldl
echo $C/variants/libdzopt_$name.so
