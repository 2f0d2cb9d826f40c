#!/bin/bash
# Build an A/B variant of libfme_b200.so into variants/libfme_<name>.so (git-ignored, travels with gpurun).
#   tools/build_variant.sh swar8 "-DFME_K2_SWAR8"
#   FME_B200_LIB=variants/libfme_swar8.so python bench.py
set -e
name=$1; shift
extra="$*"
root=$(cd "$(dirname "$0")/.." && pwd)
src=$root/hm16.9-nn_fme_b200/csrc
out=$root/variants; obj=$root/variants/obj_$name
mkdir -p "$obj"
arch="-gencode arch=compute_100a,code=sm_100a"
for f in fme_capi k1_interp k2_refine k2_umma k3_nn k_misc; do
  nvcc $arch -O3 -std=c++17 -lineinfo -Xcompiler -fPIC $extra -c $src/$f.cu -o $obj/$f.o &
done
wait
nvcc $arch -shared -cudart static -o $out/libfme_$name.so $obj/*.o
echo built $out/libfme_$name.so
