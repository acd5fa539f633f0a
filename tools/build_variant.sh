#!/bin/bash
# Builds an A/B variant of libb200vit.so with extra nvcc defines, next to the product library (git-ignored, travels to the GPU box):
#   bash tools/build_variant.sh scalar_gelu -DB200_GELU_SCALAR      ->  vit-is-all-you-need_b200/b200vit/libb200vit_scalar_gelu.so
# Use it with  B200VIT_LIB=$PWD/vit-is-all-you-need_b200/b200vit/libb200vit_<name>.so python ...
set -e
name=$1; shift
root=$(cd "$(dirname "$0")/.." && pwd)
src=$root/vit-is-all-you-need_b200/csrc
out=/tmp/b200vit_variant_$name
mkdir -p $out
for f in $src/*.cu; do
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC --expt-relaxed-constexpr "$@" -c $f -o $out/$(basename $f .cu).o &
done
wait
nvcc -shared -Wno-deprecated-gpu-targets -o $root/vit-is-all-you-need_b200/b200vit/libb200vit_$name.so $out/*.o
echo $root/vit-is-all-you-need_b200/b200vit/libb200vit_$name.so
