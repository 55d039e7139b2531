#!/bin/bash
# Build A/B variants of libnufhe_b200.so into tools/variants/ (git-ignored, shipped to the GPU box by gpurun), one per
# argument; an argument is a name followed by nvcc -D flags, e.g.
#     tools/build_variants.sh "base" "asm64 -DNB_ASM64=1" "ct4 -DNB_BR_CT=4"
# then:  gpurun -- 'bash tools/variant_experiment.sh'   (times every variant on the same seeded workload and prints
# the output checksums: equal checksums = same bits).  tools/sass_stats.py --lib tools/variants/<name>.so gives the
# static instruction mix of each build first.
set -e
cd "$(dirname "$0")/../nufhe_b200/csrc"
mkdir -p ../../tools/variants
for spec in "$@"; do
  set -- $spec
  name=$1; shift
  nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 --shared -Xcompiler -fPIC "$@" \
       -o ../../tools/variants/$name.so capi.cu 2>&1 | grep -i " error" || true
  echo "$name: $(python ../../tools/sass_stats.py --lib ../../tools/variants/$name.so | grep 'per step')"
done
