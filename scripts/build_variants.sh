#!/bin/bash
# Builds variants of libmadrona_b200 with extra nvcc -D flags for one-box A/B runs:
#   scripts/build_variants.sh name1 "-DX=1" name2 "-DY=2 -DZ=3" ...
# -> madrona_b200/libmadrona_b200_<name>.so (git-ignored)
set -e
cd "$(dirname "$0")/../madrona_b200"
KERNEL_OBJS="build/kernels_physics.o build/kernels_sort.o build/kernels_render.o"
while [ $# -gt 1 ]; do
  name=$1; flags=$2; shift 2
  rm -f $KERNEL_OBJS
  make -j8 NVCCFLAGS_EXTRA="$flags" > /dev/null
  cp libmadrona_b200.so libmadrona_b200_$name.so
done
rm -f $KERNEL_OBJS
make -j8 > /dev/null
ls -la *.so
