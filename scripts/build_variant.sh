#!/bin/bash
# A/B builds of the library: scripts/build_variant.sh NAME "<extra nvcc flags>" [TU ...]
# recompiles the named translation units (default: the BLS12-381 curve TU) with the extra flags and links
# variants/NAME.so from them plus the objects of the regular build.  Use with BLAZE_B200_LIB=variants/NAME.so.
set -e
cd "$(dirname "$0")/.."
NAME=$1; FLAGS=$2; shift 2 || true
TUS=${@:-msm_curve_bls12_381}
make -s -j8 -C blaze_b200/csrc
mkdir -p build/var/$NAME variants
OBJS=""
SRCS=$(sed -n 's/^SRCS := //p' blaze_b200/csrc/Makefile | sed 's/\$(EXTRA_SRCS)//; s/\.cu//g')
for b in $SRCS; do
  o=build/obj/$b.o
  skip=0
  for t in $TUS; do [ "$t" = "$b" ] && skip=1; done
  [ $skip = 0 ] && OBJS="$OBJS $o"
done
for t in $TUS; do
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC $FLAGS -Xptxas -v \
    -c blaze_b200/csrc/$t.cu -o build/var/$NAME/$t.o 2> build/var/$NAME/$t.ptxas.log &
done
wait
for t in $TUS; do OBJS="$OBJS build/var/$NAME/$t.o"; done
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o variants/$NAME.so $OBJS -lcudart_static -ldl -lrt -lpthread
echo built variants/$NAME.so
