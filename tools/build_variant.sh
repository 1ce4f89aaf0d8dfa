#!/bin/bash
# tools/build_variant.sh NAME "-DMK_X=.. ..." [SOURCE] : variants/lib_NAME.so = the in-tree library with SOURCE (default mk_strip)
# rebuilt under extra flags (A/B runs on one GPU box: MOSAIC_B200_LIB=variants/lib_NAME.so, tools/k4_ab.sh NAME new)
set -e
cd "$(dirname "$0")/../mosaic-library_b200/csrc"
mkdir -p ../../variants
SRC=${3:-mk_strip}
ARCH="-gencode arch=compute_100a,code=sm_100a"
nvcc -O3 -std=c++17 -lineinfo $ARCH -Xcompiler -fPIC,-ffp-contract=off,-fvisibility=hidden -Xptxas -v $2 -c $SRC.cu -o /tmp/${SRC}_$1.o 2> /tmp/${SRC}_$1.log
grep -E "Used|spill" /tmp/${SRC}_$1.log | sort | uniq -c | sort -rn | head -6
OBJS=""
for o in mk_api mk_mapper mk_strip mk_preproc mk_ransac mk_knn mk_knn_tc; do
  if [ $o = $SRC ]; then OBJS="$OBJS /tmp/${SRC}_$1.o"; else OBJS="$OBJS $o.o"; fi
done
nvcc -shared $ARCH -o ../../variants/lib_$1.so $OBJS -lcudart_static -lcuda -ldl -lrt -lpthread
