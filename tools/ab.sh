#!/bin/bash
# A/B of library builds on the same GPU: tools/ab.sh SCRIPT A B ...  (variants/lib_<name>.so; "new" = the in-tree library)
S=$1; shift
for v in "$@"; do
  echo "== $v"
  if [ "$v" = new ]; then L=""; else L="MOSAIC_B200_LIB=variants/lib_$v.so"; fi
  env $L timeout 600 python $S 2>&1 | tail -12
done
