#!/bin/bash
# A/B of library builds on the same GPU: tools/k4_ab.sh A B ...  (variants/lib_<name>.so; "new" = the in-tree library)
for v in "$@"; do
  echo "== $v"
  if [ "$v" = new ]; then L=""; else L="MOSAIC_B200_LIB=variants/lib_$v.so"; fi
  env $L timeout 300 python tools/k4_time.py 2>&1 | tail -3
done
