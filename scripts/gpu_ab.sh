#!/bin/bash
# A/B of library variants on one box: LIBS="a.so b.so" [COUNTS="5920 10000 23680"] (variants built with
# `python -m openpystruct_b200.build --out openpystruct_b200/lib/a.so -D...`, loaded through OPS_B200_LIB)
mkdir -p gpurun_out
L=$PWD/openpystruct_b200/lib
for lib in $LIBS; do
  [ -f $L/$lib ] || continue
  echo "== $lib"; OPS_B200_LIB=$L/$lib timeout 300 python scripts/sweep_beams.py ${COUNTS:-5920 10000 23680} 2>&1 | grep "^B=" | tee gpurun_out/ab_${lib%.so}.txt
done
