#!/bin/bash
# A/B of library variants on one box: LIBS="a.so b.so" REPS=n
mkdir -p gpurun_out
L=openpystruct_b200/lib
for rep in $(seq 1 ${REPS:-3}); do
for lib in $LIBS; do
  [ -f $L/$lib ] || continue
  OPS_B200_LIB=$PWD/$L/$lib timeout 300 python bench.py --workload ${WL:-cfg2} --beams ${BEAMS:-10000} --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import sys,json
try:
    d=json.loads(sys.stdin.read()); print('$lib kernel_ms %.3f value %.0f frac %.4f e2e %.0f'%(d['roofline']['kernel_ms'], d['value'], d['roofline']['frac'], d['e2e']['value']))
except Exception as ex: print('$lib FAILED',ex)
" | tee -a gpurun_out/ab2.txt
done
done
