#!/bin/bash
# A/B on one box: record-copy variants of the lanes kernel (0 = compiled out, 1 = inlined, 2 = call)
mkdir -p gpurun_out
: > gpurun_out/ab_r1g.txt
L=openpystruct_b200/lib
run() {
  [ -f $L/$1 ] || return
  OPS_B200_LIB=$PWD/$L/$1 timeout 300 python bench.py --workload $2 --beams $3 --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import sys,json
try:
    d=json.loads(sys.stdin.read()); print('$1 $2 B=$3 kernel_ms %.3f value %.0f frac %.4f e2e %.0f'%(d['roofline']['kernel_ms'], d['value'], d['roofline']['frac'], d['e2e']['value']))
except Exception as ex: print('$1 $2 FAILED',ex)
" | tee -a gpurun_out/ab_r1g.txt
}
for rep in 1 2; do
for lib in libvariant_sc0.so libvariant_sc1.so libvariant_sc2.so; do
  run $lib cfg2 10000
done
done
run libvariant_sc0.so cfg3 1000000
run libvariant_sc2.so cfg3 1000000
