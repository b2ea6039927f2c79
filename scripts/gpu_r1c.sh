#!/bin/bash
# A/B of batch-size / CTA-size variants of the trimmed lanes kernel + one ncu full capture of the current build
mkdir -p gpurun_out
: > gpurun_out/ab_r1c.txt
L=openpystruct_b200/lib
run() {  # lib threads workload beams
  [ -f $L/$1 ] || return
  OPS_B200_LIB=$PWD/$L/$1 OPS_LANES_THREADS=$2 timeout 300 python bench.py --workload $3 --beams $4 --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import sys,json
try:
    d=json.loads(sys.stdin.read()); print('$1 T=$2 $3 B=$4 kernel_ms %.3f value %.0f frac %.4f e2e %.0f'%(d['roofline']['kernel_ms'], d['value'], d['roofline']['frac'], d['e2e']['value']))
except Exception as ex: print('$1 T=$2 $3 FAILED',ex)
" | tee -a gpurun_out/ab_r1c.txt
}
run libopenpystruct_b200.so 320 cfg2 10000
run libvariant_nb4.so 320 cfg2 10000
run libvariant_nb7.so 320 cfg2 10000
run libvariant_t352.so 352 cfg2 10000
run libvariant_t384.so 384 cfg2 10000
run libvariant_t384.so 384 cfg2 42624
run libopenpystruct_b200.so 320 cfg2 35520
timeout 900 ncu --set full --clock-control none --import-source on -k regex:beamopt_lanes_kernel -s 3 -c 1 -f -o gpurun_out/prof_r1c python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_r1c.log 2>&1 ; tail -2 gpurun_out/ncu_r1c.log
