#!/bin/bash
# 384-thread instance for many-round batches; ke widening on the FP64 pipe (variant)
mkdir -p gpurun_out
: > gpurun_out/ab_r1d.txt
L=openpystruct_b200/lib
run() {  # lib threads workload beams
  [ -f $L/$1 ] || return
  OPS_B200_LIB=$PWD/$L/$1 OPS_LANES_THREADS=$2 timeout 300 python bench.py --workload $3 --beams $4 --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import sys,json
try:
    d=json.loads(sys.stdin.read()); print('$1 T=$2 $3 B=$4 kernel_ms %.3f value %.0f frac %.4f e2e %.0f'%(d['roofline']['kernel_ms'], d['value'], d['roofline']['frac'], d['e2e']['value']))
except Exception as ex: print('$1 T=$2 $3 FAILED',ex)
" | tee -a gpurun_out/ab_r1d.txt
}
run libopenpystruct_b200.so 999 cfg2 10000
run libvariant_kem.so 999 cfg2 10000
run libopenpystruct_b200.so 999 cfg2 10000
run libvariant_kem.so 999 cfg2 10000
run libopenpystruct_b200.so 999 cfg3 1000000
run libopenpystruct_b200.so 320 cfg3 1000000
run libvariant_kem.so 999 cfg3 1000000
OPS_B200_LIB=$PWD/$L/libvariant_kem.so timeout 900 python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -3 | tee gpurun_out/pytest_kem.log
