#!/bin/bash
# Session pass: parity suite on the trimmed lanes kernel, then A/B against the previous build (variants are profiling-only libs)
mkdir -p gpurun_out
echo "== pytest gpu" ; timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 | tee gpurun_out/pytest_gpu_r1b.log
: > gpurun_out/ab_r1b.txt
L=openpystruct_b200/lib
run() {  # lib threads workload beams
  [ -f $L/$1 ] || return
  OPS_B200_LIB=$PWD/$L/$1 OPS_LANES_THREADS=$2 timeout 300 python bench.py --workload $3 --beams $4 --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import sys,json
try:
    d=json.loads(sys.stdin.read()); print('$1 T=$2 $3 B=$4 kernel_ms %.3f value %.0f frac %.4f e2e %.0f es %.0f'%(d['roofline']['kernel_ms'], d['value'], d['roofline']['frac'], d['e2e']['value'], d.get('early_stop_mode',{}).get('value',0)))
except Exception as ex: print('$1 T=$2 $3 FAILED',ex)
" | tee -a gpurun_out/ab_r1b.txt
}
for lib in ${LIBS:-libvariant_base.so libopenpystruct_b200.so}; do
  run $lib 320 cfg2 10000
  run $lib 320 cfg2 35520
done
