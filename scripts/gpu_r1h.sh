#!/bin/bash
# single GPU: parity suite on the build with separate scatter instances, A/B against the scatter-free variant
mkdir -p gpurun_out
echo "== pytest gpu" ; timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/pytest_gpu_r1h.log
: > gpurun_out/ab_r1h.txt
L=openpystruct_b200/lib
run() {
  [ -f $L/$1 ] || return
  OPS_B200_LIB=$PWD/$L/$1 timeout 300 python bench.py --workload $2 --beams $3 --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import sys,json
try:
    d=json.loads(sys.stdin.read()); print('$1 $2 B=$3 kernel_ms %.3f value %.0f frac %.4f e2e %.0f'%(d['roofline']['kernel_ms'], d['value'], d['roofline']['frac'], d['e2e']['value']))
except Exception as ex: print('$1 $2 FAILED',ex)
" | tee -a gpurun_out/ab_r1h.txt
}
for rep in 1 2; do
for lib in libvariant_sc0.so libopenpystruct_b200.so; do
  run $lib cfg2 10000
done
done
