#!/bin/bash
# first GPU pass of the shared-memory-state kernels: parity, then throughput against the lanes kernel
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "smem or other_discretisations or fine_discretisation" > gpurun_out/pytest_wide1.log 2>&1
tail -5 gpurun_out/pytest_wide1.log
: > gpurun_out/wide1.txt
run() {  # solver threads workload beams
  OPS_WIDE_THREADS=$2 OPS_LANES_THREADS=$2 timeout 300 python bench.py --solver $1 --workload $3 --beams $4 --steps 3 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import sys,json
try:
    d=json.loads(sys.stdin.read()); print('solver=$1 T=$2 $3 B=$4 kernel_ms %.3f value %.0f frac %.4f e2e %.0f'%(d['roofline']['kernel_ms'], d['value'], d['roofline']['frac'], d['e2e']['value']))
except Exception as ex: print('solver=$1 T=$2 $3 FAILED',ex)
" | tee -a gpurun_out/wide1.txt
}
run 0 320 cfg2 10000
run 3 512 cfg2 10000
run 3 448 cfg2 10000
run 3 384 cfg2 10000
run 3 320 cfg2 10000
run 3 512 cfg2 56832
run 0 320 cfg2 56832
run 0 512 cfg5 4096
run 0 256 cfg5 4096
run 2 256 cfg5 4096
