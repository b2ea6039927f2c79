#!/bin/bash
# Quick perf probe: lanes kernel thread sweep + ncu full capture (no pytest).
mkdir -p gpurun_out
for thr in ${SWEEP:-256 320}; do
  OPS_LANES_THREADS=$thr timeout 300 python bench.py --steps 5 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import sys,json
try:
    d=json.loads(sys.stdin.read()); print('lanes $thr value %.0f beams/s  kernel_ms %.3f  frac %.4f e2e %.0f'%(d['value'],d['roofline']['kernel_ms'],d['roofline']['frac'],d['e2e']['value']))
except Exception as ex: print('FAILED',ex)
" | tee -a gpurun_out/lanes_sweep_quick.txt
done
if [ -n "$NCU" ]; then
timeout 900 ncu --set full --clock-control none --import-source on -k regex:beamopt_lanes_kernel -s 3 -c 1 -f -o gpurun_out/prof_quick python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_quick.log 2>&1 ; tail -2 gpurun_out/ncu_quick.log
fi
