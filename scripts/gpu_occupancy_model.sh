#!/bin/bash
# Steady-state iteration time L(w) of the lanes kernel against resident warps per scheduler (profiling only):
# B = 148 SMs x (T/8 beams per CTA) x 6 full rounds, so no partially filled round distorts the fit.
mkdir -p gpurun_out
: > gpurun_out/occupancy_model.txt
for T in 64 128 192 256 288 320; do
  B=$((148 * T / 8 * 6))
  OPS_LANES_THREADS=$T timeout 300 python bench.py --workload cfg2 --beams $B --steps 3 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import sys,json
try:
    d=json.loads(sys.stdin.read()); ms=d['roofline']['kernel_ms']
    print('T=$T B=$B kernel_ms %.3f  us_per_round_epoch %.3f  value %.0f'%(ms, ms*1e3/600/6, d['value']))
except Exception as ex: print('T=$T FAILED',ex)
" | tee -a gpurun_out/occupancy_model.txt
done
