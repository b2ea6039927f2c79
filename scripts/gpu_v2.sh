#!/bin/bash
# Second GPU pass: both solvers through the parity suite, layout/occupancy sweep of the three-moment kernel, ncu.
mkdir -p gpurun_out
echo "== pytest gpu" ; timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 | tee gpurun_out/pytest_gpu_v2.log
for lay in smem mv all; do
  for thr in 64 128 192 256; do
    echo "== bench flex layout=$lay threads=$thr"
    OPS_FLEX_LAYOUT=$lay OPS_FLEX_THREADS=$thr timeout 300 python bench.py --steps 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import sys,json
try:
    d=json.loads(sys.stdin.read()); print('$lay $thr value %.0f beams/s  kernel_ms %.2f  frac %.4f e2e %.0f'%(d['value'],d['roofline']['kernel_ms'],d['roofline']['frac'],d['e2e']['value']))
except Exception as ex: print('FAILED',ex)
" | tee -a gpurun_out/flex_sweep.txt
  done
done
echo "== bench default" ; timeout 600 python bench.py --steps 5 2>&1 | tail -1 | tee gpurun_out/bench_v2.json
echo "== ncu full (default layout)" ; timeout 900 ncu --set full --clock-control none --import-source on -k regex:beamopt_flex_kernel -s 3 -c 1 -f -o gpurun_out/prof_r1_flex python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full_flex.log 2>&1 ; tail -2 gpurun_out/ncu_full_flex.log
ls -la gpurun_out
