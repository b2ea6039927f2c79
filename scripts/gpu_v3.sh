#!/bin/bash
# Third GPU pass: parity suite over the three solvers (incl. the 8-lanes-per-beam kernel), thread sweep, bench, ncu.
mkdir -p gpurun_out
echo "== pytest gpu" ; timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 | tee gpurun_out/pytest_gpu_v3.log
for thr in 192 256 288 320; do
  echo "== bench lanes threads=$thr"
  OPS_LANES_THREADS=$thr timeout 300 python bench.py --steps 5 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import sys,json
try:
    d=json.loads(sys.stdin.read()); print('lanes $thr value %.0f beams/s  kernel_ms %.3f  frac %.4f e2e %.0f'%(d['value'],d['roofline']['kernel_ms'],d['roofline']['frac'],d['e2e']['value']))
except Exception as ex: print('FAILED',ex)
" | tee -a gpurun_out/lanes_sweep.txt
done
echo "== bench default" ; timeout 600 python bench.py --steps 10 2>&1 | tail -1 | tee gpurun_out/bench_v3.json
echo "== ncu launches" ; timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_v3.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launch_v3.log 2>&1 ; tail -2 gpurun_out/ncu_launch_v3.log
echo "== ncu full" ; timeout 900 ncu --set full --clock-control none --import-source on -k regex:beamopt_lanes_kernel -s 3 -c 1 -f -o gpurun_out/prof_v3_lanes python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full_v3.log 2>&1 ; tail -2 gpurun_out/ncu_full_v3.log
ls -la gpurun_out
