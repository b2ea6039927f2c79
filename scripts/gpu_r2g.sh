#!/bin/bash
# full parity suite + default bench line + cfg3/cfg4 on the fused packed-pass build
mkdir -p gpurun_out
echo "== pytest gpu"; timeout 1800 python -m pytest tests -m gpu -q 2>&1 | tail -15 | tee gpurun_out/r2g_pytest_gpu.log
echo "== bench"; timeout 900 python bench.py 2>&1 | tail -1 | tee gpurun_out/r2g_bench.json | cut -c1-400
for wl in cfg3 cfg4; do
  echo "== bench $wl" ; timeout 900 python bench.py --workload $wl --steps 3 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/r2g_bench_$wl.json | cut -c1-300
done
