#!/bin/bash
# Round 2, first pass on one B200: issue-model micro-benchmarks, the tightened parity suite on the round-1 kernels,
# default bench line as today's baseline on this pool.
mkdir -p gpurun_out
echo "== ubench"; timeout 300 scripts/ubench/ubench 2>&1 | tee gpurun_out/r2a_ubench.txt | tail -45
echo "== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -30 | tee gpurun_out/r2a_pytest_gpu.log
echo "== bench"; timeout 900 python bench.py --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/r2a_bench.json | cut -c1-300
