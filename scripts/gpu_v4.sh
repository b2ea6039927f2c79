#!/bin/bash
# Fourth GPU pass: parity (incl. shared-I load cases, session), default bench, the other BASELINE configs.
mkdir -p gpurun_out
echo "== pytest gpu" ; timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/pytest_gpu_v4.log
echo "== bench cfg2" ; timeout 600 python bench.py --steps 10 2>&1 | tail -1 | tee gpurun_out/bench_v4_cfg2.json | cut -c1-600
echo "== bench cfg4" ; timeout 600 python bench.py --workload cfg4 --steps 3 2>&1 | tail -1 | tee gpurun_out/bench_v4_cfg4.json | cut -c1-600
echo "== bench cfg3" ; timeout 600 python bench.py --workload cfg3 --steps 3 2>&1 | tail -1 | tee gpurun_out/bench_v4_cfg3.json | cut -c1-600
echo "== bench cfg5 (10k-beam sample)" ; timeout 900 python bench.py --workload cfg5 --beams 10000 --steps 1 2>&1 | tail -1 | tee gpurun_out/bench_v4_cfg5_10k.json | cut -c1-600
