#!/bin/bash
# parity suite + self test (verbose) + default bench line
mkdir -p gpurun_out
OPS_SELFTEST_VERBOSE=1 python -c "
from openpystruct_b200 import _cabi; print(_cabi.fastmath_selftest(1<<27))" 2>&1 | tail -3 | tee gpurun_out/selftest.txt
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 | tee gpurun_out/pytest_gpu_check.log
timeout 600 python bench.py --steps 10 ${BENCH_ARGS} 2>&1 | tail -1 | tee gpurun_out/bench_check.json | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('value %.0f beams/s  kernel_ms %.3f  frac %.4f e2e %.0f'%(d['value'],d['roofline']['kernel_ms'],d['roofline']['frac'],d['e2e']['value']))"
