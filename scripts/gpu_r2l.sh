#!/bin/bash
# after the frame optimiser / native sampler / ADVICE fixes: whole GPU suite, default bench line (with configs), reference arm
mkdir -p gpurun_out
echo "== pytest gpu"; timeout 2400 python -m pytest tests -m gpu -q 2>&1 | tail -15 | tee gpurun_out/r2l_pytest_gpu.log
echo "== bench"; timeout 1500 python bench.py 2>gpurun_out/r2l_bench.err | tail -1 | tee gpurun_out/r2l_bench.json | cut -c1-600; tail -5 gpurun_out/r2l_bench.err
echo "== reference arm"; timeout 900 python bench.py --impl reference --steps 3 --warmup 1 2>&1 | tail -1 | tee gpurun_out/r2l_bench_reference.json | cut -c1-300
