#!/bin/bash
# last build of the round on one B200: smoke + the driver's default bench line
mkdir -p gpurun_out
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 | tee gpurun_out/smoke_v7.log
timeout 300 python bench.py 2>&1 | tail -1 | tee gpurun_out/bench_v7.json | cut -c1-160
