#!/bin/bash
# Round-end evidence: parity suite, default bench line, reference arm, launch list, ncu full capture.
mkdir -p gpurun_out
echo "== pytest gpu" ; timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 | tee gpurun_out/pytest_gpu_final.log
echo "== smoke" ; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee gpurun_out/smoke_final.log
echo "== bench" ; timeout 900 python bench.py 2>&1 | tail -1 | tee gpurun_out/bench_final.json | cut -c1-300
echo "== bench reference arm" ; timeout 900 python bench.py --impl reference --steps 3 --warmup 1 2>&1 | tail -1 | tee gpurun_out/bench_final_reference.json | cut -c1-300
echo "== ncu launches" ; timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_final.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launch_final.log 2>&1 ; tail -1 gpurun_out/ncu_launch_final.log | cut -c1-200
echo "== ncu full" ; timeout 900 ncu --set full --clock-control none --import-source on -k regex:beamopt_lanes_kernel -s 3 -c 1 -f -o gpurun_out/prof_final python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full_final.log 2>&1 ; tail -2 gpurun_out/ncu_full_final.log
python -c "
from openpystruct_b200 import _cabi
import json; print(json.dumps(_cabi.pipe_probe()))" | tee gpurun_out/pipe_probe.json
