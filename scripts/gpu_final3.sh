#!/bin/bash
# Round-end evidence (final build of round 1): parity suite, smoke, default bench line, reference arm, other configs,
# launch list, ncu full capture of the production kernel.
mkdir -p gpurun_out
echo "== pytest gpu" ; timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/pytest_gpu_final3.log
echo "== smoke" ; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee gpurun_out/smoke_final3.log
echo "== bench" ; timeout 900 python bench.py 2>&1 | tail -1 | tee gpurun_out/bench_final3.json | cut -c1-200
echo "== bench reference arm" ; timeout 900 python bench.py --impl reference --steps 3 --warmup 1 2>&1 | tail -1 | tee gpurun_out/bench_final3_reference.json | cut -c1-200
for wl in cfg3 cfg4; do
  echo "== bench $wl" ; timeout 900 python bench.py --workload $wl --steps 3 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench_final3_$wl.json | cut -c1-200
done
echo "== bench cfg5 (20k beams)" ; timeout 900 python bench.py --workload cfg5 --beams 20000 --steps 3 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench_final3_cfg5_20k.json | cut -c1-200
echo "== ncu launches" ; timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_final3.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launch_final3.log 2>&1 ; tail -1 gpurun_out/ncu_launch_final3.log | cut -c1-120
echo "== ncu full" ; timeout 900 ncu --set full --clock-control none --import-source on -k regex:beamopt_lanes_kernel -s 3 -c 1 -f -o gpurun_out/prof_final3 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full_final3.log 2>&1 ; tail -1 gpurun_out/ncu_full_final3.log
