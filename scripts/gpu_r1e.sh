#!/bin/bash
# Evidence pass of the trimmed build: default bench line, reference arm, the other BASELINE configs, launch list,
# ncu full capture of the fine-discretisation kernel (cfg5)
mkdir -p gpurun_out
echo "== bench default" ; timeout 900 python bench.py 2>&1 | tail -1 | tee gpurun_out/bench_v5.json | cut -c1-250
echo "== bench reference arm" ; timeout 900 python bench.py --impl reference --steps 3 --warmup 1 2>&1 | tail -1 | tee gpurun_out/bench_v5_reference.json | cut -c1-250
for wl in cfg3 cfg4; do
  echo "== bench $wl" ; timeout 900 python bench.py --workload $wl --steps 3 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench_v5_$wl.json | cut -c1-250
done
echo "== bench cfg5 (20k beams)" ; timeout 900 python bench.py --workload cfg5 --beams 20000 --steps 3 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench_v5_cfg5_20k.json | cut -c1-250
echo "== ncu launches" ; timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_v5.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launch_v5.log 2>&1 ; tail -1 gpurun_out/ncu_launch_v5.log | cut -c1-200
echo "== ncu full cfg5" ; timeout 900 ncu --set full --clock-control none --import-source on -k regex:beamopt_wide_kernel -s 3 -c 1 -f -o gpurun_out/prof_v5_wide python bench.py --workload cfg5 --beams 4736 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full_v5_wide.log 2>&1 ; tail -2 gpurun_out/ncu_full_v5_wide.log
