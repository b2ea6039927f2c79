#!/bin/bash
# First GPU pass: parity tests, smoke, bench (smem + global variants), ncu launch list + full capture.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > gpurun_out/gpu.txt 2>&1
lscpu | head -20 > gpurun_out/cpu.txt 2>&1
echo "== pytest gpu" ; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 | tee gpurun_out/pytest_gpu.log
echo "== smoke" ; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5 | tee gpurun_out/smoke.log
echo "== bench" ; timeout 900 python bench.py --steps 5 --warmup 3 2>&1 | tail -3 | tee gpurun_out/bench_smem.json
echo "== bench global ws 8 ctas/sm" ; OPS_BEAMOPT_GLOBAL=1 OPS_BEAMOPT_CTAS_PER_SM=8 timeout 600 python bench.py --steps 3 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench_global8.json
echo "== bench global ws 16 ctas/sm" ; OPS_BEAMOPT_GLOBAL=1 OPS_BEAMOPT_CTAS_PER_SM=16 timeout 600 python bench.py --steps 3 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench_global16.json
echo "== ncu launches" ; timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r1.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launch_bench.log 2>&1 ; tail -2 gpurun_out/ncu_launch_bench.log
echo "== ncu full" ; timeout 900 ncu --set full --clock-control none --import-source on -k regex:beamopt_kernel -s 3 -c 1 -f -o gpurun_out/prof_r1 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full_bench.log 2>&1 ; tail -2 gpurun_out/ncu_full_bench.log
ls -la gpurun_out
