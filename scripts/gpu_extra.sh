#!/bin/bash
# Extra round-2 evidence on ONE B200 (gpurun -- 'bash scripts/gpu_extra.sh [tag]'): early-stop decision parity at full size
# against both oracles (flip_rate.py), ncu --set full of the one-warp-per-beam kernel (1000-element beams) and of the
# 8-load-case team instance of the lanes kernel, the default bench line.
TAG=${1:-r02}
mkdir -p gpurun_out
echo "== bench"; timeout 1500 python bench.py 2>gpurun_out/${TAG}_bench.err | tail -1 | tee gpurun_out/${TAG}_bench.json | cut -c1-300
echo "== flip rate"; timeout 1200 python scripts/flip_rate.py > gpurun_out/${TAG}_flip_rate_40k_beams.json 2>gpurun_out/${TAG}_flip_rate.err; tail -3 gpurun_out/${TAG}_flip_rate.err; grep -c flip_rate gpurun_out/${TAG}_flip_rate_40k_beams.json
echo "== ncu full (wide kernel, 1000 elements)"; SWEEP_WORKLOAD=cfg5 SWEEP_EPOCHS=30 timeout 900 ncu --set full --clock-control none --import-source on -k regex:beamopt_wide_kernel -s 3 -c 1 -f -o gpurun_out/${TAG}_wide python scripts/sweep_beams.py 1776 > gpurun_out/${TAG}_ncu_wide.log 2>&1; tail -1 gpurun_out/${TAG}_ncu_wide.log | cut -c1-200
echo "== ncu full (8 load cases)"; SWEEP_WORKLOAD=cfg4 SWEEP_EPOCHS=60 timeout 900 ncu --set full --clock-control none --import-source on -k regex:beamopt_lanes -s 3 -c 1 -f -o gpurun_out/${TAG}_cases8 python scripts/sweep_beams.py 740 > gpurun_out/${TAG}_ncu_cases8.log 2>&1; tail -1 gpurun_out/${TAG}_ncu_cases8.log | cut -c1-200
