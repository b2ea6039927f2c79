#!/bin/bash
# One-warp-per-beam kernel (1000-element beams) on ONE B200: its GPU parity tests, kernel time against the number of
# beams per SM, ncu --set full of a steady-state launch (gpurun -- 'bash scripts/gpu_wide.sh [tag]').
TAG=${1:-r02}
mkdir -p gpurun_out
echo "== pytest (fine discretisation)"; timeout 1200 python -m pytest tests -m gpu -q -k "fine or 1000 or discretis or wide" 2>&1 | tail -3 | tee gpurun_out/${TAG}_pytest_wide.log
echo "== sweep"; SWEEP_WORKLOAD=cfg5 timeout 900 python scripts/sweep_beams.py 1628 1776 3552 17760 100000 2>&1 | tee gpurun_out/${TAG}_wide_sweep.txt
echo "== ncu full"; SWEEP_WORKLOAD=cfg5 SWEEP_EPOCHS=200 timeout 900 ncu --set full --clock-control none --import-source on -k regex:beamopt_wide_kernel -s 3 -c 1 -f -o gpurun_out/${TAG}_wide python scripts/sweep_beams.py 1776 > gpurun_out/${TAG}_ncu_wide.log 2>&1; tail -1 gpurun_out/${TAG}_ncu_wide.log | cut -c1-200
