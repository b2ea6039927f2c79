#!/bin/bash
# XU diet (reciprocal seeds derived from rsqrt(I), integer widening): selftest + parity subset, sweep, ncu
mkdir -p gpurun_out
L=$PWD/openpystruct_b200/lib
echo "== parity subset (dev_xu)"
OPS_B200_LIB=$L/dev_xu.so timeout 900 python -m pytest tests/test_gpu_parity.py -q -x -k "(three_moment_lanes and (full_loop or fixed_600 or goldens_through or 10k)) or trajectory or many_round or random_bridges or branch_free" 2>&1 | tail -5 | tee gpurun_out/r2e_parity.log
for v in dev_nbp3 dev_xu dev_xu_nbp2; do
  echo "== $v"; OPS_B200_LIB=$L/$v.so timeout 300 python scripts/sweep_beams.py 592 5920 10000 23680 2>&1 | grep "^B=" | tee gpurun_out/r2e_sweep_$v.txt
done
for v in dev_xu_t384; do
  echo "== $v"; OPS_B200_LIB=$L/$v.so timeout 300 python scripts/sweep_beams.py 7104 10000 28416 2>&1 | grep "^B=" | tee gpurun_out/r2e_sweep_$v.txt
done
OPS_B200_LIB=$L/dev_xu.so timeout 600 ncu --set full --clock-control none --import-source on -k regex:beamopt_lanes_kernel -s 3 -c 1 -f -o gpurun_out/r2e_dev_xu_5920 python scripts/sweep_beams.py 5920 > gpurun_out/r2e_ncu.log 2>&1; tail -1 gpurun_out/r2e_ncu.log
