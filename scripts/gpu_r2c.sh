#!/bin/bash
# fused packed pass (dev builds, reference discretisation only): parity subset on the default dev build, then the
# beam-count sweep of every variant against the round-1 library
mkdir -p gpurun_out
L=$PWD/openpystruct_b200/lib
echo "== parity subset (dev_nbp3)"
OPS_B200_LIB=$L/dev_nbp3.so timeout 900 python -m pytest tests/test_gpu_parity.py -q -x -k "(three_moment_lanes and (full_loop or fixed_600 or goldens_through or 10k)) or trajectory or many_round or random_bridges or branch_free" 2>&1 | tail -5 | tee gpurun_out/r2c_parity.log
for v in base_r1 dev_nbp2 dev_nbp3 dev_nbp4; do
  echo "== $v"; OPS_B200_LIB=$L/$v.so timeout 300 python scripts/sweep_beams.py 592 2368 5920 10000 23680 2>&1 | grep "^B=" | tee gpurun_out/r2c_sweep_$v.txt
done
for v in dev_t384; do
  echo "== $v"; OPS_B200_LIB=$L/$v.so timeout 300 python scripts/sweep_beams.py 7104 10000 28416 2>&1 | grep "^B=" | tee gpurun_out/r2c_sweep_$v.txt
done
for v in dev_t448 dev_t448_nbp2; do
  echo "== $v"; OPS_B200_LIB=$L/$v.so timeout 300 python scripts/sweep_beams.py 8288 10000 33152 2>&1 | grep "^B=" | tee gpurun_out/r2c_sweep_$v.txt
done
