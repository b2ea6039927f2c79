#!/bin/bash
# single-block pass with the FP64 sums of the previous batch next to the fp32 chain: parity subset, sweep
mkdir -p gpurun_out
L=$PWD/openpystruct_b200/lib
echo "== parity subset (dev_skew)"
OPS_B200_LIB=$L/dev_skew.so timeout 900 python -m pytest tests/test_gpu_parity.py -q -x -k "(three_moment_lanes and (full_loop or fixed_600 or goldens_through or 10k)) or trajectory or many_round or random_bridges or branch_free" 2>&1 | tail -5 | tee gpurun_out/r2f_parity.log
for v in base_r1 dev_nbp3 dev_skew dev_skew_nbp2; do
  echo "== $v"; OPS_B200_LIB=$L/$v.so timeout 300 python scripts/sweep_beams.py 592 5920 10000 23680 2>&1 | grep "^B=" | tee gpurun_out/r2f_sweep_$v.txt
  OPS_B200_LIB=$L/$v.so SWEEP_EARLY_STOP=1 timeout 300 python scripts/sweep_beams.py 10000 100000 2>&1 | grep "^B=" | tee -a gpurun_out/r2f_sweep_$v.txt
done
for v in dev_skew_t384; do
  echo "== $v"; OPS_B200_LIB=$L/$v.so timeout 300 python scripts/sweep_beams.py 7104 10000 28416 2>&1 | grep "^B=" | tee gpurun_out/r2f_sweep_$v.txt
done
