#!/bin/bash
# 448 threads (56 beams per SM and round) at a 144-register cap against 320 / 384 threads at 168
mkdir -p gpurun_out
L=$PWD/openpystruct_b200/lib
for v in dev_t448_r144_nbp2 dev_t448_r144_nbp3; do
  echo "== $v"; OPS_B200_LIB=$L/$v.so timeout 300 python scripts/sweep_beams.py 8288 10000 33152 100000 2>&1 | grep "^B=" | tee gpurun_out/r2h_sweep_$v.txt
done
for v in dev_t384; do
  echo "== $v"; OPS_B200_LIB=$L/$v.so timeout 300 python scripts/sweep_beams.py 7104 10000 28416 100000 2>&1 | grep "^B=" | tee gpurun_out/r2h_sweep_$v.txt
done
echo "== production lib"; timeout 300 python scripts/sweep_beams.py 5920 10000 23680 100000 2>&1 | grep "^B=" | tee gpurun_out/r2h_sweep_prod.txt
echo "== base_r1"; OPS_B200_LIB=$L/base_r1.so timeout 300 python scripts/sweep_beams.py 5920 10000 23680 100000 2>&1 | grep "^B=" | tee gpurun_out/r2h_sweep_base.txt
