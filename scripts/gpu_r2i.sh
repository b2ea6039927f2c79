#!/bin/bash
# do resident warps in lockstep cost throughput?  first-epoch start staggered per warp
mkdir -p gpurun_out
L=$PWD/openpystruct_b200/lib
for v in dev_stag0 dev_stag700 dev_stag1500 dev_stag2200; do
  echo "== $v"; OPS_B200_LIB=$L/$v.so timeout 300 python scripts/sweep_beams.py 5920 10000 23680 2>&1 | grep "^B=" | tee gpurun_out/r2i_sweep_$v.txt
done
