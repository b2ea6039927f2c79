#!/bin/bash
# ncu --set full of the fused packed pass (dev build) on one full round (5920 beams) and on 10 000 beams
mkdir -p gpurun_out
L=$PWD/openpystruct_b200/lib
for B in 5920 592; do
OPS_B200_LIB=$L/dev_nbp3.so timeout 600 ncu --set full --clock-control none --import-source on -k regex:beamopt_lanes_kernel -s 3 -c 1 -f -o gpurun_out/r2d_dev_nbp3_$B python scripts/sweep_beams.py $B > gpurun_out/r2d_ncu_$B.log 2>&1; tail -2 gpurun_out/r2d_ncu_$B.log
done
