#!/bin/bash
# ncu full capture of the scatter instance run on one GPU (OPS_FORCE_SC), to compare its stall profile with the plain instance's
mkdir -p gpurun_out
OPS_B200_LIB=$PWD/openpystruct_b200/lib/libvariant_s1.so OPS_FORCE_SC=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:beamopt_lanes_kernel -s 3 -c 1 -f -o gpurun_out/prof_sc python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full_sc.log 2>&1 ; tail -1 gpurun_out/ncu_full_sc.log
