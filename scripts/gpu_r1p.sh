#!/bin/bash
# scatter instance (OPS_FORCE_SC, one GPU) with other batch sizes: does any order ptxas picks reach the plain instance's speed?
mkdir -p gpurun_out
rm -f gpurun_out/ab2.txt
LIBS="libopenpystruct_b200.so" REPS=1 bash scripts/gpu_ab2.sh
sed -i 's/^libopenpystruct_b200.so/plain(nb5)/' gpurun_out/ab2.txt
export OPS_FORCE_SC=1
LIBS="libopenpystruct_b200.so libvariant_n4.so libvariant_n6.so libvariant_n7.so" REPS=2 bash scripts/gpu_ab2.sh
cp gpurun_out/ab2.txt gpurun_out/ab_r1p.txt
