#!/bin/bash
# Scatter-instance variants on ONE GPU (OPS_FORCE_SC runs the SC instance with a single destination):
# s1 = committed (Adam first, copy in the record path), s2 = record path between loss and Adam, s4 / s5 = the same two with the
# peer copy moved to the top of the loop (before the next beam is fetched); nosc = the instance without the scatter
mkdir -p gpurun_out
rm -f gpurun_out/ab2.txt
LIBS="libvariant_s1.so" REPS=1 bash scripts/gpu_ab2.sh
sed -i 's/^libvariant_s1.so/nosc(s1.so)/' gpurun_out/ab2.txt
export OPS_FORCE_SC=1
LIBS="libvariant_s1.so libvariant_s2.so libvariant_s4.so libvariant_s5.so" REPS=2 bash scripts/gpu_ab2.sh
cp gpurun_out/ab2.txt gpurun_out/ab_r1m.txt
