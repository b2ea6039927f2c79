#!/bin/bash
# six-slot batches: plain instance and many-round instance against the committed five-slot build (one GPU)
mkdir -p gpurun_out
rm -f gpurun_out/ab2.txt
LIBS="libopenpystruct_b200.so libvariant_n6.so" REPS=2 bash scripts/gpu_ab2.sh
echo "--- 213120 beams (384-thread instance)" >> gpurun_out/ab2.txt
LIBS="libopenpystruct_b200.so libvariant_n6.so" REPS=1 WL=cfg3 BEAMS=213120 bash scripts/gpu_ab2.sh
cp gpurun_out/ab2.txt gpurun_out/ab_r1q.txt
