#!/bin/bash
# Variant lottery on one box (same epoch-loop instructions, different register allocation / placement), then the
# fine-discretisation kernel with a 384-thread bound and larger batches
mkdir -p gpurun_out
rm -f gpurun_out/ab2.txt
LIBS="libvariant_sc0.so libopenpystruct_b200.so libvariant_xo.so libvariant_r160.so" REPS=2 bash scripts/gpu_ab2.sh
echo "--- cfg3 200k" | tee -a gpurun_out/ab2.txt
LIBS="libopenpystruct_b200.so libvariant_xo.so libvariant_r160.so" REPS=1 WL=cfg3 BEAMS=213120 bash scripts/gpu_ab2.sh
echo "--- cfg5 6512 beams" | tee -a gpurun_out/ab2.txt
LIBS="libopenpystruct_b200.so libvariant_w4.so libvariant_w6.so libvariant_w8.so" REPS=1 WL=cfg5 BEAMS=6512 bash scripts/gpu_ab2.sh
cp gpurun_out/ab2.txt gpurun_out/ab_r1j.txt
