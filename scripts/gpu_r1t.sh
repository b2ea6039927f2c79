#!/bin/bash
# many-round batch (384-thread instance): plain against scatter instance on one GPU
mkdir -p gpurun_out
rm -f gpurun_out/ab2.txt
LIBS="libopenpystruct_b200.so" REPS=1 WL=cfg3 BEAMS=213120 bash scripts/gpu_ab2.sh
OPS_FORCE_SC=1 LIBS="libopenpystruct_b200.so" REPS=1 WL=cfg3 BEAMS=213120 bash scripts/gpu_ab2.sh
cp gpurun_out/ab2.txt gpurun_out/ab_r1t.txt
