#!/bin/bash
# latency / throughput curve of the round-1 production kernel: beams per SM 4..40 (1..10 warps), then two rounds
mkdir -p gpurun_out
timeout 600 python scripts/sweep_beams.py 592 1184 2368 3552 4736 5920 6512 7104 8288 10000 11840 17760 23680 2>&1 | grep -v Warning | tee gpurun_out/r2b_sweep_base.txt
