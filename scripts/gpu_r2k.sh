#!/bin/bash
mkdir -p gpurun_out
echo "== frames"; timeout 1200 python -m pytest tests/test_frames.py -m gpu -q 2>&1 | tail -25 | tee gpurun_out/r2k_frames.log
