#!/bin/bash
# A/B of library variants (profiling only): each line "libfile threads"
mkdir -p gpurun_out
L=openpystruct_b200/lib
while read lib thr; do
  [ -z "$lib" ] && continue
  OPS_B200_LIB=$PWD/$L/$lib OPS_LANES_THREADS=$thr timeout 300 python bench.py --steps 5 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import sys,json
try:
    d=json.loads(sys.stdin.read()); print('$lib $thr value %.0f beams/s  kernel_ms %.3f  frac %.4f'%(d['value'],d['roofline']['kernel_ms'],d['roofline']['frac']))
except Exception as ex: print('$lib $thr FAILED',ex)
" | tee -a gpurun_out/ab.txt
done <<LIST
libopenpystruct_b200.so 320
libopenpystruct_b200.so 288
libvariant_nb4.so 320
libvariant_nb5.so 320
libvariant_nb13.so 320
libvariant_nb7_t256.so 256
libvariant_nb13_t256.so 256
libvariant_nb4_t256.so 256
LIST
