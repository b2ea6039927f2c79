#!/bin/bash
# A/B of library variants (profiling only): each line "libfile threads workload beams"
mkdir -p gpurun_out
L=openpystruct_b200/lib
while read lib thr wl beams; do
  [ -z "$lib" ] && continue
  OPS_B200_LIB=$PWD/$L/$lib OPS_LANES_THREADS=$thr timeout 300 python bench.py --workload $wl --beams $beams --steps 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import sys,json
try:
    d=json.loads(sys.stdin.read()); print('$lib T=$thr $wl B=$beams value %.0f beams/s  kernel_ms %.3f  frac %.4f'%(d['value'],d['roofline']['kernel_ms'],d['roofline']['frac']))
except Exception as ex: print('$lib $thr FAILED',ex)
" | tee -a gpurun_out/ab.txt
done <<LIST
libopenpystruct_b200.so 320 cfg2 10000
libopenpystruct_b200.so 320 cfg3 200000
libvariant_nb5_t256.so 256 cfg2 10000
libvariant_nb5_t256.so 256 cfg2 9472
libvariant_nb5_t256.so 256 cfg3 200000
LIST
