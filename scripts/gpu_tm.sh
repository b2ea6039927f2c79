#!/bin/bash
# Tensor-memory instance of the lanes kernel on one B200: parity suite with the instance forced for every 13-slot
# single-case launch, then kernel time against the register / shared-memory instance (OPS_LANES_TM = 0 / 1) and
# against variant builds (LIBS="tm544.so tm640.so").
TAG=${1:-r2t}
mkdir -p gpurun_out
L=$PWD/openpystruct_b200/lib
echo "== pytest gpu (OPS_LANES_TM=1)"; OPS_LANES_TM=1 timeout 1200 python -m pytest tests -m gpu -q -x 2>&1 | tail -8 | tee gpurun_out/${TAG}_pytest_tm.log
for tm in 0 1; do
  echo "== sweep OPS_LANES_TM=$tm"; OPS_LANES_TM=$tm timeout 300 python scripts/sweep_beams.py ${COUNTS:-5920 9472 10000 23680 100000} 2>&1 | grep "^B=" | tee gpurun_out/${TAG}_sweep_tm$tm.txt
done
echo "== early stop"; for tm in 0 1; do OPS_LANES_TM=$tm SWEEP_EARLY_STOP=1 timeout 300 python scripts/sweep_beams.py 10000 100000 2>&1 | grep "^B=" | tee gpurun_out/${TAG}_sweep_es_tm$tm.txt; done
for lib in $LIBS; do
  [ -f $L/$lib ] || continue
  echo "== $lib"; OPS_LANES_TM=1 OPS_B200_LIB=$L/$lib timeout 300 python scripts/sweep_beams.py ${COUNTS2:-10000 11840 100000} 2>&1 | grep "^B=" | tee gpurun_out/${TAG}_sweep_${lib%.so}.txt
done
