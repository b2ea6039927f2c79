#!/bin/bash
# In-kernel dataset gather on N GPUs of one box, A/B of library builds (LIBS="a.so b.so", loaded through OPS_B200_LIB;
# "default" = the in-tree build): identity check of the default build, then the weak-scaling bench line per build, the
# NCCL-gather variant and one GPU of the same box.
N=${1:-8}; TAG=${2:-r2s}
mkdir -p gpurun_out
L=$PWD/openpystruct_b200/lib
echo "== multi_gpu_check on $N GPUs"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 tests/multi_gpu_check.py 2>&1 | grep -E "multi_gpu_check ok|Error|error" | tee gpurun_out/${TAG}_multi_gpu_check_n$N.log
i=0
for lib in ${LIBS:-default}; do
  i=$((i + 1))
  if [ $lib = default ]; then unset OPS_B200_LIB; else export OPS_B200_LIB=$L/$lib; fi
  for rep in 1 2; do
    echo "== bench N=$N, $lib (run $rep)"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2954$i bench.py --gpus $N --steps 20 --warmup 5 --no-cpu-baseline --no-configs 2>/dev/null | tail -1 | tee gpurun_out/${TAG}_bench_n${N}_${lib%.so}_$rep.json | cut -c1-160
  done
done
unset OPS_B200_LIB
echo "== bench N=$N, NCCL gather"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus $N --steps 20 --warmup 5 --gather nccl --no-cpu-baseline --no-configs 2>/dev/null | tail -1 | tee gpurun_out/${TAG}_bench_n${N}_nccl.json | cut -c1-160
echo "== bench N=1"; timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-configs 2>/dev/null | tail -1 | tee gpurun_out/${TAG}_bench_n1.json | cut -c1-160
