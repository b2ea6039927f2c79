#!/bin/bash
# Dataset gather of fixed-epoch batches on one multi-GPU box: chunks + peer-copy kernel (default beyond two GPUs) against the
# copy inside the kernel (OPS_SCATTER_IN_KERNEL=1) at N = $1 and N / 2, and one GPU of the same box.
N=${1:-8}; TAG=${2:-r2z}
mkdir -p gpurun_out
for n in $N $((N / 2)); do
  for mode in pipe inkernel; do
    if [ $mode = inkernel ]; then export OPS_SCATTER_IN_KERNEL=1; unset OPS_SCATTER_PIPELINED; else unset OPS_SCATTER_IN_KERNEL; export OPS_SCATTER_PIPELINED=1; fi
    echo "== bench N=$n $mode"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2954$n bench.py --gpus $n --steps 20 --warmup 5 --no-cpu-baseline --no-configs 2>/dev/null | tail -1 | tee gpurun_out/${TAG}_bench_n${n}_$mode.json | cut -c1-160
  done
done
unset OPS_SCATTER_IN_KERNEL OPS_SCATTER_PIPELINED
echo "== bench N=1"; timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-configs 2>/dev/null | tail -1 | tee gpurun_out/${TAG}_bench_n1.json | cut -c1-160
