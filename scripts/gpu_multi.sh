#!/bin/bash
# N GPUs of one box (gpurun --gpus N -- 'bash scripts/gpu_multi.sh N [tag]'): the multi-GPU identity check, the driver's
# weak-scaling bench line at N, N/2, ..., 1 on the same box, the NCCL-gather variant and the 1 M-beam strong-scaling line
N=${1:-2}; TAG=${2:-r02}
mkdir -p gpurun_out
echo "== multi_gpu_check on $N GPUs"; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 tests/multi_gpu_check.py 2>&1 | grep -E "multi_gpu_check ok|Error|error" | tee gpurun_out/${TAG}_multi_gpu_check_n$N.log
n=$N
while [ $n -ge 2 ]; do
  echo "== bench N=$n"; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2954$n bench.py --gpus $n --steps 20 --warmup 5 2>/dev/null | tail -1 | tee gpurun_out/${TAG}_scale_n$n.json | cut -c1-200
  n=$((n / 2))
done
echo "== bench N=1"; timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-configs 2>/dev/null | tail -1 | tee gpurun_out/${TAG}_scale_n1.json | cut -c1-200
echo "== bench N=$N, NCCL gather"; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus $N --steps 20 --warmup 5 --gather nccl 2>/dev/null | tail -1 | tee gpurun_out/${TAG}_bench_n${N}_nccl.json | cut -c1-200
echo "== bench N=$N, cfg3 (1 M beams sharded)"; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29552 bench.py --gpus $N --workload cfg3 --steps 3 --warmup 3 2>/dev/null | tail -1 | tee gpurun_out/${TAG}_bench_n${N}_cfg3.json | cut -c1-200
