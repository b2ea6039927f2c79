#!/bin/bash
# two GPUs: the 2-GPU pytest (in-kernel gather == NCCL gather, empty shards, fall-back, differing-cases check), bench N=2
# with the NVLink byte counters around it, N=1 on the same box
mkdir -p gpurun_out
echo "== pytest 2 GPUs"; timeout 1200 python -m pytest tests/test_multi_gpu.py -m gpu -q 2>&1 | tail -15 | tee gpurun_out/r2m_pytest_2gpu.log
nvidia-smi nvlink -gt d -i 0 > gpurun_out/r2m_nvlink_before.txt 2>&1
echo "== bench N=2"; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 20 --warmup 5 2>gpurun_out/r2m_bench_n2.err | tail -1 | tee gpurun_out/r2m_bench_n2.json | cut -c1-500
nvidia-smi nvlink -gt d -i 0 > gpurun_out/r2m_nvlink_after.txt 2>&1
tail -3 gpurun_out/r2m_bench_n2.err
echo "== bench N=2 nccl gather"; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus 2 --steps 20 --warmup 5 --gather nccl 2>/dev/null | tail -1 | tee gpurun_out/r2m_bench_n2_nccl.json | cut -c1-300
echo "== bench N=1"; timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-configs 2>/dev/null | tail -1 | tee gpurun_out/r2m_bench_n1.json | cut -c1-300
