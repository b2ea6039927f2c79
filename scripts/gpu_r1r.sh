#!/bin/bash
# six-slot batches in the 320-thread scatter instance: multi-GPU check + weak-scaling line on 2 GPUs
mkdir -p gpurun_out
echo "== multi_gpu_check N=2" ; timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/multi_gpu_check.py 2>&1 | grep "multi_gpu_check\|Error\|error" | tail -6 | tee gpurun_out/multi_gpu_check_nb6_n2.log
echo "== bench N=2" ; timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 10 --warmup 3 2>&1 | grep '^{' | tail -1 | tee gpurun_out/bench_nb6_n2.json | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('N=2 value %.0f ms/step %.3f kernel_ms %.3f e2e %.0f'%(d['value'], d['ms_per_step'], d['roofline']['kernel_ms'], d['e2e']['value']))"
