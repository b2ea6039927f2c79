#!/bin/bash
# last check of the round-end library on a 2-GPU box: single-GPU parity suite + smoke, then the multi-GPU check
mkdir -p gpurun_out
echo "== pytest gpu" ; CUDA_VISIBLE_DEVICES=0 timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 | tee gpurun_out/pytest_gpu_last.log
echo "== smoke" ; CUDA_VISIBLE_DEVICES=0 timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
echo "== pytest multi gpu" ; timeout 900 python -m pytest tests/test_multi_gpu.py -m gpu -q 2>&1 | tail -2
echo "== bench N=2" ; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 10 --warmup 3 2>&1 | grep '^{' | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('N=2 value %.0f ms/step %.3f kernel_ms %.3f e2e %.0f'%(d['value'], d['ms_per_step'], d['roofline']['kernel_ms'], d['e2e']['value']))"
