#!/bin/bash
# 8-GPU box, round-end build: multi-GPU check on 8 ranks, weak-scaling lines at N = 8, 4, 2, 1 (in-kernel gather)
mkdir -p gpurun_out
tr() { echo "python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port $2"; }
echo "== multi_gpu_check N=8" ; timeout 600 $(tr 8 29511) tests/multi_gpu_check.py 2>&1 | grep "multi_gpu_check\|Error\|error" | tail -8 | tee gpurun_out/multi_gpu_check_final_n8.log
for N in 8 4 2; do
  echo "== bench N=$N (peer)" ; timeout 600 $(tr $N 2951$N) bench.py --gpus $N --steps 10 --warmup 3 2>&1 | grep '^{' | tail -1 | tee gpurun_out/bench_scale_n${N}.json | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('N=$N value %.0f ms/step %.3f kernel_ms %.3f e2e %.0f'%(d['value'], d['ms_per_step'], d['roofline']['kernel_ms'], d['e2e']['value']))"
done
echo "== bench N=1" ; timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench_scale_n1.json | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('N=1 value %.0f ms/step %.3f kernel_ms %.3f e2e %.0f'%(d['value'], d['ms_per_step'], d['roofline']['kernel_ms'], d['e2e']['value']))"
