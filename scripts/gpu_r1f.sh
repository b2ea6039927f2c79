#!/bin/bash
# N-GPU pass (gpurun --gpus N): in-kernel dataset gather against the NCCL gather, then the weak-scaling bench lines
N=${N:-2}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
echo "== multi_gpu_check" ; timeout 600 $TR --master-port 29511 tests/multi_gpu_check.py 2>&1 | grep -v "^W\|^\*\*\*\|OMP_NUM" | tail -8 | tee gpurun_out/multi_gpu_check_n$N.log
for g in peer nccl; do
  echo "== bench N=$N gather=$g" ; timeout 600 $TR --master-port 29512 bench.py --gpus $N --steps 10 --warmup 3 --gather $g 2>&1 | grep '^{' | tail -1 | tee gpurun_out/bench_n${N}_$g.json | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('N=$N $g value %.0f ms/step %.3f kernel_ms %.3f'%(d['value'], d['ms_per_step'], d['roofline']['kernel_ms']))"
done
echo "== bench N=1" ; timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench_n1_ref.json | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('N=1 value %.0f ms/step %.3f kernel_ms %.3f'%(d['value'], d['ms_per_step'], d['roofline']['kernel_ms']))"
if [ -n "$CFG3" ]; then
for g in peer nccl; do
  echo "== bench cfg3 N=$N gather=$g" ; timeout 900 $TR --master-port 29513 bench.py --gpus $N --workload cfg3 --steps 3 --warmup 3 --gather $g 2>&1 | grep '^{' | tail -1 | tee gpurun_out/bench_cfg3_n${N}_$g.json | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('cfg3 N=$N $g value %.0f ms/step %.3f kernel_ms %.3f'%(d['value'], d['ms_per_step'], d['roofline']['kernel_ms']))"
done
fi
