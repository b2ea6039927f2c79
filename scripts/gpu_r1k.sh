#!/bin/bash
# N-GPU pass of the round-end build: two-GPU pytest, the full multi-GPU check, weak-scaling bench lines (peer / nccl)
N=${N:-2}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
echo "== pytest multi gpu" ; timeout 900 python -m pytest tests/test_multi_gpu.py -m gpu -q 2>&1 | tail -2 | tee gpurun_out/pytest_multi_gpu_n$N.log
echo "== multi_gpu_check N=$N" ; timeout 600 $TR --master-port 29511 tests/multi_gpu_check.py 2>&1 | grep "multi_gpu_check\|Error\|error" | tail -8 | tee gpurun_out/multi_gpu_check_final_n$N.log
for g in peer nccl; do
  echo "== bench N=$N gather=$g" ; timeout 600 $TR --master-port 29512 bench.py --gpus $N --steps 10 --warmup 3 --gather $g 2>&1 | grep '^{' | tail -1 | tee gpurun_out/bench_final_n${N}_$g.json | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('N=$N $g value %.0f ms/step %.3f kernel_ms %.3f e2e %.0f'%(d['value'], d['ms_per_step'], d['roofline']['kernel_ms'], d['e2e']['value']))"
done
if [ -n "$CFG3" ]; then
  echo "== bench cfg3 N=$N gather=peer" ; timeout 900 $TR --master-port 29513 bench.py --gpus $N --workload cfg3 --steps 3 --warmup 3 2>&1 | grep '^{' | tail -1 | tee gpurun_out/bench_final_cfg3_n${N}_peer.json | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('cfg3 N=$N peer value %.0f ms/step %.3f kernel_ms %.3f e2e %.0f'%(d['value'], d['ms_per_step'], d['roofline']['kernel_ms'], d['e2e']['value']))"
fi
if [ -n "$ALSO4" ]; then
  TR4="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1"
  echo "== bench N=4 gather=peer" ; timeout 600 $TR4 --master-port 29514 bench.py --gpus 4 --steps 10 --warmup 3 2>&1 | grep '^{' | tail -1 | tee gpurun_out/bench_final_n4_peer.json | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('N=4 peer value %.0f ms/step %.3f kernel_ms %.3f'%(d['value'], d['ms_per_step'], d['roofline']['kernel_ms']))"
fi
echo "== bench N=1" ; timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench_final_n1_ref.json | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('N=1 value %.0f ms/step %.3f kernel_ms %.3f e2e %.0f'%(d['value'], d['ms_per_step'], d['roofline']['kernel_ms'], d['e2e']['value']))"
