#!/bin/bash
# contiguous epoch loop (record path moved behind the Adam step in source order): A/B on one box, then the 2-GPU lines
mkdir -p gpurun_out
: > gpurun_out/ab_r1i.txt
L=openpystruct_b200/lib
run() {
  [ -f $L/$1 ] || return
  OPS_B200_LIB=$PWD/$L/$1 CUDA_VISIBLE_DEVICES=0 timeout 300 python bench.py --workload $2 --beams $3 --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import sys,json
try:
    d=json.loads(sys.stdin.read()); print('$1 $2 B=$3 kernel_ms %.3f value %.0f frac %.4f e2e %.0f es %.0f'%(d['roofline']['kernel_ms'], d['value'], d['roofline']['frac'], d['e2e']['value'], d['early_stop_mode']['value']))
except Exception as ex: print('$1 $2 FAILED',ex)
" | tee -a gpurun_out/ab_r1i.txt
}
for rep in 1 2; do
for lib in libvariant_sc0.so libopenpystruct_b200.so; do
  run $lib cfg2 10000
done
done
run libvariant_sc0.so cfg3 1000000
run libopenpystruct_b200.so cfg3 1000000
N=2
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29511 tests/multi_gpu_check.py 2>&1 | grep "multi_gpu_check" | tee -a gpurun_out/ab_r1i.txt
for g in peer nccl; do
  timeout 600 $TR --master-port 29512 bench.py --gpus $N --steps 10 --warmup 3 --gather $g 2>&1 | grep '^{' | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('N=$N $g value %.0f ms/step %.3f kernel_ms %.3f'%(d['value'], d['ms_per_step'], d['roofline']['kernel_ms']))" | tee -a gpurun_out/ab_r1i.txt
done
