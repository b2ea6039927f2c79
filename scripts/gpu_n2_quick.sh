mkdir -p gpurun_out
echo "== multi_gpu_check"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 tests/multi_gpu_check.py 2>&1 | grep -E "multi_gpu_check ok|Error|error|assert" | tee gpurun_out/r2w_multi_gpu_check_n2.log
for mode in pipe inkernel; do
  if [ $mode = inkernel ]; then export OPS_SCATTER_IN_KERNEL=1; else unset OPS_SCATTER_IN_KERNEL; fi
  echo "== bench N=2 $mode"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu-baseline --no-configs 2>gpurun_out/r2w_n2_$mode.err | tail -1 | tee gpurun_out/r2w_bench_n2_$mode.json | cut -c1-160
done
unset OPS_SCATTER_IN_KERNEL
echo "== bench N=1"; timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-configs 2>/dev/null | tail -1 | tee gpurun_out/r2w_bench_n1.json | cut -c1-160
