#!/bin/bash
# Round-end evidence on ONE B200 (gpurun -- 'bash scripts/gpu_evidence.sh [tag]'): GPU parity suite, smoke, default bench
# line, reference arm, ncu launch list of the bench command, ncu --set full of the production kernel and of the FP64
# probe, compute-sanitizer on small runs.  Everything lands in gpurun_out/<tag>_*; the summaries are copied to profiles/.
TAG=${1:-r02}
mkdir -p gpurun_out
echo "== pytest gpu"; timeout 2400 python -m pytest tests -m gpu -q 2>&1 | tail -6 | tee gpurun_out/${TAG}_pytest_gpu.log
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee gpurun_out/${TAG}_smoke.log
echo "== bench"; timeout 1500 python bench.py 2>gpurun_out/${TAG}_bench.err | tail -1 | tee gpurun_out/${TAG}_bench.json | cut -c1-300
echo "== reference arm"; timeout 900 python bench.py --impl reference --steps 3 --warmup 1 2>/dev/null | tail -1 | tee gpurun_out/${TAG}_bench_reference_arm.json | cut -c1-200
echo "== ncu launch list"; timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-configs > gpurun_out/${TAG}_ncu_launch.log 2>&1; tail -1 gpurun_out/${TAG}_ncu_launch.log | cut -c1-160
echo "== ncu full (lanes kernel)"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:beamopt_lanes_kernel -s 3 -c 1 -f -o gpurun_out/${TAG}_lanes python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-configs > gpurun_out/${TAG}_ncu_full.log 2>&1; tail -1 gpurun_out/${TAG}_ncu_full.log | cut -c1-160
echo "== ncu (FP64 probe)"; timeout 600 ncu --clock-control none -k regex:fp64_probe_kernel -s 1 -c 1 --metrics sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,gpu__time_duration.sum --csv --log-file gpurun_out/${TAG}_fp64_probe_ncu.csv python -c "
from openpystruct_b200 import _cabi
print(_cabi.fp64_peak_probe(1 << 16))" > gpurun_out/${TAG}_probe.log 2>&1; tail -1 gpurun_out/${TAG}_probe.log
echo "== ncu full (frame kernel)"; timeout 600 ncu --set full --clock-control none -k regex:frameopt_kernel -c 1 -f -o gpurun_out/${TAG}_frames python -c "
from openpystruct_b200 import frames
import random
p = frames.FrameOptParams(num_epochs=100, early_stop=False)
rng = random.Random(0)
r = frames.optimise_frames([frames.draw_frame(p, rng) for _ in range(592)], p)
print(len(r), r[0]['epochs'])" > gpurun_out/${TAG}_ncu_frames.log 2>&1; tail -1 gpurun_out/${TAG}_ncu_frames.log
bash scripts/gpu_sanitize.sh ${TAG}
