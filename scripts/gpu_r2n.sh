#!/bin/bash
mkdir -p gpurun_out
echo "== new tests"; timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "native_sampler or more_than_five or shared_inertia or drop_in or columnar" 2>&1 | tail -8 | tee gpurun_out/r2n_tests.log
echo "== bench"; timeout 1500 python bench.py --no-cpu-baseline 2>gpurun_out/r2n_bench.err | tail -1 | tee gpurun_out/r2n_bench.json | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('value %.0f frac %.4f e2e %.0f'%(d['value'],d['roofline']['frac'],d['e2e']['value'])); print('api', d['e2e']['api_e2e']); print({k:(round(v['value']),round(v.get('roofline_frac',0),4)) for k,v in d['configs'].items()})"
tail -3 gpurun_out/r2n_bench.err
