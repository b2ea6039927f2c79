#!/bin/bash
# compute-sanitizer (memcheck + racecheck) on a small batch of every kernel family, and the flip-rate report
mkdir -p gpurun_out
cat > /tmp/san.py <<'PY'
import numpy as np, sys
sys.path.insert(0, '.')
from openpystruct_b200 import _cabi, sampling
from openpystruct_b200.params import BeamOptParams
from tests.helpers import seeded_cases
for solver, nc in ((0, 1), (0, 8), (1, 1), (2, 1)):
    p = BeamOptParams.for_script("SC").replace(max_e=12, solver=solver, num_cases=nc)
    cases = seeded_cases(p, 48 * nc, seed=9, flag=1 if nc == 1 else 0)
    fixed, fn, fv, L = sampling.pack_cases(p.num_nodes, p.max_forces, cases, nc)
    out = _cabi.run_host(p, fixed, fn, fv, L, device=0)
    print("solver", solver, "cases", nc, "ok", int((out["status"] == 0).sum()), "epochs", int(out["epochs"].min()), int(out["epochs"].max()))
PY
for tool in memcheck racecheck; do
  echo "== compute-sanitizer --tool $tool"
  timeout 1200 compute-sanitizer --tool $tool --print-limit 5 python /tmp/san.py 2>&1 | grep -E "solver|ERROR SUMMARY|RACECHECK SUMMARY|Error|hazard" | head -20 | tee -a gpurun_out/sanitizer.txt
done
echo "== flip rate" ; timeout 1200 python scripts/flip_rate.py 2>&1 | tail -60 | tee gpurun_out/flip_rate.json
