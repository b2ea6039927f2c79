#!/bin/bash
# compute-sanitizer on the round-end build: lanes kernel (plain, 8 load cases, scatter instance with one destination),
# fine-discretisation kernel, pipelined session with three chunked launches
mkdir -p gpurun_out
cat > /tmp/san2.py <<'PY'
import os, sys
import numpy as np
sys.path.insert(0, '.')
from openpystruct_b200 import _cabi, sampling
from openpystruct_b200.params import BeamOptParams
from tests.helpers import seeded_cases
mode = sys.argv[1]
jobs = {"lanes": ((0, 1, 101, 48), (0, 8, 101, 24)), "wide": ((0, 1, 1001, 12),), "session": ((0, 1, 101, 148 * 40 * 2 + 500),)}[mode]
for solver, nc, nn, B in jobs:
    p = BeamOptParams.for_script("SC").replace(max_e=6 if mode == "session" else 12, solver=solver, num_cases=nc, num_nodes=nn,
                                                early_stop=False)
    rollers = [100, 300, 700, 850, 1000] if nn == 1001 else None
    base = seeded_cases(p, min(B, 256) * nc, seed=9, flag=0, roller_nodes=rollers)
    cases = (base * ((B * nc + len(base) - 1) // len(base)))[:B * nc]
    fixed, fn, fv, L = sampling.pack_cases(p.num_nodes, p.max_forces, cases, nc)
    out = _cabi.run_host(p, fixed, fn, fv, L, device=0)
    print(mode, "force_sc" if os.environ.get("OPS_FORCE_SC") else "", "nodes", nn, "cases", nc, "beams", B, "ok", int((out["status"] == 0).sum()),
          "epochs", int(out["epochs"].min()), int(out["epochs"].max()), flush=True)
PY
run() {  # tool mode [env]
  echo "== compute-sanitizer --tool $1 ($2 $3)" | tee -a gpurun_out/sanitizer2.txt
  env $3 timeout 240 compute-sanitizer --tool $1 --print-limit 5 python /tmp/san2.py $2 2>&1 | grep -E "^lanes|^wide|^session|ERROR SUMMARY|RACECHECK SUMMARY|Error|hazard" | head -12 | tee -a gpurun_out/sanitizer2.txt
}
: > gpurun_out/sanitizer2.txt
run memcheck lanes A=1
run memcheck lanes OPS_FORCE_SC=1
run memcheck wide A=1
run memcheck session A=1
run racecheck lanes OPS_FORCE_SC=1
run racecheck wide A=1
