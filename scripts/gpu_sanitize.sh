#!/bin/bash
# compute-sanitizer (memcheck + racecheck) on small runs of every kernel family of the current build
TAG=${1:-r02}
mkdir -p gpurun_out
cat > /tmp/san.py <<'PY'
import os, sys, random
import numpy as np
sys.path.insert(0, '.')
from openpystruct_b200 import _cabi, sampling, frames
from openpystruct_b200.params import BeamOptParams
from tests.helpers import seeded_cases
mode = sys.argv[1]
if mode == "frames":
    p = frames.FrameOptParams(num_epochs=6, early_stop=False)
    r = frames.optimise_frames([(1, 1), (3, 4), (10, 10), (2, 9)], p)
    print("frames", [x["epochs"] for x in r], [x["status"] for x in r], flush=True)
    sys.exit(0)
jobs = {"lanes": ((0, 1, 101, 48, 1), (0, 8, 101, 24, 1), (0, 1, 64, 40, 1), (0, 1, 101, 44, 0)), "wide": ((0, 1, 1001, 12, 0),),
        "session": ((0, 1, 101, 148 * 40 * 2 + 500, 0),), "ldlt": ((1, 1, 101, 64, 1), (2, 1, 101, 64, 1))}[mode]
for solver, nc, nn, B, es in jobs:
    p = BeamOptParams.for_script("SC").replace(max_e=6 if mode == "session" else 14, solver=solver, num_cases=nc, num_nodes=nn,
                                                early_stop=bool(es), patience=2)
    rollers = [100, 300, 700, 850, 1000] if nn == 1001 else ([6, 19, 45, 54, 63] if nn == 64 else None)
    base = seeded_cases(p, min(B, 256) * nc, seed=9, flag=0, roller_nodes=rollers)
    cases = (base * ((B * nc + len(base) - 1) // len(base)))[:B * nc]
    fixed, fn, fv, L = sampling.pack_cases(p.num_nodes, p.max_forces, cases, nc)
    out = _cabi.run_host(p, fixed, fn, fv, L, device=0)
    print(mode, "force_sc" if os.environ.get("OPS_FORCE_SC") else "", "solver", solver, "nodes", nn, "cases", nc, "beams", B, "ok",
          int((out["status"] == 0).sum()), "epochs", int(out["epochs"].min()), int(out["epochs"].max()), flush=True)
PY
OUT=gpurun_out/${TAG}_compute_sanitizer.txt
run() {  # tool mode [env]
  echo "== compute-sanitizer --tool $1 ($2 $3)" | tee -a $OUT
  env $3 timeout 300 compute-sanitizer --tool $1 --print-limit 5 python /tmp/san.py $2 2>&1 | grep -E "^lanes|^wide|^session|^frames|^ldlt|ERROR SUMMARY|RACECHECK SUMMARY|Error|hazard" | head -12 | tee -a $OUT
}
: > $OUT
run memcheck lanes A=1
run memcheck lanes OPS_FORCE_SC=1
run memcheck wide A=1
run memcheck frames A=1
run memcheck session A=1
run memcheck ldlt A=1
run racecheck lanes OPS_FORCE_SC=1
run racecheck frames A=1
run racecheck wide A=1
