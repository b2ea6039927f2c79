#!/usr/bin/env python
"""Early-stop decision parity at full size (SURVEY 7 "hard parts": report the flip rate and the decision
margin).  10 000 beams of BASELINE configs[1] with each script's effective early-stop constants, CUDA
path (through the C ABI) vs the plain-C oracle; prints one JSON object.  GPU box only."""
import json
import os
import sys
from concurrent.futures import ThreadPoolExecutor

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from openpystruct_b200 import _cabi, sampling                      # noqa: E402
from openpystruct_b200.params import BeamOptParams                  # noqa: E402
from tests.helpers import oracle_run, seeded_cases                  # noqa: E402

B = int(os.environ.get("FLIP_BEAMS", "10000"))
out = {}
for script, flag in (("SC", 0), ("MC", 0), ("GPU", 0), ("SC", 1)):
    p = BeamOptParams.for_script(script)
    cases = seeded_cases(p, B, seed=2024, flag=flag)
    fixed, fn, fv, L = sampling.pack_cases(p.num_nodes, p.max_forces, cases)
    g = _cabi.run_host(p, fixed, fn, fv, L, device=0)
    chunks = np.array_split(np.arange(B), 16)
    with ThreadPoolExecutor(16) as ex:
        parts = list(ex.map(lambda c: oracle_run(p, fixed[c], fn[c], fv[c], L[c]), chunks))
    o = {k: np.concatenate([q[k] for q in parts]) for k in parts[0]}
    ok = (o["status"] == 0) & (g["status"] == 0)
    same = (o["epochs"] == g["epochs"]) & ok
    relI = np.abs(o["I"] - g["I"]) / o["I"]
    out[f"{script}_flag{flag}"] = {
        "beams": B, "status_equal": bool(np.array_equal(o["status"], g["status"])),
        "stop_epoch_flips": int((~same & ok).sum()), "flip_rate": float((~same & ok).mean()),
        "epochs_mean": float(o["epochs"].mean()),
        "max_abs_epoch_difference": int(np.abs(o["epochs"].astype(int) - g["epochs"].astype(int)).max()),
        "loss_bit_identical_fraction_among_same_stop": float((o["loss"][same] == g["loss"][same]).mean()),
        "I_bit_identical_fraction_among_same_stop": float((o["I"][same] == g["I"][same]).all(axis=1).mean()),
        "max_rel_dI_among_same_stop": float(relI[same].max()),
        "max_rel_dI_all": float(relI[ok].max()),
    }
print(json.dumps(out, indent=1))
