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
    def oracle(prec):
        with ThreadPoolExecutor(16) as ex:
            parts = list(ex.map(lambda c: oracle_run(p, fixed[c], fn[c], fv[c], L[c], prec), chunks))
        return {k: np.concatenate([q[k] for q in parts]) for k in parts[0]}

    def compare(o, x):
        ok = (o["status"] == 0) & (x["status"] == 0)
        same = (o["epochs"] == x["epochs"]) & ok
        relI = np.abs(o["I"] - x["I"]) / o["I"]
        return {
            "status_equal": bool(np.array_equal(o["status"], x["status"])),
            "stop_epoch_flips": int((~same & ok).sum()), "flip_rate": float((~same & ok).mean()),
            "max_abs_epoch_difference": int(np.abs(o["epochs"].astype(int) - x["epochs"].astype(int)).max()),
            "loss_bit_identical_fraction_among_same_stop": float((o["loss"][same] == x["loss"][same]).mean()),
            "I_bit_identical_fraction_among_same_stop": float((o["I"][same] == x["I"][same]).all(axis=1).mean()),
            "max_rel_dI_among_same_stop": float(relI[same].max()),
        }

    o64, o80 = oracle(0), oracle(1)
    out[f"{script}_flag{flag}"] = {
        "beams": B, "epochs_mean": float(o64["epochs"].mean()),
        # the reference's arithmetic (FP64 banded Cholesky) restated on the CPU
        "cuda_vs_fp64_oracle": compare(o64, g),
        # the same loop with the FE solve in 80-bit arithmetic: who is closer to the exact solve?
        "cuda_vs_80bit_fe_oracle": compare(o80, g),
        "fp64_oracle_vs_80bit_fe_oracle": compare(o80, o64),
    }
print(json.dumps(out, indent=1))
