#!/usr/bin/env python
"""Dump the outputs of the one-warp-per-beam kernel for a bitwise A/B of two library builds (GPU box):
    OPS_B200_LIB=<lib> python scripts/wide_dump.py out.npz
1000-element beams (rollers x10) and 400-element random bridges, early-stopped and with a fixed epoch count."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from openpystruct_b200 import _cabi, sampling                        # noqa: E402
from openpystruct_b200.params import BeamOptParams                   # noqa: E402
from tests.helpers import seeded_cases                               # noqa: E402

out = {}
for tag, nn, flag, rollers, early, max_e, B in (("n1000_fix", 1001, 0, [100, 300, 700, 850, 1000], False, 120, 600),
                                                ("n1000_es", 1001, 0, [100, 300, 700, 850, 1000], True, 600, 300),
                                                ("n400_rand", 401, 1, None, True, 300, 500)):
    p = BeamOptParams.for_script("MC").replace(num_nodes=nn, early_stop=early, max_e=max_e)
    cases = seeded_cases(p, B, seed=77, flag=flag, roller_nodes=rollers) if rollers else seeded_cases(p, B, seed=77, flag=flag)
    fixed, fn, fv, L = sampling.pack_cases(p.num_nodes, p.max_forces, cases)
    r = _cabi.run_host(p, fixed, fn, fv, L, device=0)
    for k, v in r.items():
        out[f"{tag}_{k}"] = np.asarray(v)
np.savez(sys.argv[1], **out)
print("dumped", len(out), "arrays", {k: int(v.sum()) for k, v in out.items() if k.endswith("epochs")})
