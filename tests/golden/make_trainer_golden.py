"""Freeze the reference trainer's data block (OpenPyStruct_PINN_MultiCase.py:1-369, executed verbatim by
oracle/reference_loader.run_reference_trainer_block) on a dataset the PRODUCT's writer emitted
(dataset.columnar_from_run -> save_json), as tests/golden/trainer_goldens.npz.  The beams' records come from the CPU
oracle here (build container, no GPU): the fixture pins the trainer-side arithmetic, the GPU test feeds the same
columnar dataset through dataset.trainer_preprocess on the device.
Needs /root/reference (build container only):  python tests/golden/make_trainer_golden.py"""
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import reference_loader as rl                      # noqa: E402
from openpystruct_b200 import dataset, sampling                  # noqa: E402
from openpystruct_b200.params import BeamOptParams              # noqa: E402
from tests.helpers import oracle_run, seeded_cases              # noqa: E402

RUNS = ((6, 0.5, 0), (4, 0.0, 1))                               # (n_cases, c, numpy seed); the script's defaults first


def main():
    p = BeamOptParams.for_script("MC").replace(max_e=60)
    cases = seeded_cases(p, 96, seed=17, flag=1)                # random bridges: ragged roller / force lists
    fixed, fn, fv, L = sampling.pack_cases(p.num_nodes, p.max_forces, cases)
    out = oracle_run(p, fixed, fn, fv, L)
    col = dataset.columnar_from_run(p, cases, out)
    arrays = {"col:" + k: np.asarray(v) for k, v in col.items()}
    with tempfile.TemporaryDirectory() as td:
        path = os.path.join(td, "d.json")
        dataset.save_json(col, path)
        for i, (n_cases, c, seed) in enumerate(RUNS):
            ns = rl.run_reference_trainer_block(path, seed, {"n_cases = 6 ": f"n_cases = {n_cases} ", "c = 0.5 ": f"c = {c} "})
            assert ns["n_cases"] == n_cases and ns["c"] == c
            arrays[f"{i}:config"] = np.array([n_cases, c, seed, ns["train_split"]], np.float64)
            for k in ("X_train_flat", "X_val_flat", "Y_train_std", "Y_val_std", "train_idx", "val_idx", "I_grouped",
                      "roller_grouped", "force_val_grouped"):
                arrays[f"{i}:{k}"] = np.asarray(ns[k])
            print(i, n_cases, c, seed, ns["X_train_flat"].shape, ns["Y_train_std"].shape, ns["Y_val_std"].shape)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "trainer_goldens.npz"), **arrays)


if __name__ == "__main__":
    main()
