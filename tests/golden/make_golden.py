"""Regenerates tests/golden/reference_goldens.{npz,json} -- BUILD CONTAINER ONLY.

Runs the reference's OWN source (/root/reference, read-only, never copied) for the hot path on top
of oracle/opensees_shim.py (OpenSeesPy is not installable here; PARITY UNPINNED at that boundary)
with the container's torch (2.11.0 CPU), and freezes inputs + outputs as fixtures that travel to
the GPU box.  Usage:  python tests/golden/make_golden.py

Cases
  SC  seeds 0..7   SingleCore generate_sample, flag=0   (tol 5e-3, patience 5)
  MC  seeds 0..7   MultiCore  generate_sample, flag=0   (tol 5e-3, patience 10 [def default], last node zeroed)
  GPU seeds 0..1   GPU script generate_sample(device='cpu'), flag=0 (tol 1e-2, patience 100)
  SC1 seeds 0..5   SingleCore generate_sample, flag=1   (random L, 1-4 random rollers)
  BO  seed  0      OpenPyStruct_BeamOpt.py module-level loop (5 spaced rollers, 5 loads, UDL -5000, 1000 epochs)
"""
from __future__ import annotations

import io
import json
import os
import random
import sys
import types
import contextlib

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import opensees_shim                      # noqa: E402
from oracle.reference_loader import (REFERENCE_DIR, run_reference_sample, _install_shim, record_losses)  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def _record(tag, seed, result, tr, meta):
    n = len(result["I_values"])
    rec = {
        "I_values": np.asarray(result["I_values"], np.float32),
        "shear_forces": np.asarray(result["shear_forces"], np.float32),
        "bending_moments": np.asarray(result["bending_moments"], np.float32),
        "rotations": np.asarray(result["rotations"], np.float64),
        "deflections": np.asarray(result["deflections"], np.float64),
        # the last analysed model: inertias handed to OpenSees and its f64 element forces
        "I_last": np.asarray(tr.I[-1], np.float64),
        "V64_last": np.asarray(tr.V[-1], np.float64),
        "M64_last": np.asarray(tr.M[-1], np.float64),
        # first solve (uniform I_0): f64 forces
        "V64_first": np.asarray(tr.V[0], np.float64),
        "M64_first": np.asarray(tr.M[0], np.float64),
        # trajectory of the parameters at a few epochs (value handed to epoch k's analysis)
        "I_trace": np.asarray([tr.I[k] for k in meta["trace_epochs"]], np.float32),
        # float(total_loss) of every epoch: what the early-stop test compares (SingleCore:211-219) -- the decision
        # margin |loss - (best - tolerance)| of a run is computed from it (tests/test_gpu_parity.py)
        "loss_trace": np.asarray(tr.loss, np.float32),
    }
    assert rec["loss_trace"].size == len(tr.I), (rec["loss_trace"].size, len(tr.I))
    assert n == rec["I_last"].size
    return rec


def run_generator_cases():
    cases = []
    plan = [("SC", 0, range(8)), ("MC", 0, range(8)), ("GPU", 0, range(2)), ("SC", 1, range(6))]
    for which, flag, seeds in plan:
        for seed in seeds:
            with contextlib.redirect_stdout(io.StringIO()), contextlib.redirect_stderr(io.StringIO()):
                result, tr, params = run_reference_sample(which, seed, flag=flag)
            epochs = len(tr.I)
            trace_epochs = sorted({k for k in (0, 1, 2, 5, 20, 100, epochs - 1) if k < epochs})
            meta = {
                "script": which, "flag": flag, "seed": seed, "epochs": epochs,
                "L": float(result["L"]), "roller_nodes": [int(t) for t in result["roller_nodes"]],
                "force_nodes": [int(t) for t in result["force_nodes"]],
                "force_values": [float(v) for v in result["force_values"]],
                "num_nodes": int(result["num_nodes"]),
                "tolerance": float(params["tolerance"]),
                "patience": 10 if which == "MC" else int(params["patience"]),
                "max_e": int(params["max_e"]), "uniform_udl": float(params["uniform_udl"]),
                "zero_last_node": which == "MC", "trace_epochs": trace_epochs,
                "roller_x_locations": [float(v) for v in result["roller_x_locations"]],
                "force_x_locations": [float(v) for v in result["force_x_locations"]],
            }
            cases.append((meta, _record(which, seed, result, tr, meta)))
            print(f"{which} flag={flag} seed={seed}: epochs={epochs}", flush=True)
    return cases


def run_beamopt_case(seed=0):
    """OpenPyStruct_BeamOpt.py is a flat script: execute it up to the plotting section."""
    _install_shim()
    mpl = types.ModuleType("matplotlib")
    plt = types.ModuleType("matplotlib.pyplot")
    mpl.pyplot = plt
    sys.modules.setdefault("matplotlib", mpl)
    sys.modules.setdefault("matplotlib.pyplot", plt)
    path = os.path.join(REFERENCE_DIR, "OpenPyStruct_BeamOpt.py")
    text = open(path).read()
    cut = text.index("# Plot loss history")
    tr_I, tr_V, tr_M, tr_loss = [], [], [], []
    real_analyze = opensees_shim.analyze

    def analyze(n=1):
        rc = real_analyze(n)
        d = opensees_shim._D
        tags = sorted(d.elements)
        tr_I.append([d.elements[t][4] for t in tags])
        tr_M.append([float(d.ele_forces[t][2]) for t in tags])
        tr_V.append([float(d.ele_forces[t][1]) for t in tags])
        return rc

    opensees_shim.analyze = analyze
    ns = {"__name__": "_reference_BO", "__file__": path}
    try:
        random.seed(seed)
        with contextlib.redirect_stdout(io.StringIO()), record_losses(tr_loss):
            exec(compile(text[:cut], path, "exec"), ns)
    finally:
        opensees_shim.analyze = real_analyze
    epochs = len(tr_I)
    nn = ns["num_nodes"]

    class T:
        I, V, M, loss = tr_I, tr_V, tr_M, tr_loss
    result = {
        "I_values": ns["I_tensor"].detach().numpy().tolist(),
        "shear_forces": np.asarray(tr_V[-1], np.float32).tolist(),
        "bending_moments": np.asarray(tr_M[-1], np.float32).tolist(),
        "rotations": [opensees_shim.nodeDisp(i, 3) for i in range(1, nn + 1)],
        "deflections": [opensees_shim.nodeDisp(i, 2) for i in range(1, nn + 1)],
    }
    trace_epochs = sorted({k for k in (0, 1, 2, 5, 20, 100, epochs - 1) if k < epochs})
    xs = ns["node_positions"]
    meta = {
        "script": "BO", "flag": 0, "seed": seed, "epochs": epochs, "L": float(ns["L"]),
        "roller_nodes": [int(t) for t in ns["roller_nodes"]],
        "force_nodes": [int(t) for t in ns["force_nodes"]],
        "force_values": [float(v) for v in ns["force_values"]],
        "num_nodes": int(nn), "tolerance": float(ns["tolerance"]), "patience": int(ns["patience"]),
        "max_e": int(ns["num_epochs"]), "uniform_udl": float(ns["uniform_udl"]),
        "zero_last_node": False, "trace_epochs": trace_epochs,
        "roller_x_locations": [float(xs[t - 1]) for t in ns["roller_nodes"]],
        "force_x_locations": [float(xs[t - 1]) for t in ns["force_nodes"]],
    }
    print(f"BO seed={seed}: epochs={epochs}", flush=True)
    return meta, _record("BO", seed, result, T, meta)


def main():
    import torch
    torch.set_num_threads(1)
    cases = run_generator_cases()
    cases.append(run_beamopt_case(0))
    metas = [m for m, _ in cases]
    arrays = {}
    for i, (_, rec) in enumerate(cases):
        for k, v in rec.items():
            arrays[f"{i}:{k}"] = v
    np.savez_compressed(os.path.join(HERE, "reference_goldens.npz"), **arrays)
    with open(os.path.join(HERE, "reference_goldens.json"), "w") as fh:
        json.dump({"torch": torch.__version__, "numpy": np.__version__,
                   # torch.sum's vector layout is ISA-dependent in principle (8-float AVX2 lanes assumed by the kernel's
                   # summation order; SURVEY App. B found AVX2 / AVX512 / default builds bit-identical in practice)
                   "cpu_capability": torch.backends.cpu.get_cpu_capability(),
                   "note": "reference source on oracle/opensees_shim.py; parity unpinned at the OpenSees boundary",
                   "cases": metas}, fh, indent=1)
    print(f"wrote {len(cases)} cases")


if __name__ == "__main__":
    main()
