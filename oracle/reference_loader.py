"""TEST INFRASTRUCTURE ONLY -- runs the reference's OWN generator source on top of the shim.

Works only where ``/root/reference`` is mounted (the build container).  Nothing here is
reachable from ``-m gpu`` tests, ``smoke()`` or ``bench.py``: those use the committed
fixtures in ``tests/golden/`` that ``tests/golden/make_golden.py`` produces with this module.

The reference scripts are flat modules whose tail (``if __name__ == '__main__': main()``
followed by a module-level re-load of the JSON they just wrote, SingleCore:271-298) cannot
be imported.  We therefore compile the script text only up to the ``__main__`` guard and
execute it in a fresh namespace with ``openseespy.opensees`` resolved to
``oracle.opensees_shim`` -- ``setup_model`` / ``generate_sample`` then run verbatim.
No reference source is copied into the repository.
"""
from __future__ import annotations

import os
import random
import sys
import types

from . import opensees_shim

REFERENCE_DIR = os.environ.get("OPENPYSTRUCT_REFERENCE_DIR", "/root/reference")

SCRIPTS = {
    "SC": "OpenPyStruct_BeamOpt_training_SingleCore.py",
    "MC": "OpenPyStruct_BeamOpt_training_MultiCore.py",
    "GPU": "OpenPyStruct_BeamOpt_training_GPU.py",
}


def reference_available() -> bool:
    return all(os.path.isfile(os.path.join(REFERENCE_DIR, f)) for f in SCRIPTS.values())


def _install_shim():
    pkg = types.ModuleType("openseespy")
    pkg.opensees = opensees_shim
    pkg.__path__ = []
    sys.modules["openseespy"] = pkg
    sys.modules["openseespy.opensees"] = opensees_shim


def load_generator(which: str) -> dict:
    """Namespace of reference script ``which`` ('SC' | 'MC' | 'GPU') with main() not run."""
    _install_shim()
    path = os.path.join(REFERENCE_DIR, SCRIPTS[which])
    with open(path, "r") as fh:
        text = fh.read()
    cut = text.index('if __name__ == "__main__":')
    ns = {"__name__": f"_reference_{which}", "__file__": path}
    exec(compile(text[:cut], path, "exec"), ns)
    return ns


class Trace:
    """Per-epoch record taken at the shim boundary (what OpenSees would have seen/returned)."""

    def __init__(self):
        self.I = []       # element inertias handed to ops.element (fp32 values widened), per epoch
        self.M = []       # eleResponse[2] per element (f64), per epoch
        self.V = []       # eleResponse[1] per element (f64), per epoch
        self.loss = []    # float(total_loss) per epoch: the fp32 value the early-stop test compares (SingleCore:211)


class record_losses:
    """Context manager: every ``Tensor.backward()`` of a scalar appends ``float(tensor)`` to ``sink`` -- the reference
    calls ``total_loss.backward()`` exactly once per epoch (SingleCore:202, BeamOpt:169, FrameOpt:196), so this
    captures the loss trajectory without touching its source."""

    def __init__(self, sink):
        self.sink = sink

    def __enter__(self):
        import torch
        self._torch, self._real = torch, torch.Tensor.backward
        sink, real = self.sink, self._real

        def backward(t, *a, **k):
            if t.numel() == 1:
                sink.append(float(t.detach()))
            return real(t, *a, **k)

        torch.Tensor.backward = backward
        return self

    def __exit__(self, *exc):
        self._torch.Tensor.backward = self._real


def run_reference_sample(which: str, seed: int, *, flag: int = 0, patience=None, overrides=None,
                         trace: bool = True):
    """random.seed(seed); reference generate_sample(...) with that script's own main() arguments.

    Returns (result_dict, Trace, params) where params are the module-level constants in effect.
    """
    ns = load_generator(which)
    if overrides:
        ns.update(overrides)
    ns["flag"] = flag
    tr = Trace()
    real_analyze = opensees_shim.analyze

    def analyze(n=1):
        rc = real_analyze(n)
        if trace and rc == 0:
            d = opensees_shim._D
            tags = sorted(d.elements)
            tr.I.append([d.elements[t][4] for t in tags])
            tr.M.append([float(d.ele_forces[t][2]) for t in tags])
            tr.V.append([float(d.ele_forces[t][1]) for t in tags])
        return rc

    opensees_shim.analyze = analyze
    try:
        random.seed(seed)
        args = (0, ns["num_nodes"], flag, ns["L"], ns["node_positions"],
                list(ns["roller_nodes"]), list(ns["available_nodes"]))
        if which == "MC":
            # MultiCore:259-261 -- patience is NOT forwarded, the def default (10) applies.
            kwargs = {} if patience is None else {"patience": patience}
        elif which == "SC":
            kwargs = {"patience": ns["patience"] if patience is None else patience}   # SingleCore:257
        else:
            kwargs = {"patience": ns["patience"] if patience is None else patience, "device": "cpu"}
        with record_losses(tr.loss):
            result = ns["generate_sample"](*args, **kwargs)
    finally:
        opensees_shim.analyze = real_analyze
    params = {k: ns[k] for k in ("E", "nu", "G", "A", "L_max", "num_nodes", "max_force", "min_force",
                                 "uniform_udl", "I_0", "max_e", "lr", "gamma", "alpha_moment",
                                 "alpha_shear", "tolerance", "patience", "L_min", "N_rollers_max",
                                 "M_forces_max")}
    return result, tr, params


# ------------------------------------------------------------------------------------------------
# Frame optimiser (groundwork for the next round): OpenPyStruct_FrameOpt_Discrete_Beta.py
# is one flat script -- geometry drawn with ``random.randint`` (:50-51), Adam loop at module level (:179-206),
# then matplotlib plots.  It is executed verbatim up to the plots on top of the shim.
# ------------------------------------------------------------------------------------------------
FRAME_SCRIPT = "OpenPyStruct_FrameOpt_Discrete_Beta.py"


class _Anything:
    """Stand-in for matplotlib.pyplot (absent in this image): every attribute is a no-op callable."""

    def __getattr__(self, name):
        return self

    def __call__(self, *a, **k):
        return self


def frame_reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_DIR, FRAME_SCRIPT))


def run_reference_frame(seed: int, overrides=None):
    """random.seed(seed); the reference frame script up to its plots.  ``overrides``: text-level replacements of
    module constants, e.g. {"num_epochs  = 5000": "num_epochs  = 300"} (the script has no functions to call).
    Returns (namespace, Trace): the namespace holds num_bays, num_stories, loss_history, opt_I, best_loss, ..."""
    import io
    import contextlib
    _install_shim()
    mpl = types.ModuleType("matplotlib")
    mpl.pyplot = _Anything()
    mpl.__path__ = []
    saved = {k: sys.modules.get(k) for k in ("matplotlib", "matplotlib.pyplot")}
    sys.modules["matplotlib"] = mpl
    sys.modules["matplotlib.pyplot"] = mpl.pyplot
    path = os.path.join(REFERENCE_DIR, FRAME_SCRIPT)
    with open(path, "r") as fh:
        text = fh.read()
    cut = text.index("# Plot Loss History")
    cut = text.rindex("##############################", 0, cut)
    text = text[:cut]
    for old, new in (overrides or {}).items():
        if old not in text:
            raise KeyError(old)
        text = text.replace(old, new)
    tr = Trace()
    real_analyze = opensees_shim.analyze

    def analyze(n=1):
        rc = real_analyze(n)
        if rc == 0:
            d = opensees_shim._D
            tags = sorted(d.elements)
            tr.I.append([d.elements[t][4] for t in tags])
            tr.M.append([float(d.ele_forces[t][2]) for t in tags])
            tr.V.append([float(d.ele_forces[t][1]) for t in tags])
        return rc

    opensees_shim.analyze = analyze
    ns = {"__name__": "_reference_frame", "__file__": path}
    try:
        random.seed(seed)
        with contextlib.redirect_stdout(io.StringIO()), record_losses(tr.loss):
            exec(compile(text, path, "exec"), ns)
    finally:
        opensees_shim.analyze = real_analyze
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    return ns, tr


# ------------------------------------------------------------------------------------------------
# Trainer data block (SURVEY 8f row 3 and the loader of 3.4): OpenPyStruct_PINN_MultiCase.py from its first line up to
# "# Convert to PyTorch Tensors" (:1-369) -- hyper-parameters, pad_sequences / unify_label_with_c / fit_transform_3d /
# merge_sub_features, json.load of "StructDataLite.json", padding, grouping by n_cases, the permutation split, the
# StandardScalers and the label aggregation -- executed verbatim in a directory that holds a dataset written by the
# product (dataset.save_json) under the file name the trainer opens.
# ------------------------------------------------------------------------------------------------
TRAINER_SCRIPT = "OpenPyStruct_PINN_MultiCase.py"


def trainer_reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_DIR, TRAINER_SCRIPT))


def run_reference_trainer_block(dataset_json: str, seed: int, overrides=None) -> dict:
    """np.random.seed(seed) (the script relies on the global numpy state for its split, :260); the trainer's own source
    up to the tensors.  ``overrides``: text-level replacements of its constants, e.g. {"n_cases = 6": "n_cases = 4"}.
    Returns its namespace (X_train_flat, X_val_flat, Y_train_std, Y_val_std, train_idx, val_idx, scalers_*, ...)."""
    import contextlib
    import io
    import shutil
    import tempfile
    import numpy as np
    saved = {k: sys.modules.get(k) for k in ("matplotlib", "matplotlib.pyplot", "matplotlib.cm", "matplotlib.lines",
                                             "matplotlib.patches", "seaborn")}
    mpl = types.ModuleType("matplotlib")
    mpl.__path__ = []
    for sub in ("pyplot", "cm", "lines", "patches"):
        m = _Anything()
        setattr(mpl, sub, m)
        sys.modules["matplotlib." + sub] = m
    sys.modules["matplotlib"] = mpl
    sys.modules["seaborn"] = _Anything()
    path = os.path.join(REFERENCE_DIR, TRAINER_SCRIPT)
    with open(path, "r") as fh:
        text = fh.read()
    text = text[:text.index("# Convert to PyTorch Tensors")]
    for old, new in (overrides or {}).items():
        if old not in text:
            raise KeyError(old)
        text = text.replace(old, new)
    ns = {"__name__": "_reference_trainer", "__file__": path}
    cwd = os.getcwd()
    with tempfile.TemporaryDirectory() as td:
        shutil.copyfile(dataset_json, os.path.join(td, "StructDataLite.json"))
        os.chdir(td)
        try:
            np.random.seed(seed)
            with contextlib.redirect_stdout(io.StringIO()):
                exec(compile(text, path, "exec"), ns)
        finally:
            os.chdir(cwd)
            for k, v in saved.items():
                if v is None:
                    sys.modules.pop(k, None)
                else:
                    sys.modules[k] = v
    return ns
