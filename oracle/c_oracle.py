"""TEST INFRASTRUCTURE ONLY -- ctypes binding of the C oracle (oracle/csrc/beamopt_oracle.c).

The array layout is the product C ABI's (include/openpystruct_b200.h): beam-major, row-major,
0-based node indices in ``force_nodes`` with ``-1`` for unused slots.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "libbeamopt_oracle.so")
_lib = None


class Params(C.Structure):
    _fields_ = [
        ("struct_size", C.c_int32), ("num_nodes", C.c_int32), ("num_cases", C.c_int32),
        ("max_forces", C.c_int32), ("max_epochs", C.c_int32), ("patience", C.c_int32),
        ("early_stop", C.c_int32), ("zero_last_node", C.c_int32), ("solver", C.c_int32),
        ("reserved", C.c_int32),
        ("E", C.c_double), ("G", C.c_double), ("udl", C.c_double), ("I0", C.c_double),
        ("lr", C.c_double), ("gamma", C.c_double), ("alpha_moment", C.c_double),
        ("alpha_shear", C.c_double), ("tolerance", C.c_double), ("shear_k", C.c_double),
        ("bending_eps", C.c_double), ("clamp_min", C.c_double), ("beta1", C.c_double),
        ("beta2", C.c_double), ("adam_eps", C.c_double),
    ]


def make_params(*, num_nodes=101, num_cases=1, max_forces=4, max_epochs=600, patience=5, early_stop=True,
                zero_last_node=False, E=200e9, nu=0.3, udl=-1000.0, I0=0.5, lr=0.01, gamma=0.98,
                alpha_moment=1e-2, alpha_shear=1e-2, tolerance=5e-3, shear_k=0.03, bending_eps=1e-6,
                clamp_min=1e-8, beta1=0.9, beta2=0.999, adam_eps=1e-8) -> Params:
    return Params(C.sizeof(Params), num_nodes, num_cases, max_forces, max_epochs, patience,
                  int(early_stop), int(zero_last_node), 0, 0, E, E / (2 * (1 + nu)), udl, I0, lr, gamma,
                  alpha_moment, alpha_shear, tolerance, shear_k, bending_eps, clamp_min, beta1, beta2, adam_eps)


def build(force: bool = False) -> str:
    src = [os.path.join(_HERE, "csrc", f) for f in ("beamopt_oracle.c", "beam_fe.inc", "beamopt_oracle.h")]
    stale = (not os.path.exists(_LIB_PATH)) or any(
        os.path.exists(s) and os.path.getmtime(s) > os.path.getmtime(_LIB_PATH) for s in src)
    if force or stale:
        subprocess.run(["make", "-C", _HERE, "-B" if force else "-s"], check=True,
                       stdout=subprocess.DEVNULL)
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_LIB_PATH)
        _lib.oracle_torch_sum_f32.restype = C.c_float
        _lib.oracle_loss_grad_f32.restype = C.c_float
    return _lib


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


def torch_sum(x: np.ndarray) -> np.float32:
    x = np.ascontiguousarray(x, dtype=np.float32)
    return np.float32(lib().oracle_torch_sum_f32(_p(x, C.c_float), C.c_int64(x.size)))


def adam_schedule(p: Params) -> np.ndarray:
    t = np.zeros((max(p.max_epochs, 1), 2), dtype=np.float32)
    lib().oracle_adam_schedule(C.byref(p), _p(t, C.c_float))
    return t


def loss_grad(p: Params, I, csq, hsq):
    I = np.ascontiguousarray(I, np.float32)
    csq = np.ascontiguousarray(csq, np.float32)
    hsq = np.ascontiguousarray(hsq, np.float32)
    n = I.size
    g = np.zeros(n, np.float32)
    scr = np.zeros(2 * n, np.float32)
    l = lib().oracle_loss_grad_f32(C.byref(p), C.c_int64(n), _p(I, C.c_float), _p(csq, C.c_float),
                                   _p(hsq, C.c_float), _p(g, C.c_float), _p(scr, C.c_float))
    return np.float32(l), g


def adam_step(p: Params, neg_step, bc2_sqrt, grad, I, m, v):
    """In place on I, m, v (float32 contiguous)."""
    lib().oracle_adam_step_f32(C.byref(p), C.c_int64(I.size), C.c_float(neg_step), C.c_float(bc2_sqrt),
                               _p(np.ascontiguousarray(grad, np.float32), C.c_float),
                               _p(I, C.c_float), _p(m, C.c_float), _p(v, C.c_float))


def _check_inputs(p, fixed_uy, force_nodes, force_vals, L):
    B = L.shape[0]
    nn, Cc, F = p.num_nodes, p.num_cases, p.max_forces
    fixed_uy = np.ascontiguousarray(fixed_uy, np.uint8).reshape(B, nn)
    force_nodes = np.ascontiguousarray(force_nodes, np.int32).reshape(B, Cc, F)
    force_vals = np.ascontiguousarray(force_vals, np.float64).reshape(B, Cc, F)
    L = np.ascontiguousarray(L, np.float64)
    return B, fixed_uy, force_nodes, force_vals, L


def beamopt(p: Params, fixed_uy, force_nodes, force_vals, L, fe_precision: int = 0) -> dict:
    """fe_precision 0 = FP64 dpbsv restatement (the reference's arithmetic), 1 = 80-bit FE solve."""
    L = np.asarray(L, np.float64).reshape(-1)
    B, fixed_uy, force_nodes, force_vals, L = _check_inputs(p, fixed_uy, force_nodes, force_vals, L)
    nn, Cc = p.num_nodes, p.num_cases
    n = nn - 1
    out = {
        "I": np.zeros((B, n), np.float32), "defl": np.zeros((B, Cc, nn)), "rot": np.zeros((B, Cc, nn)),
        "shear": np.zeros((B, Cc, n), np.float32), "moment": np.zeros((B, Cc, n), np.float32),
        "epochs": np.zeros(B, np.int32), "loss": np.zeros(B, np.float32), "status": np.zeros(B, np.int32),
        # smallest early-stop decision margin of the run in fp32 ulps of the loss (oracle-side diagnostic, no product
        # counterpart): a stop epoch may legitimately differ only where this is about one ulp
        "margin_ulps": np.full(B, np.inf),
    }
    rc = lib().oracle_beamopt_margin(C.byref(p), C.c_int64(B), _p(fixed_uy, C.c_uint8), _p(force_nodes, C.c_int32),
                                     _p(force_vals, C.c_double), _p(L, C.c_double), _p(out["I"], C.c_float),
                                     _p(out["defl"], C.c_double), _p(out["rot"], C.c_double),
                                     _p(out["shear"], C.c_float), _p(out["moment"], C.c_float),
                                     _p(out["epochs"], C.c_int32), _p(out["loss"], C.c_float),
                                     _p(out["status"], C.c_int32), C.c_int(fe_precision),
                                     _p(out["margin_ulps"], C.c_double))
    if rc != 0:
        raise RuntimeError(f"oracle_beamopt rc={rc}")
    return out


def beam_solve(p: Params, fixed_uy, force_nodes, force_vals, L, I, precision: int = 0) -> dict:
    """One solve per beam (single load case); precision 0 = FP64 dpbsv restatement, 1 = 80-bit."""
    assert p.num_cases == 1
    L = np.asarray(L, np.float64).reshape(-1)
    B, fixed_uy, force_nodes, force_vals, L = _check_inputs(p, fixed_uy, force_nodes, force_vals, L)
    nn = p.num_nodes
    n = nn - 1
    I = np.ascontiguousarray(I, np.float64).reshape(B, n)
    out = {"defl": np.zeros((B, nn)), "rot": np.zeros((B, nn)), "shear": np.zeros((B, n)),
           "moment": np.zeros((B, n))}
    rc = lib().oracle_beam_solve(C.byref(p), C.c_int64(B), _p(fixed_uy, C.c_uint8), _p(force_nodes, C.c_int32),
                                 _p(force_vals, C.c_double), _p(L, C.c_double), _p(I, C.c_double),
                                 _p(out["defl"], C.c_double), _p(out["rot"], C.c_double),
                                 _p(out["shear"], C.c_double), _p(out["moment"], C.c_double), C.c_int(precision))
    out["rc"] = rc
    return out


def pack_cases(num_nodes: int, max_forces: int, cases) -> tuple:
    """cases: list of (L, roller_tags(1-based), force_tags(1-based), force_values) -> ABI arrays (C = 1)."""
    B = len(cases)
    fixed = np.zeros((B, num_nodes), np.uint8)
    fn = -np.ones((B, 1, max_forces), np.int32)
    fv = np.zeros((B, 1, max_forces), np.float64)
    L = np.zeros(B)
    for b, (Lb, rollers, ftags, fvals) in enumerate(cases):
        L[b] = Lb
        fixed[b, 0] = 1
        for t in rollers:
            fixed[b, t - 1] = 1
        for j, (t, F) in enumerate(zip(ftags, fvals)):
            fn[b, 0, j] = t - 1
            fv[b, 0, j] = F
    return fixed, fn, fv, L
