"""``torch.ops.openpystruct.beam_opt`` / ``beam_solve``: the C ABI exposed as PyTorch custom ops.

PyTorch is plumbing here (device memory, streams); the op bodies only hand raw pointers and the
current CUDA stream to ``ops_beamopt_launch`` / ``ops_beamsolve_launch``.  Registered for CUDA
only -- calling them with CPU tensors raises (there is no CPU implementation of this path).
"""
from __future__ import annotations

import ctypes as C
from typing import List, Sequence, Tuple

import torch

from . import _cabi
from .params import BeamOptParams

_INT_FIELDS = ("num_nodes", "num_cases", "max_forces", "max_epochs", "patience", "early_stop", "zero_last_node",
               "solver")
_F64_FIELDS = ("E", "G", "udl", "I0", "lr", "gamma", "alpha_moment", "alpha_shear", "tolerance",
               "shear_k", "bending_eps", "clamp_min", "beta1", "beta2", "adam_eps")


def pack_params(p: BeamOptParams) -> Tuple[List[int], List[float]]:
    cp = _cabi.to_c_params(p)
    return [int(getattr(cp, f)) for f in _INT_FIELDS], [float(getattr(cp, f)) for f in _F64_FIELDS]


def _c_params(iparams: Sequence[int], fparams: Sequence[float]) -> _cabi.OpsBeamOptParams:
    cp = _cabi.OpsBeamOptParams()
    cp.struct_size = C.sizeof(_cabi.OpsBeamOptParams)
    for f, v in zip(_INT_FIELDS, iparams):
        setattr(cp, f, int(v))
    for f, v in zip(_F64_FIELDS, fparams):
        setattr(cp, f, float(v))
    return cp


def _check_cuda(*tensors):
    dev = tensors[0].device
    for t in tensors:
        if not t.is_cuda or t.device != dev:
            raise RuntimeError("openpystruct ops need CUDA tensors on one device (no CPU fallback)")
        if not t.is_contiguous():
            raise RuntimeError("openpystruct ops need contiguous tensors")
    return dev


@torch.library.custom_op("openpystruct::beam_opt", mutates_args=(), device_types="cuda")
def beam_opt(fixed_uy: torch.Tensor, force_nodes: torch.Tensor, force_vals: torch.Tensor, L: torch.Tensor,
             schedule: torch.Tensor, iparams: Sequence[int], fparams: Sequence[float]
             ) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor,
                        torch.Tensor, torch.Tensor, torch.Tensor]:
    """fixed_uy u8[B,nn], force_nodes i32[B,C,F], force_vals f64[B,C,F], L f64[B], schedule f32[max_e,2]
    -> I f32[B,n], defl f64[B,C,nn], rot f64[B,C,nn], shear f32[B,C,n], moment f32[B,C,n],
       epochs i32[B], loss f32[B], status i32[B]."""
    dev = _check_cuda(fixed_uy, force_nodes, force_vals, L, schedule)
    cp = _c_params(iparams, fparams)
    B, nn, Cc = L.shape[0], cp.num_nodes, cp.num_cases
    n = nn - 1
    if (fixed_uy.dtype, force_nodes.dtype, force_vals.dtype, L.dtype, schedule.dtype) != \
            (torch.uint8, torch.int32, torch.float64, torch.float64, torch.float32):
        raise RuntimeError("beam_opt: dtypes must be (uint8, int32, float64, float64, float32)")
    if fixed_uy.numel() != B * nn or force_nodes.numel() != B * Cc * cp.max_forces or \
            force_vals.numel() != force_nodes.numel() or schedule.numel() < 2 * cp.max_epochs:
        raise RuntimeError("beam_opt: shape mismatch")
    I = torch.empty((B, n), dtype=torch.float32, device=dev)
    defl = torch.empty((B, Cc, nn), dtype=torch.float64, device=dev)
    rot = torch.empty((B, Cc, nn), dtype=torch.float64, device=dev)
    shear = torch.empty((B, Cc, n), dtype=torch.float32, device=dev)
    moment = torch.empty((B, Cc, n), dtype=torch.float32, device=dev)
    epochs = torch.empty((B,), dtype=torch.int32, device=dev)
    loss = torch.empty((B,), dtype=torch.float32, device=dev)
    status = torch.empty((B,), dtype=torch.int32, device=dev)
    lib = _cabi.lib()
    with torch.cuda.device(dev):
        _cabi.check(lib.ops_set_device(dev.index), "ops_set_device")
        ws_bytes = lib.ops_beamopt_workspace_bytes(C.byref(cp), B)
        ws = torch.empty((max(int(ws_bytes), 1),), dtype=torch.uint8, device=dev)
        stream = torch.cuda.current_stream(dev).cuda_stream
        rc = lib.ops_beamopt_launch(
            C.byref(cp), B, fixed_uy.data_ptr(), force_nodes.data_ptr(), force_vals.data_ptr(),
            L.data_ptr(), schedule.data_ptr(), I.data_ptr(), defl.data_ptr(), rot.data_ptr(),
            shear.data_ptr(), moment.data_ptr(), epochs.data_ptr(), loss.data_ptr(), status.data_ptr(),
            ws.data_ptr(), ws.numel(), stream)
        _cabi.check(rc, "ops_beamopt_launch")
        ws.record_stream(torch.cuda.current_stream(dev))
    return I, defl, rot, shear, moment, epochs, loss, status


@beam_opt.register_fake
def _(fixed_uy, force_nodes, force_vals, L, schedule, iparams, fparams):
    nn, Cc = int(iparams[0]), int(iparams[1])
    B, n = L.shape[0], nn - 1
    f32, f64, i32 = torch.float32, torch.float64, torch.int32
    e = lambda shape, dt: torch.empty(shape, dtype=dt, device=L.device)   # noqa: E731
    return (e((B, n), f32), e((B, Cc, nn), f64), e((B, Cc, nn), f64), e((B, Cc, n), f32),
            e((B, Cc, n), f32), e((B,), i32), e((B,), f32), e((B,), i32))


@torch.library.custom_op("openpystruct::beam_solve", mutates_args=(), device_types="cuda")
def beam_solve(fixed_uy: torch.Tensor, force_nodes: torch.Tensor, force_vals: torch.Tensor, L: torch.Tensor,
               I: torch.Tensor, iparams: Sequence[int], fparams: Sequence[float]
               ) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor]:
    """One FP64 solve per beam: I f64[B,n] -> defl f64[B,nn], rot f64[B,nn], shear f64[B,n], moment f64[B,n], status."""
    dev = _check_cuda(fixed_uy, force_nodes, force_vals, L, I)
    cp = _c_params(iparams, fparams)
    B, nn = L.shape[0], cp.num_nodes
    n = nn - 1
    if I.dtype != torch.float64 or I.numel() != B * n:
        raise RuntimeError("beam_solve: I must be float64 [B, n]")
    defl = torch.empty((B, nn), dtype=torch.float64, device=dev)
    rot = torch.empty((B, nn), dtype=torch.float64, device=dev)
    shear = torch.empty((B, n), dtype=torch.float64, device=dev)
    moment = torch.empty((B, n), dtype=torch.float64, device=dev)
    status = torch.empty((B,), dtype=torch.int32, device=dev)
    lib = _cabi.lib()
    with torch.cuda.device(dev):
        _cabi.check(lib.ops_set_device(dev.index), "ops_set_device")
        stream = torch.cuda.current_stream(dev).cuda_stream
        rc = lib.ops_beamsolve_launch(
            C.byref(cp), B, fixed_uy.data_ptr(), force_nodes.data_ptr(), force_vals.data_ptr(),
            L.data_ptr(), I.data_ptr(), defl.data_ptr(), rot.data_ptr(), shear.data_ptr(),
            moment.data_ptr(), status.data_ptr(), stream)
        _cabi.check(rc, "ops_beamsolve_launch")
    return defl, rot, shear, moment, status


@beam_solve.register_fake
def _(fixed_uy, force_nodes, force_vals, L, I, iparams, fparams):
    nn = int(iparams[0])
    B, n = L.shape[0], nn - 1
    e = lambda shape, dt: torch.empty(shape, dtype=dt, device=L.device)   # noqa: E731
    return (e((B, nn), torch.float64), e((B, nn), torch.float64), e((B, n), torch.float64),
            e((B, n), torch.float64), e((B,), torch.int32))


_schedule_cache = {}


def device_schedule(p: BeamOptParams, device) -> torch.Tensor:
    key = (p.lr, p.gamma, p.beta1, p.beta2, p.max_e, str(device))
    t = _schedule_cache.get(key)
    if t is None:
        t = torch.from_numpy(_cabi.fill_schedule(p)).to(device)
        _schedule_cache[key] = t
    return t


def optimise_beams(p: BeamOptParams, fixed_uy: torch.Tensor, force_nodes: torch.Tensor,
                   force_vals: torch.Tensor, L: torch.Tensor) -> dict:
    """Functional wrapper: device tensors in, dict of device tensors out."""
    ip, fp = pack_params(p)
    sched = device_schedule(p, L.device)
    names = ("I", "defl", "rot", "shear", "moment", "epochs", "loss", "status")
    return dict(zip(names, torch.ops.openpystruct.beam_opt(fixed_uy, force_nodes, force_vals, L, sched, ip, fp)))


def solve_beams(p: BeamOptParams, fixed_uy, force_nodes, force_vals, L, I) -> dict:
    ip, fp = pack_params(p.replace(num_cases=1))
    names = ("defl", "rot", "shear", "moment", "status")
    return dict(zip(names, torch.ops.openpystruct.beam_solve(fixed_uy, force_nodes, force_vals, L, I, ip, fp)))
