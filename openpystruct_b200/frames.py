"""Host mirror of the reference's single-frame optimiser (OpenPyStruct_FrameOpt_Discrete_Beta.py), batched.

The reference script draws ONE rectangular frame (``num_bays``, ``num_stories`` by ``random.randint``, :50-51),
optimises the moments of inertia of its columns and beams (:179-206) and plots.  Here the same loop runs as one CUDA
launch over a batch of frames behind ``ops_frameopt_launch`` (include/openpystruct_b200.h); the constants keep the
script's names, the draw order is the script's, and the outputs are the script's ``opt_I``, ``loss_history`` and
``best_loss`` per frame.  There is no CPU implementation of this path.
"""
from __future__ import annotations

import ctypes as C
import dataclasses
import random
from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _cabi


@dataclasses.dataclass(frozen=True)
class FrameOptParams:
    """Module-level constants of the script (:14-44) under their reference names."""
    max_bays: int = 10
    max_stories: int = 10
    bay_width: float = 6.0
    story_height: float = 3.0
    E: float = 200e9
    nu: float = 0.3
    A: float = 0.02
    I0: float = 5e-4
    alpha_moment: float = 1e-2
    alpha_shear: float = 1e-2
    k: float = 0.03
    lateral_load: float = 1e4
    vertical_load: float = -1e4
    num_epochs: int = 5000
    lr: float = 0.005
    tolerance: float = 1e-3
    patience: int = 10
    bending_eps: float = 1e-8        # 2 * E * I_val + 1e-8 (:155)
    clamp_min: float = 1e-8          # I_tensor.clamp_(min=1e-8) (:189)
    beta1: float = 0.9
    beta2: float = 0.999
    adam_eps: float = 1e-8
    early_stop: bool = True          # False: exactly num_epochs epochs

    @property
    def G(self) -> float:
        return self.E / (2 * (1 + self.nu))

    def replace(self, **kw) -> "FrameOptParams":
        return dataclasses.replace(self, **kw)


def to_c_params(p: FrameOptParams) -> "_cabi.OpsFrameOptParams":
    return _cabi.OpsFrameOptParams(
        C.sizeof(_cabi.OpsFrameOptParams), p.max_bays, p.max_stories, p.num_epochs, p.patience, int(p.early_stop),
        p.E, p.G, p.A, p.I0, p.alpha_moment, p.alpha_shear, p.k, p.bending_eps, p.lateral_load, p.vertical_load, p.lr,
        p.tolerance, p.bay_width, p.story_height, p.clamp_min, p.beta1, p.beta2, p.adam_eps)


def draw_frame(p: FrameOptParams = FrameOptParams(), rng=random) -> Tuple[int, int]:
    """(num_bays, num_stories) in the script's draw order (:50-51)."""
    num_bays = rng.randint(1, p.max_bays)
    num_stories = rng.randint(1, p.max_stories)
    return num_bays, num_stories


def frame_counts(num_bays: int, num_stories: int) -> Tuple[int, int]:
    """(num_columns, num_beams) (:66-70); members are numbered columns first."""
    return num_stories * (num_bays + 1), num_stories * num_bays


def max_elements(p: FrameOptParams) -> int:
    return sum(frame_counts(p.max_bays, p.max_stories))


def fill_schedule(p: FrameOptParams) -> np.ndarray:
    cp = to_c_params(p)
    table = np.zeros((max(p.num_epochs, 1), 2), np.float32)
    _cabi.check(_cabi.lib().ops_frameopt_fill_schedule(C.byref(cp), table.ctypes.data), "ops_frameopt_fill_schedule")
    return table


_schedules = {}


def optimise_frames_device(p: FrameOptParams, num_bays: torch.Tensor, num_stories: torch.Tensor) -> dict:
    """Device tensors in (int32 [B] on one CUDA device), dict of device tensors out; stream-ordered, no sync."""
    if not (num_bays.is_cuda and num_stories.is_cuda and num_bays.device == num_stories.device):
        raise RuntimeError("openpystruct_b200 frame optimiser needs CUDA tensors on one device (no CPU fallback)")
    if num_bays.dtype != torch.int32 or num_stories.dtype != torch.int32 or num_bays.shape != num_stories.shape:
        raise RuntimeError("num_bays / num_stories must be int32 tensors of one shape")
    dev, B = num_bays.device, int(num_bays.numel())
    cp, me, ne = to_c_params(p), max_elements(p), max(p.num_epochs, 1)
    key = (p.lr, p.beta1, p.beta2, p.num_epochs, str(dev))
    sched = _schedules.get(key)
    if sched is None:
        sched = _schedules[key] = torch.from_numpy(fill_schedule(p)).to(dev)
    out = {"I": torch.empty((B, me), dtype=torch.float32, device=dev),
           "loss_history": torch.empty((B, ne), dtype=torch.float32, device=dev),
           "moment": torch.empty((B, me), dtype=torch.float64, device=dev),
           "shear": torch.empty((B, me), dtype=torch.float64, device=dev),
           "best_loss": torch.empty((B,), dtype=torch.float64, device=dev),
           "epochs": torch.empty((B,), dtype=torch.int32, device=dev),
           "status": torch.empty((B,), dtype=torch.int32, device=dev)}
    lib = _cabi.lib()
    with torch.cuda.device(dev):
        _cabi.check(lib.ops_set_device(dev.index), "ops_set_device")
        rc = lib.ops_frameopt_launch(
            C.byref(cp), B, num_bays.contiguous().data_ptr(), num_stories.contiguous().data_ptr(), sched.data_ptr(),
            out["I"].data_ptr(), out["loss_history"].data_ptr(), out["moment"].data_ptr(), out["shear"].data_ptr(),
            out["best_loss"].data_ptr(), out["epochs"].data_ptr(), out["status"].data_ptr(),
            torch.cuda.current_stream(dev).cuda_stream)
        _cabi.check(rc, "ops_frameopt_launch")
    return out


def optimise_frames(frames: Sequence[Tuple[int, int]], p: FrameOptParams = FrameOptParams(), device="cuda") -> List[dict]:
    """One record per frame: the script's results under its names (``opt_I``, ``loss_history``, ``best_loss``) plus the
    frame, the epoch count and the member forces of the last analysis."""
    dev = torch.device(device)
    if dev.type != "cuda" or not torch.cuda.is_available():
        raise RuntimeError("openpystruct_b200 runs the frame optimiser on a CUDA device only (no CPU fallback)")
    if dev.index is None:
        dev = torch.device("cuda", torch.cuda.current_device())
    nb = torch.tensor([f[0] for f in frames], dtype=torch.int32, device=dev)
    ns = torch.tensor([f[1] for f in frames], dtype=torch.int32, device=dev)
    out = {k: v.cpu().numpy() for k, v in optimise_frames_device(p, nb, ns).items()}
    records = []
    for i, (bays, stories) in enumerate(frames):
        n_col, n_beam = frame_counts(bays, stories)
        n, ep = n_col + n_beam, int(out["epochs"][i])
        records.append({"num_bays": bays, "num_stories": stories, "num_columns": n_col, "num_beams": n_beam,
                        "opt_I": out["I"][i, :n].copy(), "loss_history": out["loss_history"][i, :ep].astype(np.float64),
                        "best_loss": float(out["best_loss"][i]), "epochs": ep, "status": int(out["status"][i]),
                        "bending_moments": out["moment"][i, :n].copy(), "shear_forces": out["shear"][i, :n].copy()})
    return records


def optimise_frame(num_bays: Optional[int] = None, num_stories: Optional[int] = None,
                   p: FrameOptParams = FrameOptParams(), device="cuda", rng=random) -> dict:
    """The script's run for one frame; without a frame it is drawn like the script does (global ``random``)."""
    if num_bays is None or num_stories is None:
        num_bays, num_stories = draw_frame(p, rng)
    return optimise_frames([(num_bays, num_stories)], p, device)[0]
