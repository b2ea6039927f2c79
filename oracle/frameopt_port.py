"""TEST INFRASTRUCTURE ONLY -- groundwork for SURVEY 8f row 4 (frame optimiser), not used by the product.

Restatement of the reference's single-frame optimiser (OpenPyStruct_FrameOpt_Discrete_Beta.py): a rectangular
2-D frame of ``num_bays`` x ``num_stories`` (:46-70), columns numbered before beams (:104-121), every ground node
clamped (:97-100), a lateral nodal load on the left-hand nodes above ground (:129-131), a uniform load on every
beam (:135-138), BandGeneral / Newton / LoadControl 1.0 (:141-146), loss
``sum(I) + a_m sum M^2 / (2 E I + 1e-8) + a_s sum V^2 / (G k sqrt(I))`` accumulated element by element from Python
floats M, V = eleResponse(e,'forces')[2|1] (:148-166, so autograd sees M and V as constants), torch Adam WITHOUT
learning-rate decay (:174), ``clamp_(1e-8)`` (:188-189), early stop on ``tolerance`` / ``patience`` evaluated after
the step (:194-205).  The FE half runs on oracle/opensees_shim.py; the torch half is real torch, so the fp32
operation order is the reference's.  Checked bit for bit against tests/golden/frame_goldens.npz, which
tests/golden/make_frame_golden.py froze from the reference's own source (tests/test_oracle_frame.py).

PARITY UNPINNED at the OpenSees boundary, like the beam oracle (no OpenSeesPy here, no reference tests).
"""
from __future__ import annotations

import dataclasses
from typing import List, Tuple

import numpy as np
import torch

from . import opensees_shim as ops


@dataclasses.dataclass
class FrameParams:
    E: float = 200e9
    G: float = 200e9 / 2.6
    A: float = 0.02
    I0: float = 5e-4
    alpha_moment: float = 1e-2
    alpha_shear: float = 1e-2
    k: float = 0.03
    lateral_load: float = 1e4
    vertical_load: float = -1e4
    lr: float = 0.005
    tolerance: float = 1e-3
    bay_width: float = 6.0
    story_height: float = 3.0
    patience: int = 10
    num_epochs: int = 5000


def frame_topology(num_bays: int, num_stories: int, p: FrameParams):
    """nodes {tag: (x, y)}, elements [(tag, node_i, node_j)] with the columns first, number of columns."""
    nb1 = num_bays + 1
    nodes = {s * nb1 + b + 1: (b * p.bay_width, s * p.story_height) for s in range(num_stories + 1) for b in range(nb1)}
    elements: List[Tuple[int, int, int]] = []
    for s in range(num_stories):                                  # columns, story by story
        for b in range(nb1):
            elements.append((len(elements) + 1, s * nb1 + b + 1, (s + 1) * nb1 + b + 1))
    n_col = len(elements)
    for s in range(1, num_stories + 1):                           # beams of every elevated story
        for b in range(num_bays):
            elements.append((len(elements) + 1, s * nb1 + b + 1, s * nb1 + b + 2))
    return nodes, elements, n_col


def build_and_solve(nodes, elements, n_col, inertias, p: FrameParams) -> int:
    """One static analysis of the frame with the given element inertias (Python floats)."""
    ops.wipe()
    ops.model('basic', '-ndm', 2, '-ndf', 3)
    ops.geomTransf('Linear', 1)
    for tag, (x, y) in nodes.items():
        ops.node(tag, x, y)
    for tag, (x, y) in nodes.items():
        if y == 0.0:
            ops.fix(tag, 1, 1, 1)
    for (tag, ni, nj), I_e in zip(elements, inertias):
        ops.element('elasticBeamColumn', tag, ni, nj, p.A, p.E, I_e, 1)
    ops.timeSeries('Linear', 1)
    ops.pattern('Plain', 1, 1)
    for tag, (x, y) in nodes.items():
        if x == 0.0 and y != 0.0:
            ops.load(tag, p.lateral_load, 0.0, 0.0)
    for tag, _, _ in elements[n_col:]:
        ops.eleLoad('-ele', tag, '-type', '-beamUniform', p.vertical_load, p.vertical_load)
    ops.system('BandGeneral')
    ops.numberer('RCM')
    ops.constraints('Plain')
    ops.integrator('LoadControl', 1.0)
    ops.algorithm('Newton')
    ops.analysis('Static')
    return ops.analyze(1)


def frame_optimise(num_bays: int, num_stories: int, p: FrameParams = FrameParams()) -> dict:
    nodes, elements, n_col = frame_topology(num_bays, num_stories, p)
    n = len(elements)
    I = torch.tensor([p.I0] * n, dtype=torch.float32, requires_grad=True)
    opt = torch.optim.Adam([I], lr=p.lr)
    losses, best, stall = [], float('inf'), 0
    for _ in range(p.num_epochs):
        opt.zero_grad()
        build_and_solve(nodes, elements, n_col, [I[e].item() for e in range(n)], p)
        e_m, e_v = 0.0, 0.0
        for e in range(n):
            f = ops.eleResponse(e + 1, 'forces')
            I_e = I[e]
            e_m = e_m + (f[2] ** 2) / (2 * p.E * I_e + 1e-8)
            e_v = e_v + (f[1] ** 2) / (p.G * (p.k * (I_e ** 0.5)))
        total = torch.sum(I) + p.alpha_moment * e_m + p.alpha_shear * e_v
        total.backward()
        opt.step()
        with torch.no_grad():
            I.clamp_(min=1e-8)
        cur = total.item()
        losses.append(cur)
        if cur < best - p.tolerance:
            best, stall = cur, 0
        else:
            stall += 1
        if stall >= p.patience:
            break
    return {"I": I.detach().numpy().copy(), "loss": np.array(losses, np.float64), "epochs": len(losses), "best": best,
            "num_elements": n, "num_columns": n_col}
