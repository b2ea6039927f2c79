"""TEST INFRASTRUCTURE ONLY -- groundwork for SURVEY 8f row 4 (frame optimiser), not used by the product.

The FE half of the reference's frame optimiser (OpenPyStruct_FrameOpt_Discrete_Beta.py:75-139 through OpenSees)
written the way a per-frame CUDA kernel would do it, as the algorithmic specification for that kernel:

* degrees of freedom only at the elevated nodes, (ux, uy, rz) per node in tag order, so the stiffness is SPD and
  banded with half bandwidth 3 (num_bays + 1) + 2 (a column couples a node with the one a story above);
* closed-form GLOBAL element matrices for the two orientations that occur (beams along x, columns along y) --
  no rotation matrices at run time;
* equivalent nodal loads of ``eleLoad -beamUniform Wy Wx`` (the reference passes the same value for the
  transverse AND the axial distributed load, :138);
* banded Cholesky (``dpbsv``) for ``K u = f``; the reference's ``BandGeneral`` LU solves the same SPD system;
* global end forces ``[Fx_i, Fy_i, Mz_i]`` = ``eleResponse(e,'forces')[0:3]`` incl. the fixed-end terms, of which
  the loss reads ``[1]`` ("shear": for a COLUMN this is its axial force) and ``[2]``;
* the explicit partial gradient of the loss with M, V constant (what autograd sees, :148-166).

Checked against oracle/opensees_shim.py and against torch autograd in tests/test_oracle_frame.py.
"""
from __future__ import annotations

import numpy as np
from scipy.linalg import solveh_banded

from .frameopt_port import FrameParams, frame_topology


def _k_beam(E, A, I, L):
    """Global stiffness of a member along +x, DOF order (ux_i, uy_i, rz_i, ux_j, uy_j, rz_j)."""
    a, b, c, d, t = E * A / L, 12 * E * I / L ** 3, 6 * E * I / L ** 2, 4 * E * I / L, 2 * E * I / L
    return np.array([[a, 0, 0, -a, 0, 0],
                     [0, b, c, 0, -b, c],
                     [0, c, d, 0, -c, t],
                     [-a, 0, 0, a, 0, 0],
                     [0, -b, -c, 0, b, -c],
                     [0, c, t, 0, -c, d]], float)


def _k_column(E, A, I, L):
    """Global stiffness of a member along +y: local axial = global y, local transverse = -global x."""
    a, b, c, d, t = E * A / L, 12 * E * I / L ** 3, 6 * E * I / L ** 2, 4 * E * I / L, 2 * E * I / L
    return np.array([[b, 0, -c, -b, 0, -c],
                     [0, a, 0, 0, -a, 0],
                     [-c, 0, d, c, 0, t],
                     [-b, 0, c, b, 0, c],
                     [0, -a, 0, 0, a, 0],
                     [-c, 0, t, c, 0, d]], float)


def frame_solve(num_bays: int, num_stories: int, inertias, p: FrameParams = FrameParams()):
    """-> (u [nodes above ground, 3], forces [elements, 6] global end forces like eleResponse(e,'forces'))."""
    nodes, elements, n_col = frame_topology(num_bays, num_stories, p)
    nb1 = num_bays + 1
    nfree = 3 * num_stories * nb1
    hbw = 3 * nb1 + 2
    ab = np.zeros((hbw + 1, nfree))                       # upper band storage of dpbsv
    f = np.zeros(nfree)

    def dofs(tag):                                        # -1 for the clamped ground nodes
        i = tag - 1 - nb1
        return [3 * i, 3 * i + 1, 3 * i + 2] if i >= 0 else [-1, -1, -1]

    ke, fe, idx = [], [], []
    w = p.vertical_load
    for (tag, ni, nj), I in zip(elements, inertias):
        column = tag <= n_col
        L = p.story_height if column else p.bay_width
        k = _k_column(p.E, p.A, float(I), L) if column else _k_beam(p.E, p.A, float(I), L)
        # equivalent nodal loads of the uniform loads (transverse w and axial w on beams only)
        q = np.zeros(6) if column else np.array([w * L / 2, w * L / 2, w * L * L / 12, w * L / 2, w * L / 2, -w * L * L / 12])
        d = dofs(ni) + dofs(nj)
        for r in range(6):
            if d[r] < 0:
                continue
            f[d[r]] += q[r]
            for c in range(6):
                if d[c] >= d[r]:
                    ab[hbw - (d[c] - d[r]), d[c]] += k[r, c]
        ke.append(k); fe.append(q); idx.append(d)
    for tag, (x, y) in nodes.items():
        if x == 0.0 and y != 0.0:
            f[dofs(tag)[0]] += p.lateral_load
    u = solveh_banded(ab, f, lower=False)
    forces = np.zeros((len(elements), 6))
    for e, (k, q, d) in enumerate(zip(ke, fe, idx)):
        ue = np.array([u[i] if i >= 0 else 0.0 for i in d])
        forces[e] = k @ ue - q
    return u.reshape(-1, 3), forces


def frame_loss_and_partial_gradient(I, M, V, p: FrameParams = FrameParams()):
    """total loss and d(total)/dI_e with M, V held constant (FP64 closed form of what autograd computes in fp32)."""
    I = np.asarray(I, float); M = np.asarray(M, float); V = np.asarray(V, float)
    den = 2 * p.E * I + 1e-8
    gk = p.G * p.k
    total = I.sum() + p.alpha_moment * (M ** 2 / den).sum() + p.alpha_shear * (V ** 2 / (gk * np.sqrt(I))).sum()
    grad = 1.0 - p.alpha_moment * M ** 2 * 2 * p.E / den ** 2 - p.alpha_shear * V ** 2 / gk * 0.5 * I ** -1.5
    return total, grad
