"""Groundwork for SURVEY 8f row 4 (frame optimiser): the OpenSees shim on frames (known answers) and the torch-path
restatement of the reference's frame optimiser against fixtures frozen from the reference's own source."""
import os

import numpy as np
import pytest

from oracle import frameopt_port as fp
from oracle import opensees_shim as ops

HERE = os.path.dirname(os.path.abspath(__file__))


def _model():
    ops.wipe()
    ops.model('basic', '-ndm', 2, '-ndf', 3)
    ops.geomTransf('Linear', 1)


def _analyse(system):
    ops.timeSeries('Linear', 1)
    ops.system(system)
    ops.numberer('RCM')
    ops.constraints('Plain')
    ops.integrator('LoadControl', 1.0)
    ops.algorithm('Newton')
    ops.analysis('Static')
    assert ops.analyze(1) == 0


@pytest.mark.parametrize("system", ["BandGeneral", "BandSPD"])
def test_vertical_cantilever_tip_load(system):
    """Clamped column of height h with a horizontal tip load P: u = P h^3 / 3EI, theta = -P h^2 / 2EI (clockwise),
    base moment P h; the element's global end forces balance the load."""
    E, A, I, h, P = 200e9, 0.02, 5e-4, 3.0, 1e4
    _model()
    ops.node(1, 0.0, 0.0); ops.node(2, 0.0, h)
    ops.fix(1, 1, 1, 1)
    ops.element('elasticBeamColumn', 1, 1, 2, A, E, I, 1)
    ops.pattern('Plain', 1, 1)
    ops.load(2, P, 0.0, 0.0)
    _analyse(system)
    assert ops.nodeDisp(2, 1) == pytest.approx(P * h ** 3 / (3 * E * I), rel=1e-12)
    assert ops.nodeDisp(2, 3) == pytest.approx(-P * h ** 2 / (2 * E * I), rel=1e-12)
    f = ops.eleResponse(1, 'forces')
    assert f[0] == pytest.approx(-P, rel=1e-12) and f[3] == pytest.approx(P, rel=1e-12)
    assert abs(f[2]) == pytest.approx(P * h, rel=1e-12) and abs(f[5]) < 1e-6


def test_symmetric_portal_under_uniform_load():
    """Fixed-base portal, span L, height h, UDL w on the beam, no sway by symmetry: with beta = (I_b/L)/(I_c/h) the knee
    moment is w L^2 / (6 (2 + beta)) and the base moment half of it (slope-deflection; axial shortening made
    negligible with a large area)."""
    E, A, Ic, Ib, h, L, w = 200e9, 1e4, 4e-4, 9e-4, 3.0, 6.0, -1e4
    _model()
    ops.node(1, 0.0, 0.0); ops.node(2, L, 0.0); ops.node(3, 0.0, h); ops.node(4, L, h)
    ops.fix(1, 1, 1, 1); ops.fix(2, 1, 1, 1)
    ops.element('elasticBeamColumn', 1, 1, 3, A, E, Ic, 1)
    ops.element('elasticBeamColumn', 2, 2, 4, A, E, Ic, 1)
    ops.element('elasticBeamColumn', 3, 3, 4, A, E, Ib, 1)
    ops.pattern('Plain', 1, 1)
    ops.eleLoad('-ele', 3, '-type', '-beamUniform', w, 0.0)
    _analyse("BandGeneral")
    beta = (Ib / L) / (Ic / h)
    knee = abs(w) * L ** 2 / (6 * (2 + beta))
    col = ops.eleResponse(1, 'forces')
    assert abs(col[5]) == pytest.approx(knee, rel=1e-6)
    assert abs(col[2]) == pytest.approx(knee / 2, rel=1e-6)
    assert abs(ops.nodeDisp(3, 1)) < 1e-9 and ops.nodeDisp(3, 3) == pytest.approx(-ops.nodeDisp(4, 3), rel=1e-9)
    # vertical equilibrium: the two columns carry w L
    assert col[1] + ops.eleResponse(2, 'forces')[1] == pytest.approx(abs(w) * L, rel=1e-12)


def test_frame_equilibrium_of_the_reference_load_pattern():
    p = fp.FrameParams()
    nodes, elements, n_col = fp.frame_topology(3, 2, p)
    assert fp.build_and_solve(nodes, elements, n_col, [p.I0] * len(elements), p) == 0
    ops.reactions()
    ground = [t for t, (x, y) in nodes.items() if y == 0.0]
    rx = sum(ops.nodeReaction(t, 1) for t in ground)
    ry = sum(ops.nodeReaction(t, 2) for t in ground)
    # two loaded left-hand nodes; three beams per story, two stories; eleLoad's second value is the AXIAL load
    assert rx == pytest.approx(-(2 * p.lateral_load + 6 * p.vertical_load * p.bay_width), rel=1e-9)
    assert ry == pytest.approx(-6 * p.vertical_load * p.bay_width, rel=1e-9)


@pytest.mark.parametrize("key", ["s2", "s4", "s11"])
def test_port_reproduces_the_reference_runs(key):
    g = np.load(os.path.join(HERE, "golden", "frame_goldens.npz"))
    bays, stories, cap, epochs = (int(v) for v in g[key + "_shape"])
    c = g[key + "_consts"]
    p = fp.FrameParams(E=c[0], G=c[1], A=c[2], I0=c[3], alpha_moment=c[4], alpha_shear=c[5], k=c[6], lateral_load=c[7],
                       vertical_load=c[8], lr=c[9], tolerance=c[10], bay_width=c[11], story_height=c[12],
                       patience=int(g[key + "_patience"][0]), num_epochs=cap)
    r = fp.frame_optimise(bays, stories, p)
    assert r["epochs"] == epochs
    assert np.array_equal(r["loss"], g[key + "_loss"])
    assert np.array_equal(r["I"], g[key + "_I"])


@pytest.mark.parametrize("bays,stories", [(1, 1), (1, 2), (4, 5), (8, 9), (10, 10)])
def test_band_restatement_of_the_frame_solve_matches_the_shim(bays, stories):
    """oracle/frame_fe.py (DOFs at the elevated nodes only, closed-form global element matrices, band Cholesky) against
    the shim's dense assembly + BandGeneral on random inertias: displacements and the end forces the loss reads."""
    from oracle import frame_fe
    p = fp.FrameParams()
    nodes, elements, n_col = fp.frame_topology(bays, stories, p)
    rng = np.random.default_rng(bays * 100 + stories)
    I = np.exp(rng.uniform(np.log(1e-5), np.log(0.5), len(elements)))
    assert fp.build_and_solve(nodes, elements, n_col, list(I), p) == 0
    u, forces = frame_fe.frame_solve(bays, stories, I, p)
    want_u = np.array([ops.nodeDisp(t) for t in sorted(nodes) if nodes[t][1] != 0.0])
    want_f = np.array([ops.eleResponse(t, 'forces') for t, _, _ in elements])
    assert np.max(np.abs(u - want_u)) <= 1e-9 * np.max(np.abs(want_u))
    scale = np.max(np.abs(want_f))
    assert np.max(np.abs(forces - want_f)) <= 1e-8 * scale


def test_partial_gradient_closed_form_is_what_autograd_computes():
    import torch
    from oracle import frame_fe
    p = fp.FrameParams()
    rng = np.random.default_rng(3)
    n = 40
    I = np.exp(rng.uniform(np.log(1e-5), np.log(0.5), n)); M = rng.normal(0, 5e4, n); V = rng.normal(0, 3e4, n)
    It = torch.tensor(I, dtype=torch.float64, requires_grad=True)
    e_m, e_v = 0.0, 0.0
    for e in range(n):
        e_m = e_m + (M[e] ** 2) / (2 * p.E * It[e] + 1e-8)
        e_v = e_v + (V[e] ** 2) / (p.G * (p.k * (It[e] ** 0.5)))
    total = torch.sum(It) + p.alpha_moment * e_m + p.alpha_shear * e_v
    total.backward()
    want_total, want_grad = frame_fe.frame_loss_and_partial_gradient(I, M, V, p)
    assert float(total.detach()) == pytest.approx(want_total, rel=1e-13)
    assert np.allclose(It.grad.numpy(), want_grad, rtol=1e-12, atol=0)
