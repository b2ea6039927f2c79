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
