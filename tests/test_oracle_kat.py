"""Known-answer tests pinning the CPU oracle's FE half to closed-form Euler-Bernoulli results
(the reference ships no tests).  Covers the OpenSees shim, the torch-path port and
the C restatement (FP64 and 80-bit)."""
import numpy as np
import pytest

from oracle import c_oracle, opensees_shim as ops
from oracle.beamopt_port import fe_solve

E, I0, L, n, w = 200e9, 0.5, 200.0, 100, -1000.0


def shim_solve(rollers, loads, udl, inertia=I0):
    ops.wipe()
    ops.model('basic', '-ndm', 2, '-ndf', 3)
    for i, x in enumerate(np.linspace(0, L, n + 1)):
        ops.node(i + 1, x, 0.0)
    ops.fix(1, 1, 1, 0)
    for r in rollers:
        ops.fix(r, 0, 1, 0)
    ops.geomTransf('Linear', 1)
    for e in range(n):
        ops.element('elasticBeamColumn', e + 1, e + 1, e + 2, 0.01, E, inertia, 1)
    ops.timeSeries('Linear', 1)
    ops.pattern('Plain', 1, 1)
    for nd, F in loads:
        ops.load(nd, 0.0, F, 0.0)
    if udl:
        for e in range(1, n + 1):
            ops.eleLoad('-ele', e, '-type', '-beamUniform', udl, udl)
    ops.system('BandSPD'); ops.numberer('RCM'); ops.constraints('Plain')
    ops.integrator('LoadControl', 1.0); ops.algorithm('Linear'); ops.analysis('Static')
    assert ops.analyze(1) == 0
    uy = np.array([ops.nodeDisp(i, 2) for i in range(1, n + 2)])
    th = np.array([ops.nodeDisp(i, 3) for i in range(1, n + 2)])
    V = np.array([ops.eleResponse(e, 'forces')[1] for e in range(1, n + 1)])
    M = np.array([ops.eleResponse(e, 'forces')[2] for e in range(1, n + 1)])
    return uy, th, V, M


def port_solve(rollers, loads, udl):
    fixed = np.zeros(n + 1, bool); fixed[0] = True
    for r in rollers:
        fixed[r - 1] = True
    f = np.zeros(n + 1)
    for nd, F in loads:
        f[nd - 1] += F
    return fe_solve(np.full(n, I0), L, fixed, f, udl, E)


def c_solve(rollers, loads, udl, precision):
    p = c_oracle.make_params(udl=udl, max_forces=4)
    fixed, fn, fv, Ls = c_oracle.pack_cases(n + 1, 4, [(L, rollers, [nd for nd, _ in loads], [F for _, F in loads])])
    o = c_oracle.beam_solve(p, fixed, fn[:, 0], fv[:, 0], Ls, np.full((1, n), I0), precision)
    assert o["rc"] == 0
    return o["defl"][0], o["rot"][0], o["shear"][0], o["moment"][0]


SOLVERS = {
    "shim": lambda r, l, u: shim_solve(r, l, u),
    "port": port_solve,
    "c_f64": lambda r, l, u: c_solve(r, l, u, 0),
    "c_f80": lambda r, l, u: c_solve(r, l, u, 1),
}


@pytest.mark.parametrize("name", SOLVERS)
def test_simply_supported_udl(name):
    uy, th, V, M = SOLVERS[name]([n + 1], [], w)
    tol = 1e-12 if name == "c_f80" else 2e-9
    assert uy[50] == pytest.approx(5 * w * L ** 4 / (384 * E * I0), rel=tol)
    assert th[0] == pytest.approx(w * L ** 3 / (24 * E * I0), rel=tol)
    assert V[0] == pytest.approx(-w * L / 2, rel=tol)          # global Fy at node-i end, up positive
    assert M[50] == pytest.approx(w * L ** 2 / 8, rel=tol)     # Mz at node-i end = -(sagging moment)
    assert abs(M[0]) < 1e-3


@pytest.mark.parametrize("name", SOLVERS)
def test_midspan_point_load(name):
    P = -1e5
    uy, th, V, M = SOLVERS[name]([n + 1], [(51, P)], 0.0)
    tol = 1e-12 if name == "c_f80" else 2e-9
    assert uy[50] == pytest.approx(P * L ** 3 / (48 * E * I0), rel=tol)
    assert M[50] == pytest.approx(P * L / 4, rel=tol)
    assert V[49] == pytest.approx(-P / 2, rel=tol)
    assert V[50] == pytest.approx(P / 2, rel=tol)


@pytest.mark.parametrize("name", SOLVERS)
def test_default_bridge_overhang_is_statically_determinate(name):
    # rollers [10,30,70,85,100] (1-based): node 101 is a free tip; last element carries only its own UDL
    uy, th, V, M = SOLVERS[name]([10, 30, 70, 85, 100], [], w)
    Le = L / n
    assert V[-1] == pytest.approx(-w * Le, rel=1e-8)
    assert M[-1] == pytest.approx(-w * Le ** 2 / 2, rel=1e-8)
    for r in (1, 10, 30, 70, 85, 100):
        assert uy[r - 1] == 0.0


@pytest.mark.parametrize("name", ["port", "c_f64"])
def test_equilibrium_and_agreement_with_shim(name):
    loads = [(20, -3e5), (55, -1e5), (77, -2.2e5)]
    ref = shim_solve([10, 30, 70, 85, 100], loads, w)
    got = SOLVERS[name]([10, 30, 70, 85, 100], loads, w)
    for a, b in zip(got, ref):
        assert np.max(np.abs(a - b)) <= 1e-10 * np.max(np.abs(b))
    # nodal equilibrium from the element end forces: R_i = Fy_i(e=i) + Fy_j(e=i-1) - P_i with
    # Fy_j(e) = -V_e - w*Le; reactions vanish at free nodes and carry the whole load at supports
    uy, th, V, M = got
    Le = L / n
    P = np.zeros(n + 1)
    for nd, F in loads:
        P[nd - 1] += F
    R = np.zeros(n + 1)
    R[:-1] += V
    R[1:] += -V - w * Le
    R -= P
    supports = [0, 9, 29, 69, 84, 99]
    free = np.setdiff1d(np.arange(n + 1), supports)
    scale = abs(sum(F for _, F in loads) + w * L)
    assert np.max(np.abs(R[free])) < 1e-8 * scale
    assert R[supports].sum() == pytest.approx(-(sum(F for _, F in loads) + w * L), rel=1e-9)


def test_f64_vs_extended_precision_default_bridge():
    rng = np.random.default_rng(3)
    p = c_oracle.make_params()
    cases = [(L, [10, 30, 70, 85, 100], [15 + 3 * k, 60 + k], [-2e5, -1e5]) for k in range(8)]
    fixed, fn, fv, Ls = c_oracle.pack_cases(n + 1, 4, cases)
    I = np.exp(rng.uniform(np.log(3e-3), np.log(0.9), (8, n))).astype(np.float32).astype(np.float64)
    a = c_oracle.beam_solve(p, fixed, fn[:, 0], fv[:, 0], Ls, I, 0)
    b = c_oracle.beam_solve(p, fixed, fn[:, 0], fv[:, 0], Ls, I, 1)
    for k in ("defl", "rot", "shear", "moment"):
        err = np.max(np.abs(a[k] - b[k]), axis=1) / np.max(np.abs(b[k]), axis=1)
        assert err.max() < 1e-9, (k, err.max())
