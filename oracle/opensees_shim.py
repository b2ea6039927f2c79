"""TEST INFRASTRUCTURE ONLY -- CPU oracle, never imported by the product package.

A drop-in stand-in for the subset of ``openseespy.opensees`` that the reference's
beam-optimisation generators cross.  OpenSeesPy itself is a
third-party dependency of the reference (environment.yml:13-14, un-pinned) and is
not installable in this image, so its published algorithm is restated here:

* ``ElasticBeam2d`` (Euler-Bernoulli, no shear deformation) basic stiffness
  ``[[EA/L,0,0],[0,4EI/L,2EI/L],[0,2EI/L,4EI/L]]``,
* ``LinearCrdTransf2d`` basic<->global transformation,
* ``-beamUniform`` element load -> fixed-end forces ``q0 = (-wa*L/2, -wt*L^2/12, +wt*L^2/12)``
  and ``p0 = (-wa*L, -wt*L/2, -wt*L/2)``,
* ``Plain`` constraint handler (constrained equations dropped),
* ``BandSPD`` (LAPACK ``dpbsv`` -- reached here through ``scipy.linalg.solveh_banded``),
* ``LoadControl 1.0`` + ``Linear`` algorithm: one linear solve at load factor 1.0,
* ``eleResponse(e, 'forces')``: global resisting force of the element, fixed-end terms included.

* ``BandGeneral`` (LAPACK ``dgbsv`` through ``scipy.linalg.solve_banded``) for the frame optimiser
  (OpenPyStruct_FrameOpt_Discrete_Beta.py:75-139; elements of any orientation, grounded nodes ``fix(tag,1,1,1)``).

Call sites restated (reference file:line):
  SingleCore:93-124 (setup_model), :176-190 (wipe/analysis/analyze/eleResponse), :224-232 (nodeDisp)
  MultiCore:97-128, :177-191, :222-223 ; GPU:98-129, :138, :185-207, :246-250 ; BeamOpt:95-126, :136-142

PARITY UNPINNED: the reference ships no tests, golden vectors or fixtures for this
boundary and OpenSeesPy cannot run here, so this restatement is pinned only by the
closed-form Euler-Bernoulli known-answer tests in tests/test_oracle_kat.py.

The module keeps OpenSees' process-global domain semantics (one model, ``wipe()`` clears it).
"""
from __future__ import annotations

import numpy as np
from scipy.linalg import solve_banded, solveh_banded

__all__ = [
    "wipe", "model", "node", "fix", "geomTransf", "element", "timeSeries", "pattern",
    "load", "eleLoad", "system", "numberer", "constraints", "integrator", "algorithm",
    "analysis", "analyze", "eleResponse", "nodeDisp", "reactions", "nodeReaction",
]


class _Domain:
    def __init__(self):
        self.ndm = 2
        self.ndf = 3
        self.nodes = {}          # tag -> (x, y)
        self.fixes = {}          # tag -> (fx, fy, fz)
        self.transf = {}         # tag -> type
        self.elements = {}       # tag -> (ni, nj, A, E, I, transf)
        self.nodal_loads = []    # (node, fx, fy, mz)
        self.ele_loads = {}      # ele tag -> [wt, wa] accumulated
        self.system = None
        self.analysis_type = None
        self.disp = None         # dict tag -> np.array(3)
        self.ele_forces = {}     # tag -> np.array(6)
        self.load_factor = 0.0
        self.integrator_dlambda = 1.0


_D = _Domain()


def wipe():
    global _D
    _D = _Domain()


def model(*args):
    # ops.model('basic', '-ndm', 2, '-ndf', 3)   (SingleCore:93)
    a = list(args)
    if "-ndm" in a:
        _D.ndm = int(a[a.index("-ndm") + 1])
    if "-ndf" in a:
        _D.ndf = int(a[a.index("-ndf") + 1])
    if _D.ndm != 2 or _D.ndf != 3:
        raise NotImplementedError("shim supports -ndm 2 -ndf 3 only")


def node(tag, x, y=0.0):
    _D.nodes[int(tag)] = (float(x), float(y))


def fix(tag, fx, fy, fz):
    _D.fixes[int(tag)] = (int(fx), int(fy), int(fz))


def geomTransf(kind, tag, *rest):
    if kind != "Linear":
        raise NotImplementedError("shim supports the Linear transformation only")
    _D.transf[int(tag)] = kind


def element(kind, tag, ni, nj, A, E, I, transf, *rest):
    if kind != "elasticBeamColumn":
        raise NotImplementedError(kind)
    _D.elements[int(tag)] = (int(ni), int(nj), float(A), float(E), float(I), int(transf))


def timeSeries(kind, tag, *rest):
    if kind != "Linear":
        raise NotImplementedError(kind)


def pattern(kind, tag, ts, *rest):
    if kind != "Plain":
        raise NotImplementedError(kind)


def load(node_tag, fx, fy, mz):
    _D.nodal_loads.append((int(node_tag), float(fx), float(fy), float(mz)))


def eleLoad(*args):
    # ops.eleLoad('-ele', e, '-type', '-beamUniform', Wy, Wx)   (SingleCore:117)
    a = list(args)
    i_ele = a.index("-ele")
    i_type = a.index("-type")
    tags = [int(t) for t in a[i_ele + 1:i_type]]
    if a[i_type + 1] != "-beamUniform":
        raise NotImplementedError(a[i_type + 1])
    vals = a[i_type + 2:]
    wt = float(vals[0])
    wa = float(vals[1]) if len(vals) > 1 else 0.0
    for t in tags:
        acc = _D.ele_loads.setdefault(t, [0.0, 0.0])
        acc[0] += wt
        acc[1] += wa


def system(kind, *rest):
    _D.system = kind


def numberer(*a):
    pass


def constraints(kind, *rest):
    if kind != "Plain":
        raise NotImplementedError(kind)


def integrator(kind, dlam=1.0, *rest):
    if kind != "LoadControl":
        raise NotImplementedError(kind)
    _D.integrator_dlambda = float(dlam)


def algorithm(kind, *rest):
    if kind not in ("Linear", "Newton"):
        raise NotImplementedError(kind)


def analysis(kind, *rest):
    if kind != "Static":
        raise NotImplementedError(kind)
    _D.analysis_type = kind


def _element_geometry(ni, nj):
    xi, yi = _D.nodes[ni]
    xj, yj = _D.nodes[nj]
    dx, dy = xj - xi, yj - yi
    L = float(np.hypot(dx, dy))
    return L, dx / L, dy / L


def _local_stiffness(A, E, I, L):
    EAoL = E * A / L
    EIoL = E * I / L
    k = np.zeros((6, 6))
    k[0, 0] = k[3, 3] = EAoL
    k[0, 3] = k[3, 0] = -EAoL
    a = 12.0 * EIoL / (L * L)
    b = 6.0 * EIoL / L
    c = 4.0 * EIoL
    d = 2.0 * EIoL
    k[1, 1] = a;  k[1, 2] = b;  k[1, 4] = -a; k[1, 5] = b
    k[2, 1] = b;  k[2, 2] = c;  k[2, 4] = -b; k[2, 5] = d
    k[4, 1] = -a; k[4, 2] = -b; k[4, 4] = a;  k[4, 5] = -b
    k[5, 1] = b;  k[5, 2] = d;  k[5, 4] = -b; k[5, 5] = c
    return k


def _rotation(c, s):
    T = np.zeros((6, 6))
    for o in (0, 3):
        T[o, o] = c;  T[o, o + 1] = s
        T[o + 1, o] = -s; T[o + 1, o + 1] = c
        T[o + 2, o + 2] = 1.0
    return T


def _fixed_end_local(wt, wa, L):
    """Local resisting force of the element at zero displacement (= -equivalent nodal loads).

    ElasticBeam2d::addLoad (beamUniform): V = wt*L/2, M = V*L/6, P = wa*L;
    p0 = (-P, -V, -V), q0 = (-P/2, -M, +M); LinearCrdTransf2d::getGlobalResistingForce maps
    (q, p0) -> pl = (-q0+p0[0], (q1+q2)/L + p0[1], q1, q0, -(q1+q2)/L + p0[2], q2).
    """
    V = 0.5 * wt * L
    M = V * L / 6.0
    P = wa * L
    q0, q1, q2 = -0.5 * P, -M, M
    p0 = (-P, -V, -V)
    return np.array([-q0 + p0[0], (q1 + q2) / L + p0[1], q1, q0, -(q1 + q2) / L + p0[2], q2])


def _resisting_local(ul, sec, w):
    """ElasticBeam2d::getResistingForce through the basic system (LinearCrdTransf2d::update)."""
    A, E, I, L = sec
    wt, wa = w
    ub0 = ul[3] - ul[0]
    chord = (ul[1] - ul[4]) / L
    ub1 = ul[2] + chord
    ub2 = ul[5] + chord
    EoverL = E / L
    EIoverL2 = 2.0 * I * EoverL
    EIoverL4 = 2.0 * EIoverL2
    V = 0.5 * wt * L
    M = V * L / 6.0
    P = wa * L
    q0 = A * EoverL * ub0 - 0.5 * P
    q1 = EIoverL4 * ub1 + EIoverL2 * ub2 - M
    q2 = EIoverL2 * ub1 + EIoverL4 * ub2 + M
    Vb = (q1 + q2) / L
    return np.array([-q0 - P, Vb - V, q1, q0, -Vb - V, q2])


def analyze(nsteps=1):
    """One LoadControl step per call: lambda += dlambda, linear solve, commit.  Returns 0 on success."""
    tags = sorted(_D.nodes)
    index = {t: i for i, t in enumerate(tags)}
    ndof = 3 * len(tags)
    for _ in range(int(nsteps)):
        _D.load_factor += _D.integrator_dlambda
        lam = _D.load_factor
        K = np.zeros((ndof, ndof))
        f = np.zeros(ndof)
        cache = {}
        hbw = 0
        for et, (ni, nj, A, E, I, _tr) in _D.elements.items():
            L, c, s = _element_geometry(ni, nj)
            kl = _local_stiffness(A, E, I, L)
            T = _rotation(c, s)
            kg = T.T @ kl @ T
            dofs = np.r_[3 * index[ni] + np.arange(3), 3 * index[nj] + np.arange(3)]
            K[np.ix_(dofs, dofs)] += kg
            hbw = max(hbw, int(dofs.max() - dofs.min()))
            fe = np.zeros(6)
            if et in _D.ele_loads:
                wt, wa = _D.ele_loads[et]
                fe = _fixed_end_local(lam * wt, lam * wa, L)
                f[dofs] -= T.T @ fe
            cache[et] = (dofs, (A, E, I, L), T, (lam * _D.ele_loads[et][0], lam * _D.ele_loads[et][1])
                         if et in _D.ele_loads else (0.0, 0.0))
        for (nt, fx, fy, mz) in _D.nodal_loads:
            b = 3 * index[nt]
            f[b] += lam * fx
            f[b + 1] += lam * fy
            f[b + 2] += lam * mz
        free = np.ones(ndof, dtype=bool)
        for nt, flags in _D.fixes.items():
            b = 3 * index[nt]
            for k in range(3):
                if flags[k]:
                    free[b + k] = False
        fidx = np.nonzero(free)[0]
        Kff = K[np.ix_(fidx, fidx)]
        ff = f[fidx]
        n = len(fidx)
        kd = min(hbw, n - 1)
        try:
            if _D.system == "BandGeneral":
                # BandGeneral: LAPACK dgbsv (band LU with partial pivoting), the frame optimiser's system
                # (OpenPyStruct_FrameOpt_Discrete_Beta.py:132)
                gb = np.zeros((2 * kd + 1, n))
                for d in range(-kd, kd + 1):
                    diag = np.diagonal(Kff, d)
                    if d >= 0:
                        gb[kd - d, d:] = diag
                    else:
                        gb[kd - d, :n + d] = diag
                uf = solve_banded((kd, kd), gb, ff, check_finite=True)
            else:
                # BandSPD: upper banded storage for dpbsv.
                ab = np.zeros((kd + 1, n))
                for d in range(kd + 1):
                    ab[kd - d, d:] = np.diagonal(Kff, d)
                uf = solveh_banded(ab, ff, lower=False, check_finite=True)
        except Exception:
            return -3
        u = np.zeros(ndof)
        u[fidx] = uf
        _D.disp = {t: u[3 * index[t]:3 * index[t] + 3].copy() for t in tags}
        _D.ele_forces = {}
        for et, (dofs, sec, T, w) in cache.items():
            _D.ele_forces[et] = T.T @ _resisting_local(T @ u[dofs], sec, w)
        _D._K = K
        _D._f = f
        _D._u = u
    return 0


def eleResponse(tag, *what):
    if not what or what[0] not in ("forces", "force", "globalForce", "globalForces"):
        raise NotImplementedError(what)
    return [float(v) for v in _D.ele_forces[int(tag)]]


def nodeDisp(tag, dof=None):
    d = _D.disp[int(tag)]
    if dof is None:
        return [float(v) for v in d]
    return float(d[int(dof) - 1])


def reactions():
    _D._R = _D._K @ _D._u - _D._f


def nodeReaction(tag, dof=None):
    tags = sorted(_D.nodes)
    i = tags.index(int(tag))
    r = _D._R[3 * i:3 * i + 3]
    if dof is None:
        return [float(v) for v in r]
    return float(r[int(dof) - 1])
