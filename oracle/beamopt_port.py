"""TEST INFRASTRUCTURE ONLY -- "torch path" port of the reference's per-sample optimisation loop.

A restatement (not a copy) of ``generate_sample`` of the reference generators that can travel to
the GPU box, where ``/root/reference`` and OpenSeesPy do not exist:

* FP64 half (OpenSees in the reference, SingleCore:89-124,176-190,221-232): Euler-Bernoulli
  bending system on (uy, theta) DOFs, constrained equations dropped (``Plain``), banded Cholesky
  ``dpbsv`` through ``scipy.linalg.solveh_banded`` (``BandSPD``), element end forces through the
  basic system (``ElasticBeam2d`` / ``LinearCrdTransf2d``).  The axial system that the reference also
  solves (``eleLoad ... w, w`` puts ``w`` on Wx too, SingleCore:117) is decoupled from every emitted
  field for a straight horizontal beam and is dropped.
* FP32 half (SingleCore:163-219): the REAL torch ops -- ``torch.sum``, autograd, ``torch.optim.Adam``
  (single-tensor CPU path), ``ExponentialLR``, ``clamp_`` -- in the reference's order.

``bench.py`` times this port as the CPU baseline (``kind: "port"``); the tests use it as the checker
for seeds beyond the committed goldens.  It is validated here against the reference's own source
executed on the shim (tests/test_oracle_port.py, tests/golden/).

PARITY UNPINNED (see oracle/opensees_shim.py).
"""
from __future__ import annotations

import dataclasses
import random
from typing import List, Optional, Sequence

import numpy as np
from scipy.linalg import solveh_banded


@dataclasses.dataclass
class BeamOptParams:
    """Module-level constants of the generators (SingleCore:20-49 / MultiCore:20-52 / GPU:21-56 / BeamOpt:24-48)."""
    E: float = 200e9
    nu: float = 0.3
    num_nodes: int = 101
    uniform_udl: float = -1000.0
    I_0: float = 0.5
    max_e: int = 600
    lr: float = 0.01
    gamma: float = 0.98
    alpha_moment: float = 1e-2
    alpha_shear: float = 1e-2
    tolerance: float = 5e-3
    patience: int = 5
    shear_k: float = 0.03            # A_approx = 0.03 * I**0.5   (SingleCore:196)
    bending_eps: float = 1e-6        # 2*E*I + 1e-6               (SingleCore:195)
    clamp_min: float = 1e-8          # I_tensor.clamp_(min=1e-8)   (SingleCore:208)
    early_stop: bool = True          # False -> fixed epoch count (benchmark mode)
    zero_last_node: bool = False     # MultiCore:222-223 emits 0.0 for the last node

    @property
    def G(self) -> float:
        return self.E / (2 * (1 + self.nu))

    @property
    def num_elements(self) -> int:
        return self.num_nodes - 1

    @staticmethod
    def for_script(which: str) -> "BeamOptParams":
        if which == "SC":
            return BeamOptParams(tolerance=5e-3, patience=5)
        if which == "MC":   # def-default patience=10 shadows the module constant (MultiCore:130 vs :44)
            return BeamOptParams(tolerance=5e-3, patience=10, zero_last_node=True)
        if which == "GPU":
            return BeamOptParams(tolerance=1e-2, patience=100)
        if which == "BO":
            return BeamOptParams(tolerance=1e-2, patience=10, max_e=1000, uniform_udl=-5000.0)
        raise ValueError(which)


# ----------------------------------------------------------------------------------------------
# FP64 half: one linear solve + force recovery
# ----------------------------------------------------------------------------------------------

def fe_solve(I: np.ndarray, L: float, fixed_uy: np.ndarray, f_uy: np.ndarray, w: float, E: float):
    """Solve the beam for element inertias ``I`` (float64 view of the fp32 parameters).

    fixed_uy[nn] bool (node 0 is always pinned), f_uy[nn] nodal vertical loads.
    Returns (uy[nn], theta[nn], V[n], M[n]) -- V, M = eleResponse(e,'forces')[1], [2].
    """
    n = I.shape[0]
    nn = n + 1
    Le = L / n
    k = E * I / Le ** 3
    N = 2 * nn
    # upper banded storage, half bandwidth 3
    ab = np.zeros((4, N))
    a12 = 12.0 * k
    a6 = 6.0 * Le * k
    a4 = 4.0 * Le * Le * k
    a2 = 2.0 * Le * Le * k
    iu = 2 * np.arange(n)            # uy dof of node e
    # diagonal
    np.add.at(ab[3], iu, a12)
    np.add.at(ab[3], iu + 1, a4)
    np.add.at(ab[3], iu + 2, a12)
    np.add.at(ab[3], iu + 3, a4)
    # first super-diagonal: (uy_e,th_e)=6Le k ; (th_e,uy_e+1)=-6Le k ; (uy_e+1,th_e+1)=-6Le k
    np.add.at(ab[2], iu + 1, a6)
    np.add.at(ab[2], iu + 2, -a6)
    np.add.at(ab[2], iu + 3, -a6)
    # second: (uy_e,uy_e+1)=-12k ; (th_e,th_e+1)=2Le^2 k
    np.add.at(ab[1], iu + 2, -a12)
    np.add.at(ab[1], iu + 3, a2)
    # third: (uy_e,th_e+1)=6Le k
    np.add.at(ab[0], iu + 3, a6)
    f = np.zeros(N)
    f[0::2] = f_uy
    f[0::2][:-1] += 0.5 * w * Le
    f[0::2][1:] += 0.5 * w * Le
    f[1] += w * Le * Le / 12.0
    f[N - 1] -= w * Le * Le / 12.0
    fixed = np.zeros(N, dtype=bool)
    fixed[0::2] = fixed_uy
    fixed[0] = True
    free = ~fixed
    # Plain handler: drop constrained equations.  Build the reduced band by dense gather of the band.
    idx = np.nonzero(free)[0]
    m = idx.size
    abr = np.zeros((4, m))
    pos = -np.ones(N, dtype=np.int64)
    pos[idx] = np.arange(m)
    for d in range(4):
        cols = np.arange(d, N)
        rows = cols - d
        keep = free[cols] & free[rows]
        c = pos[cols[keep]]
        r = pos[rows[keep]]
        dd = c - r                      # reduced distance (<= d)
        abr[3 - dd, c] = ab[3 - d, cols[keep]]
    u_f = solveh_banded(abr, f[idx], lower=False)
    u = np.zeros(N)
    u[idx] = u_f
    uy = u[0::2]
    th = u[1::2]
    chord = (uy[:-1] - uy[1:]) / Le
    v1 = th[:-1] + chord
    v2 = th[1:] + chord
    EoverL = E / Le
    c2 = 2.0 * I * EoverL
    c4 = 2.0 * c2
    Vfe = 0.5 * w * Le
    Mfe = Vfe * Le / 6.0
    q1 = c4 * v1 + c2 * v2 - Mfe
    q2 = c2 * v1 + c4 * v2 + Mfe
    V = (q1 + q2) / Le - Vfe
    return uy, th, V, q1


# ----------------------------------------------------------------------------------------------
# sampling (host side, kept in the reference's call order so seeds line up)
# ----------------------------------------------------------------------------------------------

DEFAULT_ROLLERS = (10, 30, 70, 85, 100)   # 1-based node tags, [10,30,70,85,num_nodes-1] (SingleCore:62)


def sample_case(p: BeamOptParams, flag: int = 0, *, L_max: float = 200.0, L_min: float = 15.0,
                N_rollers_max: int = 4, M_forces_max: int = 4, max_force: float = -355857,
                roller_nodes: Optional[Sequence[int]] = None, rng=random):
    """One draw in the order of SingleCore:133-160: [L, rollers if flag] randint, sample, uniform*k."""
    num_nodes = p.num_nodes
    min_force = max_force / 10
    if flag == 1:
        L = L_min + rng.uniform(0, L_max)
        rollers, avail = [], list(range(2, num_nodes))
        num_rollers = rng.randint(1, N_rollers_max)
        first = rng.choice(avail)
        rollers.append(first)
        avail.remove(first)
        for _ in range(num_rollers - 1):
            if avail:
                r = rng.choice(avail)
                rollers.append(r)
                avail.remove(r)
    else:
        L = L_max
        rollers = list(roller_nodes if roller_nodes is not None
                       else (10, 30, 70, 85, num_nodes - 1))
        avail = [t for t in range(2, num_nodes) if t not in rollers]
    k = rng.randint(1, M_forces_max)
    k = min(k, len(avail))
    force_nodes = rng.sample(avail, k)
    force_values = [rng.uniform(min_force, max_force) for _ in force_nodes]
    return L, rollers, force_nodes, force_values


# ----------------------------------------------------------------------------------------------
# the loop
# ----------------------------------------------------------------------------------------------

def optimise_beam(p: BeamOptParams, L: float, roller_nodes: Sequence[int], force_nodes: Sequence[int],
                  force_values: Sequence[float], *, trace: Optional[list] = None):
    """Port of the epoch loop SingleCore:163-232.  Node tags are 1-based like the reference."""
    import torch
    from torch.optim.lr_scheduler import ExponentialLR

    n = p.num_elements
    nn = p.num_nodes
    fixed = np.zeros(nn, dtype=bool)
    fixed[0] = True
    for t in roller_nodes:
        fixed[t - 1] = True
    f_uy = np.zeros(nn)
    for t, F in zip(force_nodes, force_values):
        f_uy[t - 1] += F
    E, G = p.E, p.G

    I_tensor = torch.tensor([p.I_0] * n, dtype=torch.float32, requires_grad=True)
    optimizer = torch.optim.Adam([I_tensor], lr=p.lr)
    scheduler = ExponentialLR(optimizer, gamma=p.gamma)
    best_loss = float("inf")
    counter = 0
    epochs = 0
    status = 0
    uy = th = None
    Vt = Mt = None
    loss_val = float("nan")
    for _epoch in range(p.max_e):
        optimizer.zero_grad()
        I64 = I_tensor.detach().numpy().astype(np.float64)
        try:
            uy, th, V, M = fe_solve(I64, L, fixed, f_uy, p.uniform_udl, E)
        except Exception:
            status = 1
            break
        if not (np.all(np.isfinite(uy)) and np.all(np.isfinite(th))):
            status = 1
            break
        Mt = torch.tensor(M.tolist(), dtype=torch.float32)
        Vt = torch.tensor(V.tolist(), dtype=torch.float32)
        bending_energy = torch.sum((Mt ** 2) / (2 * E * I_tensor + p.bending_eps))
        A_approx = p.shear_k * I_tensor ** 0.5
        shear_energy = torch.sum(Vt ** 2 / (G * A_approx))
        primary = torch.sum(I_tensor)
        total = primary + p.alpha_moment * bending_energy + p.alpha_shear * shear_energy
        if trace is not None:
            trace.append({"I": I_tensor.detach().numpy().copy(), "M": M.copy(), "V": V.copy(),
                          "loss": total.item()})
        total.backward()
        optimizer.step()
        scheduler.step()
        with torch.no_grad():
            I_tensor.clamp_(min=p.clamp_min)
        epochs += 1
        loss_val = total.item()
        if p.early_stop:
            if loss_val < best_loss - p.tolerance:
                best_loss = loss_val
                counter = 0
            else:
                counter += 1
            if counter >= p.patience:
                break
    if uy is None:
        return None
    defl = uy.copy()
    rot = th.copy()
    if p.zero_last_node:
        defl[-1] = 0.0
        rot[-1] = 0.0
    return {
        "I_values": I_tensor.detach().numpy().copy(),
        "shear_forces": Vt.numpy().copy(),
        "bending_moments": Mt.numpy().copy(),
        "rotations": rot,
        "deflections": defl,
        "epochs": epochs,
        "loss": np.float32(loss_val),
        "status": status,
    }


def generate_sample(p: BeamOptParams, flag: int = 0, rng=random, **sample_kw) -> Optional[dict]:
    """Sampling + loop + the 13-key record of SingleCore:235-249."""
    L, rollers, force_nodes, force_values = sample_case(p, flag, rng=rng, **sample_kw)
    out = optimise_beam(p, L, rollers, force_nodes, force_values)
    if out is None or out["status"] != 0:
        return None
    node_positions = np.linspace(0, L, p.num_nodes)
    return {
        "roller_x_locations": [node_positions[t - 1] for t in rollers],
        "force_x_locations": [node_positions[t - 1] for t in force_nodes],
        "force_values": force_values,
        "I_values": out["I_values"].tolist(),
        "shear_forces": out["shear_forces"].tolist(),
        "bending_moments": out["bending_moments"].tolist(),
        "node_positions": node_positions.tolist(),
        "roller_nodes": rollers,
        "force_nodes": force_nodes,
        "num_nodes": p.num_nodes,
        "L": L,
        "rotations": out["rotations"].tolist(),
        "deflections": out["deflections"].tolist(),
        "_epochs": out["epochs"],
    }


def _worker(args):
    import torch
    torch.set_num_threads(1)
    p, seed, count, flag = args
    rng = random.Random(seed)
    done = 0
    for _ in range(count):
        if generate_sample(p, flag, rng=rng) is not None:
            done += 1
    return done


class PortPool:
    """Process pool over beams (the MultiCore:258 joblib/loky pattern) kept alive across timed steps."""

    def __init__(self, p: BeamOptParams, workers: int, flag: int = 0):
        import multiprocessing as mp
        self.p, self.workers, self.flag = p, workers, flag
        self.pool = mp.get_context("spawn").Pool(workers)
        self.pool.map(_worker, [(p, 0, 0, flag)] * workers)          # import torch in every worker
        self.calls = 0

    def run(self, beams: int, seed: int = 0):
        """Optimise `beams` freshly sampled beams; returns (beams_done, seconds)."""
        import time
        w = self.workers
        per = [beams // w + (1 if i < beams % w else 0) for i in range(w)]
        self.calls += 1
        jobs = [(self.p, seed + 7919 * i + 104729 * self.calls, c, self.flag) for i, c in enumerate(per) if c > 0]
        t0 = time.perf_counter()
        done = sum(self.pool.map(_worker, jobs, chunksize=1))
        return done, time.perf_counter() - t0

    def close(self):
        self.pool.close()
        self.pool.join()


def timed_pool_run(p: BeamOptParams, beams: int, workers: int, seed: int = 0, flag: int = 0):
    """One-shot PortPool run.  Returns (beams_done, seconds)."""
    pool = PortPool(p, workers, flag)
    try:
        return pool.run(beams, seed)
    finally:
        pool.close()
