"""Host-side sampling of supports and loads, in the reference's draw order so that one
``random.seed(s)`` reproduces the reference's stream (SingleCore:133-160, MultiCore:137-162,
GPU:141-170; BeamOpt:56-80).  The reference never seeds (`grep seed` finds nothing); seeding is the
only addition.  Node tags are 1-based like the reference's; the ABI arrays are 0-based."""
from __future__ import annotations

import random
from typing import List, Optional, Sequence, Tuple

import numpy as np

Case = Tuple[float, List[int], List[int], List[float]]   # (L, roller tags, force tags, force values)


def fixed_bridge(num_nodes: int = 101, roller_nodes: Optional[Sequence[int]] = None):
    """roller_nodes / available_nodes of the fixed configuration (SingleCore:58-66)."""
    rollers = list(roller_nodes) if roller_nodes is not None else [10, 30, 70, 85, num_nodes - 1]
    available = [t for t in range(2, num_nodes) if t not in rollers]
    return rollers, available


def sample_case(num_nodes: int, flag: int, L: float, roller_nodes: Sequence[int],
                available_nodes: Sequence[int], *, L_max: float = 200.0, L_min: float = 15.0,
                N_rollers_max: int = 4, M_forces_max: int = 4, max_force: float = -355857,
                min_force: Optional[float] = None, rng=random) -> Case:
    """One generate_sample() draw.  Order of RNG calls (flag=1): uniform (L), randint (#rollers),
    choice x #rollers; then always: randint (#forces), sample (force nodes), uniform x #forces."""
    if min_force is None:
        min_force = max_force / 10
    if flag == 1:
        L = L_min + rng.uniform(0, L_max)
        rollers: List[int] = []
        avail = list(range(2, num_nodes))
        num_rollers = rng.randint(1, N_rollers_max)
        for _ in range(num_rollers):
            if avail:
                r = rng.choice(avail)
                rollers.append(r)
                avail.remove(r)
    else:
        rollers = list(roller_nodes)
        avail = list(available_nodes)
    k = min(rng.randint(1, M_forces_max), len(avail))
    force_nodes = rng.sample(avail, k)
    force_values = [rng.uniform(min_force, max_force) for _ in force_nodes]
    return L, rollers, force_nodes, force_values


def sample_beamopt_case(num_nodes: int = 101, L: float = 200.0, N_rollers: int = 5, M_forces: int = 5,
                        L_min: int = 15, max_force: float = -355857, rng=random) -> Case:
    """The single-beam script's draw (BeamOpt:56-80): rollers by rejection on |node distance| >= L_min."""
    avail = list(range(2, num_nodes))
    rollers = [rng.choice(avail)]
    avail.remove(rollers[0])
    for _ in range(1, N_rollers):
        if not any(all(abs(c - r) >= L_min for r in rollers) for c in avail):
            # the reference's rejection loop (BeamOpt:68-76) would spin forever on this draw
            raise RuntimeError("no node satisfies the minimum roller spacing; re-seed")
        while True:
            cand = rng.choice(avail)
            if all(abs(cand - r) >= L_min for r in rollers):
                rollers.append(cand)
                avail.remove(cand)
                break
    pool = [t for t in range(2, num_nodes) if t not in rollers]
    force_nodes = rng.sample(pool, min(M_forces, len(pool)))
    force_values = [rng.uniform(0.5 * max_force, max_force) for _ in force_nodes]
    return L, rollers, force_nodes, force_values


def pack_cases(num_nodes: int, max_forces: int, cases: Sequence[Case], num_cases: int = 1):
    """Cases -> ABI arrays.  With num_cases > 1 consecutive cases share a beam (supports of the first)."""
    assert len(cases) % num_cases == 0
    B = len(cases) // num_cases
    fixed = np.zeros((B, num_nodes), np.uint8)
    fn = np.full((B, num_cases, max_forces), -1, np.int32)
    fv = np.zeros((B, num_cases, max_forces), np.float64)
    L = np.zeros(B, np.float64)
    for i, (Lb, rollers, ftags, fvals) in enumerate(cases):
        b, c = divmod(i, num_cases)
        if len(ftags) > max_forces:
            raise ValueError("more point loads than max_forces")
        if c == 0:
            L[b] = Lb
            fixed[b, 0] = 1
            for t in rollers:
                fixed[b, t - 1] = 1
        for j, (t, F) in enumerate(zip(ftags, fvals)):
            fn[b, c, j] = t - 1
            fv[b, c, j] = F
    return fixed, fn, fv, L


# --------------------------------------------------------------------------------------------------
# native fast path: the same draws from a bit-exact replica of CPython's `random` (csrc/sampler_host.cpp)
# --------------------------------------------------------------------------------------------------
class PackedCases:
    """``count`` sampled cases as arrays: the ABI inputs (``fixed_uy, force_nodes, force_vals, L``) plus what the record
    needs besides (1-based ``roller_tags`` / ``force_tags`` padded with 0, per-case ``case_L``).  ``cases()`` rebuilds the
    list-of-tuples form of ``sample_case`` (for ``make_records``); the columnar writer takes the arrays as they are."""

    def __init__(self, num_nodes, num_cases, max_forces, fixed, fn, fv, L, roller_tags, force_tags, case_L):
        self.num_nodes, self.num_cases, self.max_forces = num_nodes, num_cases, max_forces
        self.fixed_uy, self.force_nodes, self.force_vals, self.L = fixed, fn, fv, L
        self.roller_tags, self.force_tags, self.case_L = roller_tags, force_tags, case_L

    def __len__(self):
        return int(self.case_L.shape[0])

    def abi_arrays(self):
        return self.fixed_uy, self.force_nodes, self.force_vals, self.L

    def cases(self) -> List[Case]:
        out = []
        fv = self.force_vals.reshape(len(self), self.max_forces)
        for i in range(len(self)):
            k = int(np.count_nonzero(self.force_tags[i]))
            out.append((float(self.case_L[i]), [int(t) for t in self.roller_tags[i] if t > 0],
                        [int(t) for t in self.force_tags[i, :k]], [float(v) for v in fv[i, :k]]))
        return out


class NativeSampler:
    """``random.Random(seed)`` in C: the stream of the reference's sampling statements, ~50 ns per beam."""

    def __init__(self, seed: int):
        import ctypes as C
        from . import _cabi
        if not 0 <= int(seed) < 2 ** 64:
            raise ValueError("seed must be in [0, 2^64)")
        self._C, self._lib = C, _cabi.lib()
        self._h = C.c_void_p()
        _cabi.check(self._lib.ops_sampler_create(int(seed), C.byref(self._h)), "ops_sampler_create")

    def random(self) -> float:
        return float(self._lib.ops_sampler_random(self._h))

    def randint(self, a: int, b: int) -> int:
        return int(self._lib.ops_sampler_randint(self._h, a, b))

    def draw_cases(self, count: int, num_nodes: int, flag: int, L: float, roller_nodes: Sequence[int],
                   available_nodes: Sequence[int], *, L_max: float = 200.0, L_min: float = 15.0, N_rollers_max: int = 4,
                   M_forces_max: int = 4, max_force: float = -355857, min_force: Optional[float] = None,
                   num_cases: int = 1, max_forces: int = 4) -> PackedCases:
        from . import _cabi
        if min_force is None:
            min_force = max_force / 10
        assert count % num_cases == 0
        B = count // num_cases
        rw = max(len(roller_nodes), N_rollers_max, 1)
        fixed = np.zeros((B, num_nodes), np.uint8)
        fn = np.full((B, num_cases, max_forces), -1, np.int32)
        fv = np.zeros((B, num_cases, max_forces), np.float64)
        Lb = np.zeros(B, np.float64)
        rt = np.zeros((count, rw), np.int32)
        ft = np.zeros((count, max_forces), np.int32)
        cL = np.zeros(count, np.float64)
        rn = np.ascontiguousarray(roller_nodes, np.int32)
        an = np.ascontiguousarray(available_nodes, np.int32)
        rc = self._lib.ops_sampler_draw_cases(
            self._h, count, num_nodes, flag, float(L), rn.ctypes.data, len(rn), an.ctypes.data, len(an), float(L_max),
            float(L_min), N_rollers_max, M_forces_max, float(max_force), float(min_force), num_cases, max_forces,
            fixed.ctypes.data, fn.ctypes.data, fv.ctypes.data, Lb.ctypes.data, rt.ctypes.data, rw, ft.ctypes.data,
            cL.ctypes.data)
        _cabi.check(rc, "ops_sampler_draw_cases")
        return PackedCases(num_nodes, num_cases, max_forces, fixed, fn, fv, Lb, rt, ft, cL)

    def close(self):
        if self._h:
            self._lib.ops_sampler_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:      # noqa: BLE001
            pass
