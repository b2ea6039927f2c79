"""Parity of the CUDA path, called through the C ABI, against the CPU oracle and the committed
reference fixtures.  Runs on the B200 box (`-m gpu`); reads nothing outside the repo.

Tolerances (BASELINE.json north_star): displacements and forces 1e-9 relative per FP64 solve;
optimised I 1e-5 relative after a fixed epoch count; identical early-stop decisions."""
import numpy as np
import pytest
import torch

from oracle import c_oracle
from openpystruct_b200 import _cabi, generator, ops, sampling
from openpystruct_b200.params import BeamOptParams
from tests.helpers import (goldens, golden_decision_margin, golden_params, golden_case, oracle_params, oracle_run,
                           oracle_run_mt, rel_err, seeded_cases)

pytestmark = pytest.mark.gpu

# the device solvers behind the same ABI: 0 = three-moment, 8 lanes per beam (production default),
# 1 = banded LDL^T, 2 = three-moment, thread per beam, 3 / 4 = three-moment with the optimiser state in
# shared memory, 8 / 32 lanes per beam (4 is what solver 0 runs beyond 169 nodes)
SOLVERS = [pytest.param(0, id="three_moment_lanes"), pytest.param(1, id="band_ldlt"),
           pytest.param(2, id="three_moment_thread"), pytest.param(3, id="three_moment_smem8"),
           pytest.param(4, id="three_moment_smem32")]


def gpu_run(p, fixed, fn, fv, L):
    """Through the torch custom op (device tensors) -> ops_beamopt_launch."""
    dev = torch.device("cuda", 0)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)   # noqa: E731
    out = ops.optimise_beams(p, t(fixed), t(fn), t(fv), t(L))
    torch.cuda.synchronize()
    return {k: v.cpu().numpy() for k, v in out.items()}


# "Identical early-stop decisions" (north_star), made testable: the stop test compares the fp32 loss with
# best - tolerance (SingleCore:211); the oracle reports, per beam, how close the closest such comparison of the run
# came to going the other way, in fp32 ulps of the loss (margin_ulps).  The loss of two correct implementations can
# differ in the last bit (M, V differ at 1e-11 before the fp32 cast), so a stop epoch may differ ONLY on beams whose
# margin is about one ulp -- everywhere else (>= 85 % of all beams) the decisions must be identical, and the beams
# that do differ are bounded in number as well (measured on 40 000 beams: 0.01-0.03 % on the fixed bridge).
MARGIN_ULPS = 2.0


def assert_same_decisions(o, g, max_flip_rate):
    """o = oracle, g = CUDA path; returns the mask of beams with identical stop epochs."""
    assert np.array_equal(o["status"], g["status"])
    flips = o["epochs"] != g["epochs"]
    assert (o["margin_ulps"][flips] <= MARGIN_ULPS).all(), (np.flatnonzero(flips), o["margin_ulps"][flips],
                                                            o["epochs"][flips], g["epochs"][flips])
    assert flips.mean() <= max_flip_rate, (int(flips.sum()), flips.size)
    return ~flips


def assert_matches_oracle(a, b, flag=0, truth=None):
    """a = FP64 oracle (the reference's banded-Cholesky arithmetic), b = CUDA path, truth = the same loop with the FE
    solve in 80-bit arithmetic (needed on random bridges, whose ill-conditioned K makes FP64 Cholesky itself noisy)."""
    same = assert_same_decisions(a, b, 0.002 if flag == 0 else 0.02)
    assert np.max(np.abs(a["I"][same] - b["I"][same]) / a["I"][same]) < 1e-5
    assert (a["I"][same] == b["I"][same]).mean() > 0.98
    assert rel_err(b["defl"][same, 0], a["defl"][same, 0]).max() < (1e-7 if flag == 0 else 1e-5)
    assert rel_err(b["rot"][same, 0], a["rot"][same, 0]).max() < (1e-7 if flag == 0 else 1e-5)
    assert rel_err(b["moment"][same, 0], a["moment"][same, 0]).max() < 1e-5
    assert rel_err(b["shear"][same, 0], a["shear"][same, 0]).max() < 1e-5
    assert (a["loss"][same] == b["loss"][same]).mean() > 0.98
    if truth is not None:
        st = assert_same_decisions(truth, b, 0.002)
        assert np.max(np.abs(truth["I"][st] - b["I"][st]) / truth["I"][st]) < 1e-5
        assert (truth["I"][st] == b["I"][st]).all(axis=1).mean() > 0.99
        assert (truth["loss"][st] == b["loss"][st]).mean() > 0.99


def test_library_sees_the_gpu():
    assert _cabi.lib().ops_device_count() >= 1
    assert "sm_100a" in _cabi.version()


def test_branch_free_div_sqrt_are_ieee():
    """csrc/fastmath.cuh: the production kernel's fp32 division / square-root sequences give the bits
    of the IEEE operators (which is what torch's CPU kernels compute) on 2^27 random operands."""
    r = _cabi.fastmath_selftest(1 << 27)
    assert r["samples"] >= 1 << 27
    assert (r["div"], r["sqrt"], r["rcp"]) == (0, 0, 0), r
    assert r["rcp64_max_rel_err"] < 6.7e-16, r            # <= 3 ulp of 1/x


@pytest.mark.parametrize("solver", SOLVERS)
@pytest.mark.parametrize("script,flag,count", [("SC", 0, 512), ("MC", 0, 512), ("GPU", 0, 64), ("SC", 1, 512)])
def test_full_loop_against_c_oracle(script, flag, count, solver):
    p = BeamOptParams.for_script(script).replace(solver=solver)
    cases = seeded_cases(p, count, seed=101, flag=flag)
    fixed, fn, fv, L = sampling.pack_cases(p.num_nodes, p.max_forces, cases)
    truth = oracle_run(p, fixed, fn, fv, L, 1) if (flag == 1 and solver != 1) else None
    assert_matches_oracle(oracle_run(p, fixed, fn, fv, L), gpu_run(p, fixed, fn, fv, L), flag, truth)


@pytest.mark.parametrize("solver", SOLVERS)
@pytest.mark.parametrize("flag", [0, 1])
def test_fixed_600_epochs_I_within_1e5(solver, flag):
    """Optimised I within 1e-5 relative after a fixed epoch count, on ALL beams (no stop decision involved).
    flag=1 (random bridges: short stiff spans, cond(K) up to 1e10): the tolerance is asserted against the loop with
    the 80-bit FE solve; the reference's own FP64 banded Cholesky is further from that exact solve than the CUDA
    path is, so against the FP64 oracle the bound is the FP64 oracle's own distance to the truth."""
    p = BeamOptParams.for_script("MC").replace(early_stop=False, solver=solver)
    cases = seeded_cases(p, 256, seed=102, flag=flag)
    fixed, fn, fv, L = sampling.pack_cases(p.num_nodes, p.max_forces, cases)
    a, b = oracle_run(p, fixed, fn, fv, L), gpu_run(p, fixed, fn, fv, L)
    assert (b["epochs"] == 600).all() and not b["status"].any()
    err64 = np.max(np.abs(a["I"] - b["I"]) / a["I"])
    if flag == 0:
        assert err64 < 1e-5
    if flag == 1 or solver == 0:
        t = oracle_run(p, fixed, fn, fv, L, 1)
        err80 = np.max(np.abs(t["I"] - b["I"]) / t["I"])
        own = np.max(np.abs(t["I"] - a["I"]) / t["I"])            # FP64 oracle against the 80-bit loop
        if solver != 1:                                           # (solver 1 IS an FP64 band factorisation)
            assert err80 < 1e-5, (err80, own)
        assert err64 <= 2 * own + 1e-5 and err80 <= 2 * own + 1e-5, (err64, err80, own)
    assert (b["defl"][:, 0, -1] == 0).all() and (b["rot"][:, 0, -1] == 0).all()    # MultiCore:222-223


# Reference runs whose stop epoch the CUDA path does not reproduce: torch's CPU sqrt (MKL VML) is not correctly
# rounded on 0.74 % of its inputs (SURVEY finding 5), which moves the fp32 loss by one ulp now and then; on these two
# runs (same beam: SC and MC scripts draw the same seed-0 stream) the reference's own closest decision was 0.7 ulp of
# the loss away from going the other way (epoch 216) -- the IEEE-sqrt oracles (FP64 and 80-bit FE) both stop where
# the CUDA path stops (232 / 268).  Every other run must stop on the reference's epoch.
GOLDEN_FLIPS = {("SC", 0, 0): 232, ("MC", 0, 0): 268}


@pytest.mark.parametrize("solver", SOLVERS)
def test_reference_goldens_through_run_host(solver):
    """The committed reference runs (reference source + torch, made by tests/golden/make_golden.py)
    through the host-buffer C-ABI entry ops_beamopt_run_host."""
    flipped = {}
    for m, rec in goldens():
        p = golden_params(m).replace(solver=solver)
        fixed, fn, fv, L = sampling.pack_cases(p.num_nodes, p.max_forces, [golden_case(m)])
        out = _cabi.run_host(p, fixed, fn, fv, L, device=0)
        assert out["status"][0] == 0
        key = (m["script"], m["seed"], m["flag"])
        if out["epochs"][0] != m["epochs"]:
            flipped[key] = int(out["epochs"][0])
            margin, at = golden_decision_margin(m, rec)
            assert margin < 1.0, (key, margin, at)            # the reference's own decision hung on the last bit
            continue
        # I after a fixed number of epochs never depends on the stop decision: checked below; here the
        # full early-stopped run
        assert np.max(np.abs(out["I"][0] - rec["I_values"]) / rec["I_values"]) < 1e-5
        assert rel_err(out["moment"][0, 0], rec["bending_moments"]) < 1e-6
        assert rel_err(out["shear"][0, 0], rec["shear_forces"]) < 1e-6
        assert rel_err(out["defl"][0, 0], rec["deflections"]) < 1e-6
        assert rel_err(out["rot"][0, 0], rec["rotations"]) < 1e-6
    assert flipped == GOLDEN_FLIPS, flipped


def test_reference_goldens_fixed_epoch_trajectory():
    for m, rec in goldens():
        for k, I_ref in zip(m["trace_epochs"], rec["I_trace"]):
            if k == 0:
                continue
            p = golden_params(m).replace(max_e=k, early_stop=False)
            fixed, fn, fv, L = sampling.pack_cases(p.num_nodes, p.max_forces, [golden_case(m)])
            out = _cabi.run_host(p, fixed, fn, fv, L, device=0)
            assert out["epochs"][0] == k
            assert np.max(np.abs(out["I"][0] - I_ref) / I_ref) < 1e-5, (m["script"], m["seed"], k)


def gpu_solve(p, fixed, fn, fv, L, I):
    dev = torch.device("cuda", 0)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)   # noqa: E731
    out = ops.solve_beams(p, t(fixed), t(fn), t(fv), t(L), t(I))
    torch.cuda.synchronize()
    return {k: v.cpu().numpy() for k, v in out.items()}


@pytest.mark.parametrize("solver", SOLVERS)
def test_single_solve_1e9(solver):
    p = BeamOptParams(solver=solver)
    rng = np.random.default_rng(0)
    cases = seeded_cases(p, 2000, seed=103)
    fixed, fn, fv, L = sampling.pack_cases(p.num_nodes, p.max_forces, cases)
    I = np.exp(rng.uniform(np.log(3e-3), np.log(0.9), (2000, 100))).astype(np.float32).astype(np.float64)
    cp = oracle_params(p)
    o64 = c_oracle.beam_solve(cp, fixed, fn[:, 0], fv[:, 0], L, I, 0)
    o80 = c_oracle.beam_solve(cp, fixed, fn[:, 0], fv[:, 0], L, I, 1)
    g = gpu_solve(p, fixed, fn[:, 0], fv[:, 0], L, I)
    assert not g["status"].any()
    for k in ("defl", "rot", "shear", "moment"):
        assert rel_err(g[k], o64[k]).max() < 1e-9, k
        assert rel_err(g[k], o80[k]).max() < (5e-10 if solver == 1 else 1e-11), k


@pytest.mark.parametrize("solver", SOLVERS)
def test_single_solve_on_reference_goldens(solver):
    for m, rec in goldens():
        p = golden_params(m).replace(solver=solver)
        fixed, fn, fv, L = sampling.pack_cases(p.num_nodes, p.max_forces, [golden_case(m)])
        g = gpu_solve(p, fixed, fn[:, 0], fv[:, 0], L, rec["I_last"][None, :])
        tol = 1e-9 if m["flag"] == 0 else 1e-6
        assert rel_err(g["moment"][0], rec["M64_last"]) < tol
        assert rel_err(g["shear"][0], rec["V64_last"]) < tol
        if not m["zero_last_node"]:
            assert rel_err(g["defl"][0], rec["deflections"]) < tol
            assert rel_err(g["rot"][0], rec["rotations"]) < tol


@pytest.mark.parametrize("solver", SOLVERS)
def test_edge_cases_and_mechanism(solver):
    p = BeamOptParams.for_script("SC").replace(max_e=40, solver=solver)
    cases = [
        (200.0, [10, 30, 70, 85, 100], [], []),
        (200.0, [101], [51], [-1e5]),
        (15.0, [2], [3, 4, 5, 6], [-3.5e5] * 4),
        (215.0, [100], [2, 50, 99, 60], [-3e5, -2e5, -1e5, -5e4]),
        (200.0, [10, 30, 70, 85, 100], [50, 50], [-1e5, -1e5]),
        (200.0, [10, 11, 12, 100, 101], [5, 11, 60], [-1e5, -2e5, -3e5]),   # adjacent rollers, load on a roller
        (200.0, [50], [51, 100, 101], [-1e5, -1e5, -5e4]),                  # long overhang with tip load
        (200.0, [], [50], [-1e5]),                                   # pin only: singular -> status 1
    ]
    fixed, fn, fv, L = sampling.pack_cases(p.num_nodes, p.max_forces, cases)
    a, b = oracle_run(p, fixed, fn, fv, L), gpu_run(p, fixed, fn, fv, L)
    assert a["status"].tolist() == [0] * 7 + [1] == b["status"].tolist()
    ok = a["status"] == 0
    assert np.array_equal(a["epochs"][ok], b["epochs"][ok])
    assert np.max(np.abs(a["I"][ok] - b["I"][ok]) / a["I"][ok]) < 1e-5


@pytest.mark.parametrize("solver", SOLVERS)
def test_empty_and_ragged_batches(solver):
    p = BeamOptParams.for_script("SC").replace(max_e=25, solver=solver)
    dev = torch.device("cuda", 0)
    z = ops.optimise_beams(p, torch.zeros((0, 101), dtype=torch.uint8, device=dev),
                           torch.zeros((0, 1, 4), dtype=torch.int32, device=dev),
                           torch.zeros((0, 1, 4), dtype=torch.float64, device=dev),
                           torch.zeros((0,), dtype=torch.float64, device=dev))
    assert z["I"].shape == (0, 100) and z["epochs"].shape == (0,)
    cases = seeded_cases(p, 333, seed=104)
    fixed, fn, fv, L = sampling.pack_cases(p.num_nodes, p.max_forces, cases)
    whole = gpu_run(p, fixed, fn, fv, L)
    for lo, hi in [(0, 1), (1, 34), (34, 333)]:           # any split of the batch gives the same bytes
        part = gpu_run(p, fixed[lo:hi], fn[lo:hi], fv[lo:hi], L[lo:hi])
        for k in whole:
            assert np.array_equal(part[k], whole[k][lo:hi]), k


def test_beamopt_script_config_five_loads():
    """OpenPyStruct_BeamOpt.py's constants (5 rollers by rejection sampling, 5 loads, UDL -5000, tolerance 1e-2,
    patience 10, num_epochs = 1000, BeamOpt:24-48): the early-stopped loop and all 1000 epochs with the stop off."""
    p = BeamOptParams.for_script("BO")
    assert p.max_e == 1000
    import random
    rng = random.Random(5)
    cases = []
    while len(cases) < 48:
        try:
            cases.append(sampling.sample_beamopt_case(rng=rng))
        except RuntimeError:
            pass
    fixed, fn, fv, L = sampling.pack_cases(p.num_nodes, p.max_forces, cases)
    assert_matches_oracle(oracle_run(p, fixed, fn, fv, L), gpu_run(p, fixed, fn, fv, L))
    pf = p.replace(early_stop=False)
    a, b = oracle_run_mt(pf, fixed, fn, fv, L), gpu_run(pf, fixed, fn, fv, L)
    assert (b["epochs"] == 1000).all() and not b["status"].any()
    assert np.max(np.abs(a["I"] - b["I"]) / a["I"]) < 1e-5


@pytest.mark.parametrize("solver", SOLVERS)
def test_full_size_properties_10k_beams(solver):
    """BASELINE config 2 size (10 000 beams, default discretisation) through size-independent
    properties: determinism, supports stay at zero, nodal equilibrium of the emitted forces,
    I > 0 after the clamp, and a 256-beam slice against the oracle."""
    p = BeamOptParams.for_script("MC").replace(solver=solver)
    cases = seeded_cases(p, 10000, seed=105)
    fixed, fn, fv, L = sampling.pack_cases(p.num_nodes, p.max_forces, cases)
    a = gpu_run(p, fixed, fn, fv, L)
    b = gpu_run(p, fixed, fn, fv, L)
    for k in a:
        assert np.array_equal(a[k], b[k]), k
    assert not a["status"].any()
    assert (a["I"] >= np.float32(1e-8)).all() and np.isfinite(a["I"]).all()
    assert 200 < a["epochs"].mean() < 350 and a["epochs"].max() <= 600
    supports = [0, 9, 29, 69, 84, 99]
    assert (a["defl"][:, 0, supports] == 0).all()
    # equilibrium from the emitted fp32 end forces: R_i = V_i - V_{i-1} - w*Le - P_i vanishes at free nodes
    Le, w = 2.0, p.uniform_udl
    V = a["shear"][:, 0].astype(np.float64)
    P = np.zeros((10000, 101))
    for b_, (_, _, ft, fvv) in enumerate(cases):
        for t_, F in zip(ft, fvv):
            P[b_, t_ - 1] += F
    R = np.zeros((10000, 101))
    R[:, :-1] += V
    R[:, 1:] += -V - w * Le
    R -= P
    free = np.setdiff1d(np.arange(101), supports)
    scale = np.abs(P.sum(1) + w * 200.0)
    assert (np.abs(R[:, free]).max(1) / scale).max() < 1e-6          # fp32 storage of V
    assert np.allclose(R[:, supports].sum(1), -(P.sum(1) + w * 200.0), rtol=1e-6)
    sl = slice(4000, 4256)
    o = oracle_run(p, fixed[sl], fn[sl], fv[sl], L[sl])
    assert_matches_oracle(o, {k: v[sl] for k, v in a.items()})


def test_many_round_batches_use_the_larger_cta_and_agree_bitwise(monkeypatch):
    """Batches of three or more full rounds run the 384-thread instance of the lanes kernel (48 beams
    per SM and round); a beam's arithmetic does not depend on the CTA it runs in, so the record equals
    the 320-thread instance's bit for bit.  A slice is checked against the oracle as well."""
    p = BeamOptParams.for_script("MC")
    B = 148 * 48 * 3 + 77
    cases = seeded_cases(p, B, seed=311)
    fixed, fn, fv, L = sampling.pack_cases(p.num_nodes, p.max_forces, cases)
    monkeypatch.delenv("OPS_LANES_THREADS", raising=False)
    a = gpu_run(p, fixed, fn, fv, L)
    monkeypatch.setenv("OPS_LANES_THREADS", "320")
    b = gpu_run(p, fixed, fn, fv, L)
    monkeypatch.delenv("OPS_LANES_THREADS", raising=False)
    for k in a:
        assert np.array_equal(a[k], b[k]), k
    assert not a["status"].any()
    sl = slice(B - 200, B)
    o = oracle_run(p, fixed[sl], fn[sl], fv[sl], L[sl])
    assert_matches_oracle(o, {k: v[sl] for k, v in a.items()})


@pytest.mark.parametrize("script,flag,early_stop,B", [("MC", 0, False, 148 * 64), ("MC", 0, True, 3000), ("SC", 1, True, 2000)])
def test_tensor_memory_instance_agrees_bitwise(monkeypatch, script, flag, early_stop, B):
    """beamopt_lanes_tm.cu keeps {M0, Q0}, m, v of a lane in tensor memory (tcgen05.ld / tcgen05.st) and holds 64 beams
    per SM; the planner picks it for single rounds of 52..64 beams per SM with a fixed epoch count.  Same phase
    functions, so the record equals the register / shared-memory instance's bit for bit: a full single round, ragged
    early stopping (fresh beams committed next to running ones), random bridges with rejected beams in between."""
    p = BeamOptParams.for_script(script).replace(early_stop=early_stop)
    if not early_stop:
        p = p.replace(max_e=200)
    cases = seeded_cases(p, B, seed=77, flag=flag)
    cases[5] = (200.0, [], [50], [-1e5])                                  # mechanism
    cases[41] = (200.0, [10, 20, 30, 40, 50, 60], [55], [-1e5])          # more rollers than the three-moment kernels take
    fixed, fn, fv, L = sampling.pack_cases(p.num_nodes, p.max_forces, cases)
    monkeypatch.setenv("OPS_LANES_TM", "0")
    a = gpu_run(p, fixed, fn, fv, L)
    monkeypatch.setenv("OPS_LANES_TM", "1")
    b = gpu_run(p, fixed, fn, fv, L)
    monkeypatch.delenv("OPS_LANES_TM", raising=False)
    c = gpu_run(p, fixed, fn, fv, L)                                      # the planner's own choice
    assert a["status"][5] == 1 and a["status"][41] == 3 and (a["status"] == 0).sum() == B - 2
    for k in a:
        assert np.array_equal(a[k], b[k], equal_nan=True), k
        assert np.array_equal(a[k], c[k], equal_nan=True), k


@pytest.mark.parametrize("mode,num_cases,B", [("in_kernel", 1, 700), ("in_kernel", 4, 160), ("pipelined", 1, 148 * 40 * 2 + 333),
                                              ("pipelined", 1, 901), ("pipelined", 4, 211)])
def test_dataset_gather_entry_on_one_gpu(monkeypatch, mode, num_cases, B):
    """ops_beamopt_launch_scatter with THREE destination sets that all live on this GPU (the peers' arrays of a
    multi-GPU job are just more device pointers): rows row0 .. row0 + B of every set must equal the plain launch's
    record, every other row must stay untouched.  Both forms of the gather: the copy inside the kernel (early-stopped
    batches) and the pipeline of chunked launches + peer_copy_kernel on the side stream (fixed epoch counts; several
    chunks with a ragged last one, an odd row0 so that the 808-byte rows start 8-byte aligned only)."""
    import ctypes as C
    p = BeamOptParams.for_script("MC").replace(num_cases=num_cases)
    if mode == "pipelined":
        p = p.replace(early_stop=False, max_e=7)
        monkeypatch.setenv("OPS_SCATTER_PIPELINED", "1")
    else:
        monkeypatch.setenv("OPS_SCATTER_IN_KERNEL", "1")
    cases = seeded_cases(p, B * num_cases, seed=91)
    fixed, fn, fv, L = sampling.pack_cases(p.num_nodes, p.max_forces, cases, num_cases)
    want = gpu_run(p, fixed, fn, fv, L)
    dev = torch.device("cuda", 0)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)   # noqa: E731
    d_in = [t(fixed), t(fn), t(fv), t(L)]
    nn, n, row0, total = p.num_nodes, p.num_nodes - 1, 3, B + 8
    spec = (("I", (total, n), torch.float32), ("defl", (total, num_cases, nn), torch.float64),
            ("rot", (total, num_cases, nn), torch.float64), ("shear", (total, num_cases, n), torch.float32),
            ("moment", (total, num_cases, n), torch.float32), ("epochs", (total,), torch.int32),
            ("loss", (total,), torch.float32), ("status", (total,), torch.int32))
    sets = [{name: torch.full(shape, -7, dtype=dt, device=dev) for name, shape, dt in spec} for _ in range(3)]
    dests = (_cabi.OpsBeamOptRecordArrays * 3)()
    for i, st in enumerate(sets):
        for f, (name, _, _) in zip(("I_values", "deflections", "rotations", "shear", "moment", "epochs", "loss", "status"), spec):
            setattr(dests[i], f, st[name].data_ptr())
    ip, fp = ops.pack_params(p)
    cp = ops._c_params(ip, fp)
    sched = ops.device_schedule(p, dev)
    lib = _cabi.lib()
    assert lib.ops_beamopt_scatter_supported(C.byref(cp)) == 1
    ws = torch.empty((max(int(lib.ops_beamopt_workspace_bytes(C.byref(cp), B)), 1),), dtype=torch.uint8, device=dev)
    stream = torch.cuda.current_stream(dev)
    rc = lib.ops_beamopt_launch_scatter(C.byref(cp), B, d_in[0].data_ptr(), d_in[1].data_ptr(), d_in[2].data_ptr(),
                                        d_in[3].data_ptr(), sched.data_ptr(), 3, dests, row0, ws.data_ptr(), ws.numel(),
                                        stream.cuda_stream)
    _cabi.check(rc, "ops_beamopt_launch_scatter")
    torch.cuda.synchronize()
    for i, st in enumerate(sets):
        for name, _, _ in spec:
            got = st[name].cpu().numpy()
            assert np.array_equal(got[row0:row0 + B], want[name].reshape(got[row0:row0 + B].shape), equal_nan=True), (i, name)
            assert (got[:row0] == -7).all() and (got[row0 + B:] == -7).all(), (i, name, "rows outside the launch were written")


def test_generate_samples_batched_is_a_drop_in():
    """Same entry point, arguments and record schema as the reference's generate_sample."""
    rollers, avail = sampling.fixed_bridge(101)
    node_positions = np.linspace(0, 200.0, 101)
    recs = generator.generate_samples_batched(range(8), 101, 0, 200.0, node_positions, rollers, avail,
                                              patience=5, params=BeamOptParams.for_script("SC"), seed=0)
    assert len(recs) == 8 and all(r is not None for r in recs)
    assert tuple(recs[0]) == generator.TRAINING_DATA_KEYS
    # seed 0, first sample = the reference's SC seed-0 golden (same stream, same call order)
    m, rec = goldens()[0]
    assert recs[0]["force_nodes"] == m["force_nodes"] and recs[0]["force_values"] == m["force_values"]
    import random
    random.seed(0)
    one = generator.generate_sample(0, 101, 0, 200.0, node_positions, rollers, avail, patience=5,
                                    params=BeamOptParams.for_script("SC"))
    assert one["I_values"] == recs[0]["I_values"]
    data = generator.generate_dataset(generator.GeneratorConfig.multi_core(), num_samples=64, seed=3)
    assert len(data["I_values"]) == 64 and len(data["deflections"][0]) == 101
    assert all(d[-1] == 0.0 for d in data["deflections"])


def test_three_moment_unsupported_roller_count_and_ldlt_fallback():
    p = BeamOptParams.for_script("SC").replace(max_e=5)
    cases = [(200.0, [10, 20, 30, 40, 50, 60], [55], [-1e5])]
    fixed, fn, fv, L = sampling.pack_cases(p.num_nodes, p.max_forces, cases)
    assert gpu_run(p.replace(solver=0), fixed, fn, fv, L)["status"][0] == 3
    assert gpu_run(p.replace(solver=2), fixed, fn, fv, L)["status"][0] == 3
    assert gpu_run(p.replace(solver=3), fixed, fn, fv, L)["status"][0] == 3
    assert gpu_run(p.replace(solver=4), fixed, fn, fv, L)["status"][0] == 3
    b = gpu_run(p.replace(solver=1), fixed, fn, fv, L)
    a = oracle_run(p, fixed, fn, fv, L)
    assert b["status"][0] == 0 and np.max(np.abs(a["I"] - b["I"]) / a["I"]) < 1e-5


@pytest.mark.parametrize("solver", [0, 2])
def test_fine_discretisation_1000_elements(solver):
    """BASELINE config 5 geometry (1001 nodes, rollers x10): the three-moment solve stays within 1e-9 of
    the 80-bit truth where FP64 banded Cholesky cannot (cond(K) ~ 2e10), and the loop matches the oracle."""
    p = BeamOptParams.for_script("MC").replace(num_nodes=1001, max_e=12, solver=solver)
    cases = seeded_cases(p, 64, seed=31, roller_nodes=[100, 300, 700, 850, 1000])
    fixed, fn, fv, L = sampling.pack_cases(p.num_nodes, p.max_forces, cases)
    a, b = oracle_run(p, fixed, fn, fv, L), gpu_run(p, fixed, fn, fv, L)
    assert np.array_equal(a["epochs"], b["epochs"]) and not b["status"].any()
    assert np.max(np.abs(a["I"] - b["I"]) / a["I"]) < 1e-5
    rng = np.random.default_rng(1)
    I = np.exp(rng.uniform(np.log(3e-3), np.log(0.9), (64, 1000))).astype(np.float32).astype(np.float64)
    cp = oracle_params(p)
    o80 = c_oracle.beam_solve(cp, fixed, fn[:, 0], fv[:, 0], L, I, 1)
    g = gpu_solve(p, fixed, fn[:, 0], fv[:, 0], L, I)
    for k in ("defl", "rot", "shear", "moment"):
        assert rel_err(g[k], o80[k]).max() < 1e-9, k


def test_fine_discretisation_full_length_runs():
    """BASELINE config 5 at full length on 256 beams: the early-stopped loop (MultiCore constants) and 600 fixed
    epochs of the production kernel for 1000 elements (one warp per beam).  cond(K) ~ 2e10 puts FP64 banded Cholesky
    3e-8 away from the exact solve per analysis (SURVEY finding 10), so the oracle here is the loop with the 80-bit FE
    solve: identical stop decisions (up to one-ulp margins) and I within 1e-5 on ALL beams after 600 epochs; the FP64
    oracle is reported beside it with the bound "no further from the truth than twice the FP64 oracle itself"."""
    p = BeamOptParams.for_script("MC").replace(num_nodes=1001)
    cases = seeded_cases(p, 256, seed=32, roller_nodes=[100, 300, 700, 850, 1000])
    fixed, fn, fv, L = sampling.pack_cases(p.num_nodes, p.max_forces, cases)
    t, g = oracle_run_mt(p, fixed, fn, fv, L, 1), gpu_run(p, fixed, fn, fv, L)
    same = assert_same_decisions(t, g, 0.01)
    assert np.max(np.abs(t["I"][same] - g["I"][same]) / t["I"][same]) < 1e-5
    assert rel_err(g["moment"][same, 0], t["moment"][same, 0]).max() < 1e-6
    assert rel_err(g["defl"][same, 0], t["defl"][same, 0]).max() < 1e-6
    pf = p.replace(early_stop=False)
    t, g = oracle_run_mt(pf, fixed, fn, fv, L, 1), gpu_run(pf, fixed, fn, fv, L)
    o = oracle_run_mt(pf, fixed, fn, fv, L, 0)
    assert (g["epochs"] == 600).all() and not g["status"].any()
    err80 = np.max(np.abs(t["I"] - g["I"]) / t["I"])
    own = np.max(np.abs(t["I"] - o["I"]) / t["I"])
    err64 = np.max(np.abs(o["I"] - g["I"]) / o["I"])
    assert err80 < 1e-5, (err80, own, err64)
    assert err64 <= 2 * own + 1e-5, (err64, own)


def test_session_matches_the_device_pointer_entry():
    """ops_beamopt_session_*: pinned host arrays in, pinned host arrays out, same bytes as ops_beamopt_launch;
    partial batches and reuse of one session."""
    p = BeamOptParams.for_script("MC").replace(max_e=80)
    cases = seeded_cases(p, 700, seed=106)
    fixed, fn, fv, L = sampling.pack_cases(p.num_nodes, p.max_forces, cases)
    want = gpu_run(p, fixed, fn, fv, L)
    with _cabi.Session(p, 700, device=0) as s:
        for B in (700, 123, 700):
            s.load(fixed[:B], fn[:B], fv[:B], L[:B])
            got = s.run(B)
            for k in want:
                assert np.array_equal(got[k], want[k][:B]), (k, B)
        assert s.kernel_ms > 0


@pytest.mark.parametrize("early_stop", [False, True])
def test_session_pipeline_of_chunked_launches(early_stop):
    """A session run is a pipeline: launches of whole rounds with each chunk's device->host copy on a second
    stream under the next chunk's iterations.  Whatever the chunking (fixed epochs: one round per launch, early
    stopping: six), the pinned outputs hold the bytes of the single device-pointer launch."""
    p = BeamOptParams.for_script("MC").replace(max_e=40 if not early_stop else 600, early_stop=early_stop)
    B = 148 * 40 * 2 + 1776 if not early_stop else 148 * 48 * 6 * 2 + 999        # three launches either way
    cases = seeded_cases(p, 4096, seed=108)
    f0, n0, v0, L0 = sampling.pack_cases(p.num_nodes, p.max_forces, cases)
    rep = (B + 4095) // 4096
    fixed, fn, fv, L = (np.concatenate([a] * rep)[:B] for a in (f0, n0, v0, L0))
    want = gpu_run(p, fixed, fn, fv, L)
    with _cabi.Session(p, B, device=0) as s:
        for Bi in (B, B - 333):
            s.load(fixed[:Bi], fn[:Bi], fv[:Bi], L[:Bi])
            got = s.run(Bi)
            for k in want:
                assert np.array_equal(got[k], want[k][:Bi]), (k, Bi)


@pytest.mark.parametrize("num_cases", [2, 4, 8])
def test_shared_inertia_load_cases(num_cases):
    """BASELINE config 4 (extension): C load cases per beam share one I vector, summed
    energies; one record per (beam, case) with identical I_values."""
    p = BeamOptParams.for_script("MC").replace(num_cases=num_cases)
    B = 512
    cases = seeded_cases(p, B * num_cases, seed=107)
    fixed, fn, fv, L = sampling.pack_cases(p.num_nodes, p.max_forces, cases, num_cases)
    a, b = oracle_run_mt(p, fixed, fn, fv, L), gpu_run(p, fixed, fn, fv, L)
    assert not b["status"].any()
    same = assert_same_decisions(a, b, 0.01)
    assert np.max(np.abs(a["I"][same] - b["I"][same]) / a["I"][same]) < 1e-5
    assert (a["loss"][same] == b["loss"][same]).mean() > 0.95
    ns = int(same.sum())
    for key in ("defl", "rot", "moment", "shear"):
        assert rel_err(b[key][same].reshape(ns * num_cases, -1), a[key][same].reshape(ns * num_cases, -1)).max() < 1e-6, key
    # fixed epochs: I never depends on a stop decision
    pf = p.replace(early_stop=False)
    a, b = oracle_run_mt(pf, fixed, fn, fv, L), gpu_run(pf, fixed, fn, fv, L)
    assert (b["epochs"] == 600).all() and np.max(np.abs(a["I"] - b["I"]) / a["I"]) < 1e-5
    # host records: C consecutive records per beam, identical I_values (the trainers' reshape(total, n_cases, -1))
    rollers, avail = sampling.fixed_bridge(101)
    recs = generator.generate_samples_batched(range(6), 101, 0, 200.0, np.linspace(0, 200.0, 101), rollers, avail,
                                              patience=10, params=p, seed=1)
    assert len(recs) == 6 * num_cases
    for i in range(6):
        grp = recs[i * num_cases:(i + 1) * num_cases]
        assert all(r["I_values"] == grp[0]["I_values"] for r in grp)
        assert len({tuple(r["force_values"]) for r in grp}) > 1


def test_columnar_dataset_and_trainer_preprocessing_on_device(tmp_path):
    """one launch -> columnar arrays -> JSON / npz -> the trainers' dict; the
    trainers' pre-processing block on the GPU."""
    from openpystruct_b200 import dataset
    cfg = generator.GeneratorConfig.multi_core()
    col = generator.generate_columnar(cfg, num_samples=96, seed=3)
    data = generator.generate_dataset(cfg, num_samples=96, seed=3)
    got = dataset.to_training_data(col)
    import json as _json
    for k in generator.TRAINING_DATA_KEYS:
        assert _json.loads(_json.dumps(got[k])) == _json.loads(_json.dumps(data[k], default=float)), k
    dataset.save_npz(col, tmp_path / "d.npz")
    back = dataset.load_training_data(str(tmp_path / "d.npz"))
    assert back["I_values"] == got["I_values"] and back["deflections"] == got["deflections"]
    pre = dataset.trainer_preprocess(back, n_cases=4, c=0.5, seed=0, device="cuda")
    assert pre["X_train"].is_cuda and pre["X_train"].shape[0] == int(0.8 * 24)
    assert pre["Y_train"].shape[1] == 100 + 101 + 101 and torch.isfinite(pre["Y_train"]).all()


def test_random_bridges_against_the_80bit_fe_loop():
    """flag=1 draws short, stiff spans whose K is ill conditioned: there the reference's FP64 banded
    Cholesky (the FP64 oracle) is itself several digits away from the exact solve.  Against the same loop
    with the FE half in 80-bit arithmetic the CUDA path makes identical stop decisions and is at least as
    close as the FP64 oracle is; where it differs from the FP64 oracle, the FP64 oracle differs from the truth."""
    p = BeamOptParams.for_script("SC")
    cases = seeded_cases(p, 2048, seed=2024, flag=1)
    fixed, fn, fv, L = sampling.pack_cases(p.num_nodes, p.max_forces, cases)
    t = oracle_run_mt(p, fixed, fn, fv, L, 1)
    o = oracle_run_mt(p, fixed, fn, fv, L, 0)
    g = gpu_run(p, fixed, fn, fv, L)
    same_g = assert_same_decisions(t, g, 0.002)
    same_o = o["epochs"] == t["epochs"]
    assert (~same_g).sum() <= (~same_o).sum()
    # a stop epoch that differs from the reference arithmetic's is one the reference arithmetic gets "wrong" itself
    assert ((g["epochs"] != o["epochs"]) <= (~same_o | ~same_g)).all()
    both = same_g & same_o
    err_g = np.max(np.abs(g["I"][both] - t["I"][both]) / t["I"][both])
    err_o = np.max(np.abs(o["I"][both] - t["I"][both]) / t["I"][both])
    assert err_g < 1e-5 and err_g <= 2 * err_o + 1e-7, (err_g, err_o)
    assert (g["I"][same_g] == t["I"][same_g]).all(axis=1).mean() > 0.99


@pytest.mark.parametrize("solver", [0, 3, 4])
@pytest.mark.parametrize("num_nodes", [6, 33, 64, 87, 129, 169])
def test_other_discretisations(num_nodes, solver):
    """Every template instance of the lanes kernel (4 / 8 / 13 / 21 element slots per lane, run-time n) and of the
    shared-memory-state kernels."""
    p = BeamOptParams.for_script("SC").replace(num_nodes=num_nodes, max_e=60, solver=solver)
    n = num_nodes - 1
    rollers = sorted({max(2, int(round(f * n))) for f in (0.1, 0.3, 0.7, 0.85)} | {n})
    cases = seeded_cases(p, 300, seed=num_nodes, roller_nodes=rollers)
    fixed, fn, fv, L = sampling.pack_cases(p.num_nodes, p.max_forces, cases)
    a, b = oracle_run(p, fixed, fn, fv, L), gpu_run(p, fixed, fn, fv, L)
    assert not b["status"].any()
    same = assert_same_decisions(a, b, 0.01)
    assert np.max(np.abs(a["I"][same] - b["I"][same]) / a["I"][same]) < 1e-5
    assert (a["loss"][same] == b["loss"][same]).mean() > 0.95
    assert rel_err(b["moment"][same, 0], a["moment"][same, 0]).max() < 1e-6
    assert rel_err(b["defl"][same, 0], a["defl"][same, 0]).max() < 1e-6


def test_native_sampler_paths_and_the_pipelined_stream():
    """The native sampler draws the stream of `random` (CPU test): through the GPU the columnar dataset of
    generate_columnar(seed) equals the one built from the Python sampler's cases, and the pipelined batches of
    stream_columnar concatenate to it -- with copies and with views of the session's pinned buffers."""
    cfg = generator.GeneratorConfig.multi_core()
    p = cfg.params.replace(max_e=60)
    cfg = generator.GeneratorConfig(params=p)
    want = generator.generate_columnar(cfg, num_samples=1000, seed=11, native_sampler=False)
    got = generator.generate_columnar(cfg, num_samples=1000, seed=11)
    for k in want:
        assert np.array_equal(np.asarray(want[k]), np.asarray(got[k]), equal_nan=True), k
    for reuse in (False, True):
        parts = []
        for col in generator.stream_columnar(cfg, num_samples=1000, batch_size=384, seed=11, reuse_buffers=reuse):
            parts.append({k: np.array(v, copy=True) for k, v in col.items()})
        assert [len(c["L"]) for c in parts] == [384, 384, 232]
        for k in ("I_values", "deflections", "rotations", "shear_forces", "bending_moments", "L", "node_positions"):
            assert np.array_equal(np.concatenate([c[k] for c in parts]), np.asarray(want[k])), (k, reuse)
        fv = np.concatenate([np.pad(c["force_values"], ((0, 0), (0, 4 - c["force_values"].shape[1])), constant_values=np.nan)
                             for c in parts])
        wv = np.pad(want["force_values"], ((0, 0), (0, 4 - want["force_values"].shape[1])), constant_values=np.nan)
        assert np.array_equal(fv, wv, equal_nan=True)


def test_more_than_five_rollers_are_rerun_with_the_band_solver():
    """The reference's ops.fix loop takes any number of rollers (SingleCore:101-102); the three-moment kernels take five.
    The host entry re-runs such beams with the banded LDL^T solver instead of dropping them."""
    p = BeamOptParams.for_script("SC").replace(max_e=30)
    cases = seeded_cases(p, 6, seed=9) + [(200.0, [10, 20, 30, 40, 50, 60], [55], [-1e5]),
                                          (200.0, [5, 15, 25, 45, 65, 85, 100], [50, 70], [-2e5, -1e5])]
    out = generator.optimise_cases(p, cases)
    assert not out["status"].any()
    fixed, fn, fv, L = sampling.pack_cases(p.num_nodes, p.max_forces, cases)
    a = oracle_run(p, fixed, fn, fv, L)
    assert np.array_equal(a["epochs"], out["epochs"])
    assert np.max(np.abs(a["I"] - out["I"]) / a["I"]) < 1e-5
    recs = generator.make_records(p, cases, out)
    assert all(r is not None for r in recs) and recs[-1]["roller_nodes"] == [5, 15, 25, 45, 65, 85, 100]


def test_the_ctypes_stub_of_INTEGRATION_md_runs_as_printed():
    """INTEGRATION.md section B is what a maintainer of the reference pastes into a script: the code block is taken
    from the document as printed, executed (no package import, no torch), and its records compared with the package path."""
    import os
    import re
    from tests.helpers import ROOT
    text = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    block = next(b for b in re.findall(r"```python\n(.*?)```", text, flags=re.S) if "class BeamOptimiser" in b)
    ns = {}
    cwd = os.getcwd()
    os.chdir(ROOT)
    try:
        exec(compile(block, "INTEGRATION.md", "exec"), ns)
        p = BeamOptParams.for_script("SC")
        cases = seeded_cases(p, 300, seed=21)
        opt = ns["BeamOptimiser"](max_beams=512)
        got = {k: np.array(v, copy=True) for k, v in opt(cases).items()}
        again = opt(cases[:77])                               # a second batch on the same session
        assert np.array_equal(again["I"], got["I"][:77])
        opt.close()
    finally:
        os.chdir(cwd)
    fixed, fn, fv, L = sampling.pack_cases(p.num_nodes, p.max_forces, cases)
    want = gpu_run(p, fixed, fn, fv, L)
    for k in want:
        assert np.array_equal(got[k], want[k]), k
