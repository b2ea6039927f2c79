"""The kernel's per-beam arithmetic (openpystruct_b200/csrc/beamopt_core.cuh) compiled for the host
(tests/hostsim -- debug aid, not a product path) against the CPU oracle.  This is how the exact
operation order the GPU executes is validated in the build container, which has no GPU; the same
comparisons run against the real kernel in tests/test_gpu_parity.py."""
import numpy as np
import pytest

from oracle import c_oracle
from openpystruct_b200 import sampling
from openpystruct_b200.params import BeamOptParams
from tests.helpers import (goldens, golden_params, golden_case, hostsim_run, hostsim_solve, oracle_params,
                           oracle_run, rel_err, seeded_cases)


SOLVERS = [pytest.param(0, id="three_moment_lanes"), pytest.param(1, id="band_ldlt"),
           pytest.param(2, id="three_moment_thread"), pytest.param(3, id="three_moment_smem8"),
           pytest.param(4, id="three_moment_smem32")]


@pytest.mark.parametrize("solver", SOLVERS)
@pytest.mark.parametrize("script,flag,count", [("SC", 0, 200), ("MC", 0, 200), ("GPU", 0, 40), ("SC", 1, 200)])
def test_full_loop_against_c_oracle(script, flag, count, solver):
    p = BeamOptParams.for_script(script).replace(solver=solver)
    cases = seeded_cases(p, count, seed=11, flag=flag)
    fixed, fn, fv, L = sampling.pack_cases(p.num_nodes, p.max_forces, cases)
    a = oracle_run(p, fixed, fn, fv, L)
    b = hostsim_run(p, fixed, fn, fv, L)
    assert not a["status"].any() and not b["status"].any()
    same = a["epochs"] == b["epochs"]
    # identical early-stop decisions (both sides IEEE fp32; M,V agree to ~1e-11 before the fp32 cast)
    assert same.mean() >= (0.99 if flag == 0 else 0.97)
    tol = 1e-5
    assert np.max(np.abs(a["I"][same] - b["I"][same]) / a["I"][same]) < tol
    assert (a["I"][same] == b["I"][same]).mean() > 0.98
    assert rel_err(b["defl"][same, 0], a["defl"][same, 0]).max() < (1e-7 if flag == 0 else 1e-5)
    assert np.array_equal(a["loss"][same] == b["loss"][same], np.ones(same.sum(), bool)) or \
        (a["loss"][same] == b["loss"][same]).mean() > 0.98


@pytest.mark.parametrize("solver", SOLVERS)
def test_fixed_epoch_mode_against_c_oracle(solver):
    p = BeamOptParams.for_script("MC").replace(early_stop=False, max_e=600, solver=solver)
    cases = seeded_cases(p, 50, seed=12)
    fixed, fn, fv, L = sampling.pack_cases(p.num_nodes, p.max_forces, cases)
    a = oracle_run(p, fixed, fn, fv, L)
    b = hostsim_run(p, fixed, fn, fv, L)
    assert (a["epochs"] == 600).all() and (b["epochs"] == 600).all()
    assert np.max(np.abs(a["I"] - b["I"]) / a["I"]) < 1e-5
    assert (a["defl"][:, 0, -1] == 0).all() and (b["defl"][:, 0, -1] == 0).all()     # MultiCore:222-223


@pytest.mark.parametrize("solver", SOLVERS)
def test_single_solve_1e9_on_default_bridge_and_vs_truth(solver):
    p = BeamOptParams(solver=solver)
    rng = np.random.default_rng(0)
    cases = seeded_cases(p, 200, seed=13)
    fixed, fn, fv, L = sampling.pack_cases(p.num_nodes, p.max_forces, cases)
    I = np.exp(rng.uniform(np.log(3e-3), np.log(0.9), (200, 100))).astype(np.float32).astype(np.float64)
    cp = oracle_params(p)
    o64 = c_oracle.beam_solve(cp, fixed, fn[:, 0], fv[:, 0], L, I, 0)
    o80 = c_oracle.beam_solve(cp, fixed, fn[:, 0], fv[:, 0], L, I, 1)
    h = hostsim_solve(p, fixed, fn[:, 0], fv[:, 0], L, I)
    for k in ("defl", "rot", "shear", "moment"):
        assert rel_err(h[k], o64[k]).max() < 1e-9, k      # north_star tolerance vs the dpbsv restatement
        assert rel_err(h[k], o80[k]).max() < (5e-10 if solver == 1 else 1e-11), k   # vs the 80-bit truth


@pytest.mark.parametrize("solver", SOLVERS)
def test_single_solve_on_reference_goldens(solver):
    for m, rec in goldens():
        p = golden_params(m).replace(solver=solver)
        fixed, fn, fv, L = sampling.pack_cases(p.num_nodes, p.max_forces, [golden_case(m)])
        h = hostsim_solve(p, fixed, fn[:, 0], fv[:, 0], L, rec["I_last"][None, :])
        tol = 1e-9 if m["flag"] == 0 else 1e-6
        assert rel_err(h["moment"][0], rec["M64_last"]) < tol
        assert rel_err(h["shear"][0], rec["V64_last"]) < tol


@pytest.mark.parametrize("solver", SOLVERS)
def test_goldens_full_loop(solver):
    same = 0
    for m, rec in goldens():
        p = golden_params(m).replace(solver=solver)
        fixed, fn, fv, L = sampling.pack_cases(p.num_nodes, p.max_forces, [golden_case(m)])
        b = hostsim_run(p, fixed, fn, fv, L)
        if b["epochs"][0] == m["epochs"]:
            same += 1
            assert np.max(np.abs(b["I"][0] - rec["I_values"]) / rec["I_values"]) < 1e-5
            assert rel_err(b["moment"][0, 0], rec["bending_moments"]) < 1e-6
    assert same >= 0.8 * len(goldens())


@pytest.mark.parametrize("solver", SOLVERS)
def test_edge_cases(solver):
    p = BeamOptParams.for_script("SC").replace(max_e=40, solver=solver)
    cases = [
        (200.0, [10, 30, 70, 85, 100], [], []),                       # UDL only
        (200.0, [101], [51], [-1e5]),                                 # single span, roller at the tip node
        (15.0, [2], [3, 4, 5, 6], [-3.5e5] * 4),                      # shortest random bridge, roller next to the pin
        (215.0, [100], [2, 50, 99, 60], [-3e5, -2e5, -1e5, -5e4]),   # longest, one roller
        (200.0, [10, 30, 70, 85, 100], [50, 50], [-1e5, -1e5]),       # two loads on one node accumulate
        (200.0, [10, 11, 12, 100, 101], [5, 11, 60], [-1e5, -2e5, -3e5]),  # adjacent rollers, load on a roller
        (200.0, [50], [51, 100, 101], [-1e5, -1e5, -5e4]),            # long overhang with tip load
    ]
    fixed, fn, fv, L = sampling.pack_cases(p.num_nodes, p.max_forces, cases)
    a = oracle_run(p, fixed, fn, fv, L)
    b = hostsim_run(p, fixed, fn, fv, L)
    assert np.array_equal(a["epochs"], b["epochs"]) and not b["status"].any()
    assert np.max(np.abs(a["I"] - b["I"]) / a["I"]) < 1e-5
    # mechanism: pin only (no roller) is singular -> status 1 on both sides, never a crash
    fixed2, fn2, fv2, L2 = sampling.pack_cases(p.num_nodes, p.max_forces, [(200.0, [], [50], [-1e5])])
    assert oracle_run(p, fixed2, fn2, fv2, L2)["status"][0] == 1
    assert hostsim_run(p, fixed2, fn2, fv2, L2)["status"][0] == 1


def test_three_moment_reports_unsupported_roller_count():
    p = BeamOptParams.for_script("SC").replace(max_e=5)
    cases = [(200.0, [10, 20, 30, 40, 50, 60], [55], [-1e5])]          # 6 rollers > FLEX_MAXS - 1
    fixed, fn, fv, L = sampling.pack_cases(p.num_nodes, p.max_forces, cases)
    assert hostsim_run(p, fixed, fn, fv, L, solver=0)["status"][0] == 3
    b = hostsim_run(p, fixed, fn, fv, L, solver=1)
    a = oracle_run(p, fixed, fn, fv, L)
    assert b["status"][0] == 0 and np.max(np.abs(a["I"] - b["I"]) / a["I"]) < 1e-5


@pytest.mark.parametrize("solver", [2, 4])
def test_fine_discretisation_1000_elements_three_moment(solver):
    """BASELINE config 5 geometry: 1001 nodes, rollers scaled x10; exercises the torch.sum level cascade
    (n >= 512) and the O(#supports) state.  Compared with the 80-bit truth / the fp32 oracle loop."""
    p = BeamOptParams.for_script("MC").replace(num_nodes=1001, max_e=12, solver=solver)
    cases = seeded_cases(p, 6, seed=31, roller_nodes=[100, 300, 700, 850, 1000])
    fixed, fn, fv, L = sampling.pack_cases(p.num_nodes, p.max_forces, cases)
    a = oracle_run(p, fixed, fn, fv, L)
    b = hostsim_run(p, fixed, fn, fv, L)
    assert np.array_equal(a["epochs"], b["epochs"])
    assert np.max(np.abs(a["I"] - b["I"]) / a["I"]) < 1e-5
    rng = np.random.default_rng(1)
    I = np.exp(rng.uniform(np.log(3e-3), np.log(0.9), (6, 1000))).astype(np.float32).astype(np.float64)
    cp = oracle_params(p)
    o80 = c_oracle.beam_solve(cp, fixed, fn[:, 0], fv[:, 0], L, I, 1)
    h = hostsim_solve(p, fixed, fn[:, 0], fv[:, 0], L, I)
    for k in ("defl", "rot", "shear", "moment"):
        assert rel_err(h[k], o80[k]).max() < 1e-9, k


@pytest.mark.parametrize("num_cases", [2, 4, 8])
def test_shared_inertia_load_cases_against_c_oracle(num_cases):
    """SURVEY 8a row 15 (extension): C load cases share one I vector, energies summed in case order.  The
    lanes kernel runs the cases on C adjacent groups that exchange M^2, V^2; the host model runs the
    same phase functions team by team."""
    p = BeamOptParams.for_script("MC").replace(num_cases=num_cases)
    cases = seeded_cases(p, 24 * num_cases, seed=21)
    fixed, fn, fv, L = sampling.pack_cases(p.num_nodes, p.max_forces, cases, num_cases)
    a = oracle_run(p, fixed, fn, fv, L)
    b = hostsim_run(p, fixed, fn, fv, L, solver=0)
    assert not a["status"].any() and not b["status"].any()
    assert np.array_equal(a["epochs"], b["epochs"])
    assert np.max(np.abs(a["I"] - b["I"]) / a["I"]) < 1e-5
    assert (a["loss"] == b["loss"]).mean() > 0.95
    B = len(L)
    for key in ("defl", "rot", "moment", "shear"):
        assert rel_err(b[key].reshape(B * num_cases, -1), a[key].reshape(B * num_cases, -1)).max() < 1e-6, key
    assert (b["defl"][:, :, -1] == 0).all()


@pytest.mark.parametrize("solver", [0, 3, 4])
@pytest.mark.parametrize("num_nodes", [6, 33, 64, 87, 129, 169])
def test_other_discretisations_against_c_oracle(num_nodes, solver):
    """Every template instance of the lanes kernel (4 / 8 / 13 / 21 element slots per lane) and the torch.sum
    shapes that go with them (tails, left-over vectors, no full block at all)."""
    p = BeamOptParams.for_script("SC").replace(num_nodes=num_nodes, max_e=60)
    n = num_nodes - 1
    rollers = sorted({max(2, int(round(f * n))) for f in (0.1, 0.3, 0.7, 0.85)} | {n})
    cases = seeded_cases(p, 40, seed=num_nodes, roller_nodes=rollers)
    fixed, fn, fv, L = sampling.pack_cases(p.num_nodes, p.max_forces, cases)
    a = oracle_run(p, fixed, fn, fv, L)
    b = hostsim_run(p, fixed, fn, fv, L, solver=solver)
    assert not a["status"].any() and not b["status"].any()
    assert np.array_equal(a["epochs"], b["epochs"])
    assert np.max(np.abs(a["I"] - b["I"]) / a["I"]) < 1e-5
    assert (a["loss"] == b["loss"]).mean() > 0.95
    assert rel_err(b["moment"][:, 0], a["moment"][:, 0]).max() < 1e-6


@pytest.mark.parametrize("num_nodes,num_cases", [(101, 1), (101, 4), (34, 1), (88, 2)])
def test_peer_copy_of_a_record_moves_exactly_its_rows(num_nodes, num_cases):
    """lane_copy_record (the in-kernel dataset gather): the eight lanes copy one beam's rows of all record arrays
    into the second destination -- whatever the alignment of the row (16-, 8- and 4-byte units) -- and nothing else."""
    import ctypes as C
    from tests.helpers import hostsim_lib
    hs = hostsim_lib()
    rng = np.random.default_rng(5)
    nn, n, B = num_nodes, num_nodes - 1, 7
    shapes = {"I": ((B, n), np.float32), "defl": ((B, num_cases, nn), np.float64), "rot": ((B, num_cases, nn), np.float64),
              "shear": ((B, num_cases, n), np.float32), "moment": ((B, num_cases, n), np.float32),
              "epochs": ((B,), np.int32), "loss": ((B,), np.float32), "status": ((B,), np.int32)}
    src = {k: (rng.standard_normal(sh) * 100).astype(dt) for k, (sh, dt) in shapes.items()}
    dst = {k: np.full(sh, -7, dt) for k, (sh, dt) in shapes.items()}
    want = {k: v.copy() for k, v in dst.items()}
    order = ("I", "defl", "rot", "shear", "moment", "epochs", "loss", "status")
    for row in (0, 3, 6):
        for c in range(num_cases):
            rowc = row * num_cases + c
            rc = hs.hostsim_copy_record(n, nn, C.c_int64(row), C.c_int64(rowc), int(c == 0),
                                        *[C.c_void_p(src[k].ctypes.data) for k in order],
                                        *[C.c_void_p(dst[k].ctypes.data) for k in order])
            assert rc == 0
            for k in ("defl", "rot", "shear", "moment"):
                want[k][row, c] = src[k][row, c]
        for k in ("I", "epochs", "loss", "status"):
            want[k][row] = src[k][row]
    for k in order:
        assert np.array_equal(want[k], dst[k]), k


def test_batch_size_of_the_stage_major_batches_does_not_change_the_bits(tmp_path):
    """The number of slot pairs per stage-major batch of the pass (three by default; a build knob,
    OPS_LANES_NBP) is a pure reordering of independent per-element chains: the same source built with two pairs per
    batch produces the bits of the default build (single case and four load cases)."""
    import ctypes as C
    import subprocess
    from tests import helpers
    base = helpers.hostsim_lib()
    libs = {}
    for nb in (2,):
        path = str(tmp_path / f"libhostsim_nb{nb}.so")
        subprocess.run(["g++", "-O1", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off", "-Wno-unknown-pragmas",
                        f"-DOPS_LANES_NBP={nb}", "-o", path, helpers._HS_SRC], check=True)
        libs[nb] = C.CDLL(path)
    for num_cases, beams in ((1, 24), (4, 8)):
        p = BeamOptParams.for_script("MC").replace(num_cases=num_cases, max_e=150)
        cases = helpers.seeded_cases(p, beams * num_cases, seed=41)
        fixed, fn, fv, L = sampling.pack_cases(p.num_nodes, p.max_forces, cases, num_cases)
        want = hostsim_run(p, fixed, fn, fv, L, solver=0)
        for nb, lib in libs.items():
            helpers._hs, keep = lib, helpers._hs
            try:
                got = hostsim_run(p, fixed, fn, fv, L, solver=0)
            finally:
                helpers._hs = keep
            for k in want:
                assert np.array_equal(want[k], got[k]), (nb, num_cases, k)
    assert helpers._hs is base


@pytest.mark.parametrize("nbp", [1, 2, 3])
@pytest.mark.parametrize("script,flag,early,max_e,count", [("MC", 0, True, None, 61), ("SC", 1, True, None, 45),
                                                           ("MC", 0, False, 150, 22), ("SC", 0, True, 0, 12)])
def test_tensor_memory_instance_equals_the_register_instance(script, flag, early, max_e, count, nbp):
    """beamopt_lanes_tm.cu keeps {M0, Q0}, m, v of a lane in tensor memory and enters every phase that touches it with
    the whole warp, idle groups included.  The host model runs ONE WARP of that kernel -- four groups pulling beams from
    a counter, stopping at different epochs, a plain word array per lane as the tensor memory -- and must reproduce the
    register / shared-memory instance bit for bit: ragged early stopping (fresh beams next to running ones: the
    read-modify-write commit), parked epochs that go on, rejected beams (mechanism, > 5 rollers) next to running ones,
    zero epochs, any batch size of the pass."""
    p = BeamOptParams.for_script(script).replace(early_stop=early)
    if max_e is not None:
        p = p.replace(max_e=max_e)
    cases = seeded_cases(p, count, seed=23, flag=flag)
    # rejected beams in between: a mechanism (no roller) and an unsupported support count
    cases.insert(3, (200.0, [], [50], [-1e5]))
    cases.insert(9, (200.0, [10, 20, 30, 40, 50, 60], [55], [-1e5]))
    fixed, fn, fv, L = sampling.pack_cases(p.num_nodes, p.max_forces, cases)
    want = hostsim_run(p, fixed, fn, fv, L, solver=0)
    got = hostsim_run(p, fixed, fn, fv, L, solver="tm", tm_nbp=nbp)
    assert want["status"][3] == 1 and want["status"][9] == 3
    if early and max_e is None:
        assert len(set(want["epochs"].tolist())) > 10          # the groups really stop at different epochs
    for key in want:
        assert np.array_equal(want[key], got[key], equal_nan=True), key
