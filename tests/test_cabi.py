"""The C-ABI shared library: loads, exports every symbol include/openpystruct_b200.h declares, host-only
entry points work, and compute entries fail loudly (no CPU fallback) when there is no CUDA device."""
import ctypes as C
import os
import re

import numpy as np
import pytest
import torch

from openpystruct_b200 import _cabi, build
from openpystruct_b200.params import BeamOptParams
from oracle import c_oracle
from tests.helpers import ROOT, oracle_params

HEADER = os.path.join(ROOT, "include", "openpystruct_b200.h")


def declared_functions():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(ops_[a-z_0-9]+)\s*\(", text)))


def test_library_is_built_in_tree():
    path = build.build()
    assert os.path.exists(path)
    assert os.path.dirname(path).startswith(ROOT)


def test_exports_every_declared_symbol():
    lib = _cabi.lib()
    names = declared_functions()
    assert set(names) >= {"ops_beamopt_launch", "ops_beamsolve_launch", "ops_beamopt_run_host",
                          "ops_beamopt_workspace_bytes", "ops_beamopt_fill_schedule", "ops_beamopt_version",
                          "ops_device_count", "ops_set_device"}
    for name in names:
        assert hasattr(lib, name), name
    assert set(_cabi.EXPORTS) == set(names)


def test_version_and_struct_layout():
    assert "sm_100a" in _cabi.version()
    assert C.sizeof(_cabi.OpsBeamOptParams) == 10 * 4 + 15 * 8
    assert C.sizeof(_cabi.OpsBeamOptParams) == C.sizeof(c_oracle.Params)


def test_launch_plans_on_a_b200():
    """ops_beamopt_plan is host arithmetic (148 SMs, 232 448 bytes of opt-in shared memory given: no device touched):
    which kernel family serves which configuration, and that every plan fits the SM."""
    p = BeamOptParams.for_script("MC").replace(early_stop=False)
    smem, sms = _cabi.B200_SMEM_OPTIN, _cabi.B200_SMS
    # BASELINE configs[1]: two rounds of the 320-thread register instance; one CTA per SM
    a = _cabi.launch_plan(p, 10000)
    assert (a["family"], a["threads"], a["blocks"], a["beams_per_cta"], a["scatter"]) == ("lanes", 320, 148, 40, 1)
    # a fixed-epoch batch that is ONE round of 52..64 beams per SM: the tensor-memory instance, 16 warps
    b = _cabi.launch_plan(p, 148 * 64)
    assert (b["family"], b["threads"], b["beams_per_cta"]) == ("lanes_tm", 512, 64)
    assert _cabi.launch_plan(p.replace(early_stop=True), 148 * 64)["family"] == "lanes"     # ragged stopping: registers
    # many rounds: twelve warps (three per scheduler)
    c = _cabi.launch_plan(p, 1000000)
    assert (c["family"], c["threads"], c["beams_per_cta"]) == ("lanes", 384, 48)
    # a small batch is spread over as many SMs as it has beams
    assert _cabi.launch_plan(p, 100)["blocks"] == 100
    # eight load cases: teams of eight groups (64 threads per beam), five teams per CTA
    d = _cabi.launch_plan(p.replace(num_cases=8), 100000)
    assert (d["family"], d["lanes_per_beam"], d["beams_per_cta"], d["scatter"]) == ("lanes", 64, 5, 1)
    # 1000-element beams: one warp per beam, TWELVE beams per SM (the segment statics live in registers)
    e = _cabi.launch_plan(p.replace(num_nodes=1001), 100000)
    assert (e["family"], e["lanes_per_beam"], e["threads"], e["beams_per_cta"], e["scatter"]) == ("wide", 32, 384, 12, 0)
    # the literal banded LDL^T and the thread-per-beam three-moment kernel behind the same ABI
    assert _cabi.launch_plan(p.replace(solver=1), 10000)["family"] == "thread_ldlt"
    assert _cabi.launch_plan(p.replace(solver=2), 10000)["family"] == "thread_three_moment"
    # every supported discretisation / case count: whole teams per CTA, whole warps, inside the SM's shared memory
    for nn in (5, 33, 64, 65, 101, 105, 106, 169, 170, 400, 1001, 2001):
        for nc in (1, 2, 4, 8):
            q = p.replace(num_nodes=nn, num_cases=nc)
            if nc > 1 and nn > 105:
                with pytest.raises(_cabi.CudaLibraryError):
                    _cabi.launch_plan(q, 5000)
                continue
            for B in (1, 37, 5000, 200000):
                pl = _cabi.launch_plan(q, B)
                assert pl["smem_bytes"] <= smem and 1 <= pl["blocks"] <= sms, (nn, nc, B, pl)
                assert pl["threads"] % 32 == 0 and pl["threads"] % pl["lanes_per_beam"] == 0, (nn, nc, B, pl)
                assert pl["threads"] <= 1024 and pl["beams_per_cta"] >= 1
                assert pl["family"] == ("wide" if nn > 169 else pl["family"])
    assert _cabi.launch_plan(p, 10000)["workspace_bytes"] == _cabi.lib().ops_beamopt_workspace_bytes(
        C.byref(_cabi.to_c_params(p)), 10000) or not torch.cuda.is_available()


def test_schedule_is_host_side_and_matches_torch_and_oracle():
    p = BeamOptParams()
    tab = _cabi.fill_schedule(p)
    assert tab.shape == (600, 2)
    assert np.array_equal(tab, c_oracle.adam_schedule(oracle_params(p)))
    lr = 0.01
    for t in range(1, 601):
        assert tab[t - 1, 0] == np.float32(-(lr / (1 - 0.9 ** t)))
        assert tab[t - 1, 1] == np.float32((1 - 0.999 ** t) ** 0.5)
        lr *= 0.98


def test_bad_struct_size_is_rejected():
    cp = _cabi.to_c_params(BeamOptParams())
    cp.struct_size = 12
    tab = np.zeros(1200, np.float32)
    assert _cabi.lib().ops_beamopt_fill_schedule(C.byref(cp), tab.ctypes.data) == -1


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU behaviour")
def test_compute_fails_loudly_without_a_gpu():
    assert _cabi.lib().ops_device_count() == 0
    p = BeamOptParams()
    fixed = np.zeros((1, 101), np.uint8)
    fn = -np.ones((1, 1, 4), np.int32)
    fv = np.zeros((1, 1, 4))
    with pytest.raises(_cabi.CudaLibraryError):
        _cabi.run_host(p, fixed, fn, fv, np.array([200.0]))
    from openpystruct_b200 import generate_samples_batched, sampling
    rollers, avail = sampling.fixed_bridge()
    with pytest.raises(RuntimeError, match="CUDA"):
        generate_samples_batched(range(2), 101, 0, 200.0, None, rollers, avail, patience=5, seed=0)
    with pytest.raises(RuntimeError, match="CUDA"):
        generate_samples_batched(range(2), 101, 0, 200.0, None, rollers, avail, patience=5, seed=0, device="cpu")


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU behaviour")
def test_custom_op_has_no_cpu_kernel():
    from openpystruct_b200 import ops
    p = BeamOptParams()
    ip, fp = ops.pack_params(p)
    with pytest.raises((NotImplementedError, RuntimeError)):
        torch.ops.openpystruct.beam_opt(torch.zeros((1, 101), dtype=torch.uint8),
                                        torch.zeros((1, 1, 4), dtype=torch.int32),
                                        torch.zeros((1, 1, 4), dtype=torch.float64),
                                        torch.full((1,), 200.0, dtype=torch.float64),
                                        torch.zeros((600, 2)), ip, fp)


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "openpystruct_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                text = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in text.lower().replace("no cpu fallback", ""), os.path.join(dirpath, f)
                assert "hostsim" not in text or f.endswith(".cuh"), os.path.join(dirpath, f)


def test_no_packed_product_was_contracted_into_a_packed_sum():
    """ptxas fuses mul.rn.f32x2 + add.rn.f32x2 into one FFMA2 even under -fmad=false, which would change torch's
    rounding; the lanes kernel is written so that it cannot happen, and scripts/sass_packed_audit.py proves it for the
    built library by counting the packed instructions of every kernel instance (no GPU needed: cuobjdump)."""
    import shutil
    import subprocess
    import sys
    if shutil.which("cuobjdump") is None:
        pytest.skip("cuobjdump not on PATH")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "sass_packed_audit.py"), build.build()],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "packed audit: PASS" in r.stdout, r.stdout[-3000:]
    assert r.stdout.count("ok ") >= 20
