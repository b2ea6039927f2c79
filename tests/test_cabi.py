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
