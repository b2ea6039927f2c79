"""ctypes binding of include/openpystruct_b200.h.  There is no CPU implementation behind it:
a missing library or a missing CUDA device raises."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from .build import LIB_PATH
from .params import BeamOptParams

_lib = None

EXPORTS = (
    "ops_beamopt_version", "ops_device_count", "ops_set_device", "ops_beamopt_fill_schedule",
    "ops_beamopt_workspace_bytes", "ops_beamopt_launch", "ops_beamsolve_launch", "ops_beamopt_run_host",
    "ops_fp64_peak_probe", "ops_fastmath_selftest", "ops_pipe_probe",
    "ops_beamopt_session_create", "ops_beamopt_session_arrays", "ops_beamopt_session_run",
    "ops_beamopt_session_destroy",
    "ops_beamopt_plan",
    "ops_beamopt_launch_scatter", "ops_beamopt_scatter_supported", "ops_peer_alloc", "ops_peer_open", "ops_peer_close", "ops_peer_free",
    "ops_sampler_create", "ops_sampler_destroy", "ops_sampler_random", "ops_sampler_randint", "ops_sampler_draw_cases",
    "ops_frameopt_max_elements", "ops_frameopt_fill_schedule", "ops_frameopt_launch", "ops_frameopt_run_host",
)


class OpsBeamOptParams(C.Structure):
    _fields_ = [
        ("struct_size", C.c_int32), ("num_nodes", C.c_int32), ("num_cases", C.c_int32),
        ("max_forces", C.c_int32), ("max_epochs", C.c_int32), ("patience", C.c_int32),
        ("early_stop", C.c_int32), ("zero_last_node", C.c_int32), ("solver", C.c_int32),
        ("reserved", C.c_int32),
        ("E", C.c_double), ("G", C.c_double), ("udl", C.c_double), ("I0", C.c_double),
        ("lr", C.c_double), ("gamma", C.c_double), ("alpha_moment", C.c_double),
        ("alpha_shear", C.c_double), ("tolerance", C.c_double), ("shear_k", C.c_double),
        ("bending_eps", C.c_double), ("clamp_min", C.c_double), ("beta1", C.c_double),
        ("beta2", C.c_double), ("adam_eps", C.c_double),
    ]


class OpsFrameOptParams(C.Structure):
    _fields_ = [
        ("struct_size", C.c_int32), ("max_bays", C.c_int32), ("max_stories", C.c_int32), ("max_epochs", C.c_int32),
        ("patience", C.c_int32), ("early_stop", C.c_int32),
        ("E", C.c_double), ("G", C.c_double), ("A", C.c_double), ("I0", C.c_double), ("alpha_moment", C.c_double),
        ("alpha_shear", C.c_double), ("shear_k", C.c_double), ("bending_eps", C.c_double),
        ("lateral_load", C.c_double), ("vertical_load", C.c_double), ("lr", C.c_double), ("tolerance", C.c_double),
        ("bay_width", C.c_double), ("story_height", C.c_double), ("clamp_min", C.c_double), ("beta1", C.c_double),
        ("beta2", C.c_double), ("adam_eps", C.c_double),
    ]


class OpsBeamOptHostArrays(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("fixed_uy", "force_nodes", "force_vals", "L", "I_values", "deflections",
                                          "rotations", "shear", "moment", "epochs", "loss", "status")]


class OpsBeamOptRecordArrays(C.Structure):
    """One set of dataset arrays (device pointers) of ops_beamopt_launch_scatter."""
    _fields_ = [(n, C.c_void_p) for n in ("I_values", "deflections", "rotations", "shear", "moment", "epochs", "loss",
                                          "status")]


class OpsLaunchPlanInfo(C.Structure):
    _fields_ = [("family", C.c_int32), ("threads", C.c_int32), ("blocks", C.c_int32), ("lanes_per_beam", C.c_int32),
                ("beams_per_cta", C.c_int32), ("scatter", C.c_int32), ("smem_bytes", C.c_int64),
                ("workspace_bytes", C.c_int64)]


PLAN_FAMILIES = ("lanes", "lanes_tm", "wide", "thread_three_moment", "thread_ldlt")
B200_SMS, B200_SMEM_OPTIN = 148, 232448


class CudaLibraryError(RuntimeError):
    pass


def to_c_params(p: BeamOptParams) -> OpsBeamOptParams:
    return OpsBeamOptParams(
        C.sizeof(OpsBeamOptParams), p.num_nodes, p.num_cases, p.max_forces, p.max_e, p.patience,
        int(p.early_stop), int(p.zero_last_node), int(p.solver), 0, p.E, p.G, p.uniform_udl, p.I_0, p.lr, p.gamma,
        p.alpha_moment, p.alpha_shear, p.tolerance, p.shear_k, p.bending_eps, p.clamp_min,
        p.beta1, p.beta2, p.adam_eps)


def lib():
    """The loaded CUDA library; raises CudaLibraryError when it has not been built."""
    global _lib
    if _lib is None:
        path = os.environ.get("OPS_B200_LIB", LIB_PATH)      # A/B builds of the same ABI (profiling)
        if not os.path.exists(path):
            raise CudaLibraryError(
                f"{path} is missing: build it with `python -m openpystruct_b200.build` "
                "(there is no CPU fallback for this path)")
        L = C.CDLL(path)
        L.ops_beamopt_version.restype = C.c_char_p
        L.ops_device_count.restype = C.c_int
        L.ops_set_device.argtypes = [C.c_int]
        L.ops_beamopt_fill_schedule.argtypes = [C.POINTER(OpsBeamOptParams), C.c_void_p]
        L.ops_beamopt_workspace_bytes.restype = C.c_size_t
        L.ops_beamopt_workspace_bytes.argtypes = [C.POINTER(OpsBeamOptParams), C.c_int64]
        L.ops_beamopt_launch.argtypes = [C.POINTER(OpsBeamOptParams), C.c_int64] + [C.c_void_p] * 13 + \
            [C.c_void_p, C.c_size_t, C.c_void_p]
        L.ops_beamsolve_launch.argtypes = [C.POINTER(OpsBeamOptParams), C.c_int64] + [C.c_void_p] * 10 + \
            [C.c_void_p]
        L.ops_beamopt_run_host.argtypes = [C.POINTER(OpsBeamOptParams), C.c_int64] + [C.c_void_p] * 12 + \
            [C.c_int, C.POINTER(C.c_float)]
        L.ops_fp64_peak_probe.argtypes = [C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_float), C.c_void_p]
        L.ops_fastmath_selftest.argtypes = [C.c_int64, C.POINTER(C.c_int64), C.POINTER(C.c_int64),
                                            C.POINTER(C.c_double), C.c_void_p]
        L.ops_pipe_probe.argtypes = [C.c_int, C.c_int, C.POINTER(C.c_double), C.c_void_p]
        L.ops_beamopt_session_create.argtypes = [C.POINTER(OpsBeamOptParams), C.c_int64, C.c_int,
                                                 C.POINTER(C.c_void_p)]
        L.ops_beamopt_session_arrays.argtypes = [C.c_void_p, C.POINTER(OpsBeamOptHostArrays)]
        L.ops_beamopt_session_run.argtypes = [C.c_void_p, C.c_int64, C.POINTER(C.c_float)]
        L.ops_beamopt_session_destroy.argtypes = [C.c_void_p]
        L.ops_beamopt_session_destroy.restype = None
        L.ops_beamopt_launch_scatter.argtypes = [C.POINTER(OpsBeamOptParams), C.c_int64] + [C.c_void_p] * 5 + \
            [C.c_int, C.POINTER(OpsBeamOptRecordArrays), C.c_int64, C.c_void_p, C.c_size_t, C.c_void_p]
        L.ops_beamopt_scatter_supported.argtypes = [C.POINTER(OpsBeamOptParams)]
        L.ops_beamopt_plan.argtypes = [C.POINTER(OpsBeamOptParams), C.c_int64, C.c_int32, C.c_int32,
                                       C.POINTER(OpsLaunchPlanInfo)]
        L.ops_peer_alloc.argtypes = [C.c_size_t, C.POINTER(C.c_void_p), C.c_char_p]
        L.ops_peer_open.argtypes = [C.c_char_p, C.POINTER(C.c_void_p)]
        L.ops_peer_close.argtypes = [C.c_void_p]
        L.ops_peer_free.argtypes = [C.c_void_p]
        L.ops_sampler_create.argtypes = [C.c_uint64, C.POINTER(C.c_void_p)]
        L.ops_sampler_destroy.argtypes = [C.c_void_p]
        L.ops_sampler_destroy.restype = None
        L.ops_sampler_random.argtypes = [C.c_void_p]
        L.ops_sampler_random.restype = C.c_double
        L.ops_sampler_randint.argtypes = [C.c_void_p, C.c_int, C.c_int]
        L.ops_sampler_draw_cases.argtypes = [C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_double, C.c_void_p, C.c_int32,
                                             C.c_void_p, C.c_int32, C.c_double, C.c_double, C.c_int32, C.c_int32,
                                             C.c_double, C.c_double, C.c_int32, C.c_int32] + [C.c_void_p] * 5 + \
            [C.c_int32, C.c_void_p, C.c_void_p]
        L.ops_frameopt_max_elements.argtypes = [C.POINTER(OpsFrameOptParams)]
        L.ops_frameopt_fill_schedule.argtypes = [C.POINTER(OpsFrameOptParams), C.c_void_p]
        L.ops_frameopt_launch.argtypes = [C.POINTER(OpsFrameOptParams), C.c_int64] + [C.c_void_p] * 11
        L.ops_frameopt_run_host.argtypes = [C.POINTER(OpsFrameOptParams), C.c_int64] + [C.c_void_p] * 9 + \
            [C.c_int, C.POINTER(C.c_float)]
        _lib = L
    return _lib


def check(rc: int, what: str):
    if rc == 0:
        return
    names = {-1: "OPS_E_BADARG", -2: "OPS_E_UNSUPP", -3: "OPS_E_WORKSPACE"}
    if rc < 0:
        raise CudaLibraryError(f"{what}: {names.get(rc, rc)}")
    raise CudaLibraryError(f"{what}: cudaError {rc}")


def version() -> str:
    return lib().ops_beamopt_version().decode()


def fill_schedule(p: BeamOptParams) -> np.ndarray:
    """[max_e, 2] fp32 table of (-(lr_t / bias_correction1), sqrt(bias_correction2))."""
    cp = to_c_params(p)
    table = np.zeros((max(p.max_e, 1), 2), np.float32)
    check(lib().ops_beamopt_fill_schedule(C.byref(cp), table.ctypes.data), "ops_beamopt_fill_schedule")
    return table


def run_host(p: BeamOptParams, fixed_uy, force_nodes, force_vals, L, device: int = 0) -> dict:
    """ops_beamopt_run_host: host numpy buffers in, host numpy buffers out (H2D/D2H inside)."""
    L = np.ascontiguousarray(L, np.float64).reshape(-1)
    B, nn, Cc, F = L.shape[0], p.num_nodes, p.num_cases, p.max_forces
    n = nn - 1
    fixed_uy = np.ascontiguousarray(fixed_uy, np.uint8).reshape(B, nn)
    force_nodes = np.ascontiguousarray(force_nodes, np.int32).reshape(B, Cc, F)
    force_vals = np.ascontiguousarray(force_vals, np.float64).reshape(B, Cc, F)
    out = {
        "I": np.empty((B, n), np.float32), "defl": np.empty((B, Cc, nn)), "rot": np.empty((B, Cc, nn)),
        "shear": np.empty((B, Cc, n), np.float32), "moment": np.empty((B, Cc, n), np.float32),
        "epochs": np.empty(B, np.int32), "loss": np.empty(B, np.float32), "status": np.empty(B, np.int32),
    }
    ms = C.c_float(0.0)
    cp = to_c_params(p)
    rc = lib().ops_beamopt_run_host(
        C.byref(cp), B, fixed_uy.ctypes.data, force_nodes.ctypes.data, force_vals.ctypes.data,
        L.ctypes.data, out["I"].ctypes.data, out["defl"].ctypes.data, out["rot"].ctypes.data,
        out["shear"].ctypes.data, out["moment"].ctypes.data, out["epochs"].ctypes.data,
        out["loss"].ctypes.data, out["status"].ctypes.data, device, C.byref(ms))
    check(rc, "ops_beamopt_run_host")
    out["kernel_ms"] = float(ms.value)
    return out


def launch_plan(p: BeamOptParams, B: int, sms: int = B200_SMS, smem_optin: int = B200_SMEM_OPTIN) -> dict:
    """Kernel family and launch geometry ops_beamopt_launch would use for B beams (ops_beamopt_plan: host arithmetic
    only -- no device is touched unless sms = smem_optin = 0 asks for the current device's figures)."""
    cp, info = to_c_params(p), OpsLaunchPlanInfo()
    check(lib().ops_beamopt_plan(C.byref(cp), int(B), int(sms), int(smem_optin), C.byref(info)), "ops_beamopt_plan")
    out = {name: getattr(info, name) for name, _ in OpsLaunchPlanInfo._fields_}
    out["family"] = PLAN_FAMILIES[info.family]
    return out


def fp64_peak_probe(iters: int = 1 << 16, stream: int = 0):
    """(TFLOP/s, ms) of a pure DFMA kernel on the current device -- the FP64 roofline denominator."""
    tf, ms = C.c_double(0.0), C.c_float(0.0)
    check(lib().ops_fp64_peak_probe(iters, C.byref(tf), C.byref(ms), stream), "ops_fp64_peak_probe")
    return float(tf.value), float(ms.value)


def fastmath_selftest(samples: int = 1 << 26, stream: int = 0) -> dict:
    """Bit-compare the kernel's branch-free fp32 div/sqrt with the IEEE operators on the device."""
    mism = (C.c_int64 * 3)()
    ran, worst = C.c_int64(0), C.c_double(0.0)
    check(lib().ops_fastmath_selftest(samples, mism, C.byref(ran), C.byref(worst), stream), "ops_fastmath_selftest")
    return {"div": int(mism[0]), "sqrt": int(mism[1]), "rcp": int(mism[2]), "samples": int(ran.value),
            "rcp64_max_rel_err": float(worst.value)}


class Session:
    """ops_beamopt_session_*: persistent device buffers + pinned host arrays for up to ``max_beams``
    beams per run.  ``inputs`` / ``outputs`` are numpy VIEWS of the pinned arrays: fill the inputs in
    place (or use ``load``), call ``run(B)``, read the first B rows of the outputs (valid until the
    next run)."""

    def __init__(self, p: BeamOptParams, max_beams: int, device: int = 0):
        self.p, self.max_beams, self.device = p, int(max_beams), device
        self._cp = to_c_params(p)
        self._h = C.c_void_p()
        check(lib().ops_beamopt_session_create(C.byref(self._cp), self.max_beams, device, C.byref(self._h)),
              "ops_beamopt_session_create")
        arr = OpsBeamOptHostArrays()
        check(lib().ops_beamopt_session_arrays(self._h, C.byref(arr)), "ops_beamopt_session_arrays")
        B, nn, Cc, F = self.max_beams, p.num_nodes, p.num_cases, p.max_forces
        n = nn - 1

        def view(ptr, shape, dtype):
            count = int(np.prod(shape))
            if count == 0:
                return np.zeros(shape, dtype)
            ctype = np.ctypeslib.as_ctypes_type(dtype)
            buf = (ctype * count).from_address(ptr)
            return np.ctypeslib.as_array(buf).reshape(shape)

        self.inputs = {
            "fixed_uy": view(arr.fixed_uy, (B, nn), np.uint8),
            "force_nodes": view(arr.force_nodes, (B, Cc, F), np.int32),
            "force_vals": view(arr.force_vals, (B, Cc, F), np.float64),
            "L": view(arr.L, (B,), np.float64),
        }
        self.outputs = {
            "I": view(arr.I_values, (B, n), np.float32),
            "defl": view(arr.deflections, (B, Cc, nn), np.float64),
            "rot": view(arr.rotations, (B, Cc, nn), np.float64),
            "shear": view(arr.shear, (B, Cc, n), np.float32),
            "moment": view(arr.moment, (B, Cc, n), np.float32),
            "epochs": view(arr.epochs, (B,), np.int32),
            "loss": view(arr.loss, (B,), np.float32),
            "status": view(arr.status, (B,), np.int32),
        }
        self.kernel_ms = 0.0

    def load(self, fixed_uy, force_nodes, force_vals, L) -> int:
        B = len(L)
        if B > self.max_beams:
            raise ValueError("batch larger than the session")
        self.inputs["fixed_uy"][:B] = fixed_uy
        self.inputs["force_nodes"][:B] = np.asarray(force_nodes).reshape(B, self.p.num_cases, self.p.max_forces)
        self.inputs["force_vals"][:B] = np.asarray(force_vals).reshape(B, self.p.num_cases, self.p.max_forces)
        self.inputs["L"][:B] = L
        return B

    def run(self, B: int) -> dict:
        ms = C.c_float(0.0)
        check(lib().ops_beamopt_session_run(self._h, int(B), C.byref(ms)), "ops_beamopt_session_run")
        self.kernel_ms = float(ms.value)
        return {k: v[:B] for k, v in self.outputs.items()}

    def close(self):
        if self._h:
            lib().ops_beamopt_session_destroy(self._h)
            self._h = C.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


PIPE_PROBE_OPS = ("DFMA", "FFMA", "FMUL", "FADD", "MUFU.RCP", "F2F f32<->f64", "IMAD", "LOP3", "FFMA imm")


def pipe_probe(iters: int = 1 << 14) -> dict:
    """Sustained warp instructions per clock per SM of each instruction class (diagnostic)."""
    out = {}
    for op, name in enumerate(PIPE_PROBE_OPS):
        r = C.c_double(0.0)
        check(lib().ops_pipe_probe(op, iters, C.byref(r), 0), "ops_pipe_probe")
        out[name] = float(r.value)
    return out
