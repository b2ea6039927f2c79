"""Host mirror of the reference generators' ``generate_sample`` / ``main``.

Same entry-point name, argument meaning and record schema as the reference
(``generate_sample(sample_idx, num_nodes, flag, L, node_positions, roller_nodes, available_nodes,
patience[, device]) -> dict | None``, SingleCore:126-249 / MultiCore:130-240 / GPU:131-274; the
13-key record SingleCore:235-249; dict-of-lists ``training_data`` + ``json.dump`` SingleCore:73-87,
263-264), but the inner epoch loop runs as ONE CUDA launch over a whole batch of samples.

What stays on the host, unchanged in behaviour: the sampling of supports and loads (``sampling``),
the packing of the record, the JSON dump.  What moves to the GPU: everything between
``I_tensor = torch.tensor([I_0] * num_elements ...)`` and the ``nodeDisp`` read-out.
"""
from __future__ import annotations

import dataclasses
import json
import random
from typing import List, Optional, Sequence

import numpy as np
import torch

from . import ops as _ops
from ._cabi import CudaLibraryError
from . import sampling
from .params import BeamOptParams

TRAINING_DATA_KEYS = (
    "roller_x_locations", "force_x_locations", "force_values", "I_values", "shear_forces",
    "bending_moments", "node_positions", "roller_nodes", "force_nodes", "num_nodes", "L",
    "rotations", "deflections",
)


@dataclasses.dataclass(frozen=True)
class GeneratorConfig:
    """Module-level constants of a generator script (SingleCore:20-49) that are not solver params."""
    params: BeamOptParams = BeamOptParams()
    L_max: float = 200.0
    L_min: float = 15.0
    N_rollers_max: int = 4
    M_forces_max: int = 4
    max_force: float = -355857
    num_samples: int = 100000
    random_bridge: int = 0           # flag
    roller_nodes: Optional[tuple] = None    # default [10, 30, 70, 85, num_nodes - 1]
    output_file: str = "training_data_PINN_mini.json"

    @property
    def min_force(self) -> float:
        return self.max_force / 10

    @staticmethod
    def single_core() -> "GeneratorConfig":
        return GeneratorConfig(params=BeamOptParams.for_script("SC"))

    @staticmethod
    def multi_core() -> "GeneratorConfig":
        return GeneratorConfig(params=BeamOptParams.for_script("MC"))

    @staticmethod
    def gpu() -> "GeneratorConfig":
        return GeneratorConfig(params=BeamOptParams.for_script("GPU"))


def _require_cuda(device) -> torch.device:
    dev = torch.device(device)
    if dev.type != "cuda" or not torch.cuda.is_available():
        raise RuntimeError("openpystruct_b200 runs the optimisation loop on a CUDA device only "
                           "(no CPU fallback); got device=%r, cuda available=%s"
                           % (device, torch.cuda.is_available()))
    if dev.index is None:
        dev = torch.device("cuda", torch.cuda.current_device())
    return dev


def optimise_cases(params: BeamOptParams, cases: Sequence[sampling.Case], device="cuda",
                   distributed: Optional[bool] = None, copy: bool = True) -> dict:
    """Pack sampled cases, copy them to the GPU, run the fused loop, bring the results back (numpy).

    Under ``torchrun`` (an initialised process group with more than one rank; ``distributed=False`` opts out)
    every rank must call this with the SAME cases -- they come from one seeded stream -- and the beams are
    sharded over the ranks' GPUs: contiguous blocks, no per-iteration communication, the dataset gather fused
    into the kernel's record write where the GPUs can map each other's memory, the NCCL gather otherwise
    (``distributed.py``).  Every rank gets the whole dataset back, bit-identical to the single-GPU run."""
    dev = _require_cuda(device)
    if isinstance(cases, sampling.PackedCases):
        fixed, fn, fv, L = cases.abi_arrays()
    else:
        fixed, fn, fv, L = sampling.pack_cases(params.num_nodes, params.max_forces, cases, params.num_cases)
    t = lambda a: torch.from_numpy(a).pin_memory().to(dev, non_blocking=True)   # noqa: E731
    import torch.distributed as dist
    if distributed is None:
        distributed = dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1
    if distributed:
        from . import distributed as _dist
        _require_identical_cases(fixed, fn, fv, L, dev)
        inputs = {"fixed_uy": t(fixed), "force_nodes": t(fn), "force_vals": t(fv), "L": t(L)}
        host = None
        try:
            ds = _dist.PeerDataset(params, int(L.shape[0]), device=dev)
        except _dist.PeerUnavailable:                          # raised on every rank alike, before or inside the set-up
            ds = None
        if ds is not None:
            try:
                start, stop, _per = _dist.shard_bounds(int(L.shape[0]), dist.get_rank(), dist.get_world_size())
                out = ds.optimise({k: v[start:stop].contiguous() for k, v in inputs.items()}, start)
                host = {k: v.cpu() for k, v in out.items()}      # (copies: the tensors alias the peer buffer)
            finally:
                ds.close()
        if host is None:
            host = {k: v.cpu() for k, v in _dist.optimise_beams_sharded(params, inputs).items()}
        out = {k: v.numpy() for k, v in host.items()}
    else:
        out = _session_run(params, fixed, fn, fv, L, dev, copy)
    return _rerun_unsupported(params, out, fixed, fn, fv, L, dev)


_sessions = {}


def _session_run(params: BeamOptParams, fixed, fn, fv, L, dev, copy: bool = True) -> dict:
    """Single-GPU path through the host-buffer SESSION of the C ABI (pinned staging buffers, the record of each chunk
    copied back under the next chunk's iterations): one cached session per (parameters, device), grown on demand.
    ``copy=False`` hands out VIEWS of the session's pinned arrays (valid until the next run on that session)."""
    from . import _cabi
    B = int(L.shape[0])
    key = (params, dev.index)
    sess = _sessions.get(key)
    if sess is None or sess.max_beams < B:
        if sess is not None:
            sess.close()
        if len(_sessions) >= 4:                                      # a few live parameter sets at most
            _sessions.pop(next(iter(_sessions))).close()
        sess = _sessions[key] = _cabi.Session(params, max(B, 1024), device=dev.index)
    sess.load(fixed, fn, fv, L)
    out = sess.run(B)
    return {k: v.copy() for k, v in out.items()} if copy else out      # (the pinned arrays are reused by the next run)


def _require_identical_cases(fixed, fn, fv, L, dev) -> None:
    """Under torchrun every rank samples the cases itself and only optimises its block of them; the records pair the
    gathered results with the LOCAL cases, so ranks that drew different cases (an unseeded ``random`` per process)
    would silently corrupt the dataset.  One 16-byte all_reduce of a digest of the packed inputs makes that an error."""
    import hashlib
    import torch.distributed as dist
    h = hashlib.blake2b(digest_size=8)
    for a in (fixed, fn, fv, L):
        h.update(np.ascontiguousarray(a).tobytes())
    v = int.from_bytes(h.digest(), "little") >> 1                    # 63 bits: fits int64
    lo_hi = torch.tensor([v, -v], dtype=torch.int64, device=dev)
    dist.all_reduce(lo_hi, op=dist.ReduceOp.MAX)                     # max(v) and max(-v) = -min(v)
    if int(lo_hi[0].item()) != v or int(lo_hi[1].item()) != -v:
        raise RuntimeError("openpystruct_b200: the ranks of this job sampled different cases -- under torchrun pass "
                           "the same `seed` (or an identically seeded `rng`) on every rank")


def _rerun_unsupported(params: BeamOptParams, out: dict, fixed, fn, fv, L, dev) -> dict:
    """Beams the three-moment kernels cannot take (status 3: more than 5 rollers -- the reference's ``ops.fix`` loop,
    SingleCore:101-102, takes any number) are re-run with the banded LDL^T solver behind the same ABI, on this GPU
    (under torchrun every rank does so for the few beams concerned: same inputs, same bits)."""
    from ._cabi import to_c_params  # noqa: F401  (solver constants live in the header)
    todo = np.flatnonzero(out["status"] == 3)
    if todo.size == 0 or params.solver == 1 or params.num_cases != 1:
        return out
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)   # noqa: E731
    again = _ops.optimise_beams(params.replace(solver=1), t(fixed[todo]), t(fn[todo]), t(fv[todo]), t(L[todo]))
    out = {k: np.array(v, copy=True) for k, v in out.items()}
    for k, v in again.items():
        out[k][todo] = v.cpu().numpy()
    return out


def make_records(params: BeamOptParams, cases: Sequence[sampling.Case], out: dict) -> List[Optional[dict]]:
    """The 13-key records of SingleCore:235-249, one per (beam, load case); None where status != 0
    (the MultiCore filter, MultiCore:184-186, 265)."""
    C = params.num_cases
    records: List[Optional[dict]] = []
    for i, (L, rollers, force_nodes, force_values) in enumerate(cases):
        b, c = divmod(i, C)
        if out["status"][b] != 0:
            records.append(None)
            continue
        rollers_b = cases[b * C][1]
        L_b = cases[b * C][0]
        node_positions = np.linspace(0, L_b, params.num_nodes)
        records.append({
            "roller_x_locations": [node_positions[t - 1] for t in rollers_b],
            "force_x_locations": [node_positions[t - 1] for t in force_nodes],
            "force_values": list(force_values),
            "I_values": out["I"][b].tolist(),
            "shear_forces": out["shear"][b, c].tolist(),
            "bending_moments": out["moment"][b, c].tolist(),
            "node_positions": node_positions.tolist(),
            "roller_nodes": list(rollers_b),
            "force_nodes": list(force_nodes),
            "num_nodes": params.num_nodes,
            "L": L_b,
            "rotations": out["rot"][b, c].tolist(),
            "deflections": out["defl"][b, c].tolist(),
        })
    return records


def generate_samples_batched(sample_indices: Sequence[int], num_nodes: int, flag: int, L: float,
                             node_positions, roller_nodes: Sequence[int], available_nodes: Sequence[int],
                             patience: int = 10, device="cuda", *, params: Optional[BeamOptParams] = None,
                             config: Optional[GeneratorConfig] = None, seed: Optional[int] = None,
                             rng=None) -> List[Optional[dict]]:
    """Batched ``generate_sample``: draws ``len(sample_indices)`` samples from ONE stream in index
    order (exactly what the reference's serial loop SingleCore:256-257 would draw after
    ``random.seed(seed)``), optimises them in one launch, returns the records in that order."""
    cfg = config or GeneratorConfig()
    p = (params or cfg.params).replace(num_nodes=num_nodes, patience=patience)
    if rng is None:
        rng = random if seed is None else random.Random(seed)
    elif seed is not None:
        raise ValueError("pass either seed or rng")
    # (under torchrun the ranks must draw the same cases: optimise_cases checks a digest of them collectively)
    if seed is not None and rng is random:
        random.seed(seed)
    cases = [sampling.sample_case(num_nodes, flag, L, roller_nodes, available_nodes, L_max=cfg.L_max,
                                  L_min=cfg.L_min, N_rollers_max=cfg.N_rollers_max,
                                  M_forces_max=cfg.M_forces_max, max_force=cfg.max_force,
                                  min_force=cfg.min_force, rng=rng)
             for _ in range(len(sample_indices) * p.num_cases)]
    out = optimise_cases(p, cases, device)
    return make_records(p, cases, out)


def generate_sample(sample_idx: int, num_nodes: int, flag: int, L: float, node_positions,
                    roller_nodes: Sequence[int], available_nodes: Sequence[int], patience: int = 10,
                    device="cuda", **kw) -> Optional[dict]:
    """Drop-in for the reference's per-sample function (one beam per launch; use the batched form
    for throughput).  Draws from the global ``random`` module like the reference."""
    return generate_samples_batched([sample_idx], num_nodes, flag, L, node_positions, roller_nodes,
                                    available_nodes, patience, device, **kw)[0]


def generate_dataset(config: Optional[GeneratorConfig] = None, num_samples: Optional[int] = None,
                     seed: Optional[int] = 0, device="cuda", batch_size: int = 65536) -> dict:
    """``main()`` of the generators (SingleCore:251-269): dict-of-lists over all samples, failed
    samples dropped (MultiCore:265).  Samples are drawn from one seeded stream in global order."""
    cfg = config or GeneratorConfig()
    p = cfg.params
    N = cfg.num_samples if num_samples is None else num_samples
    rollers, available = sampling.fixed_bridge(p.num_nodes, cfg.roller_nodes)
    node_positions = np.linspace(0, cfg.L_max, p.num_nodes)
    rng = random.Random(seed) if seed is not None else random
    training_data = {k: [] for k in TRAINING_DATA_KEYS}
    for start in range(0, N, batch_size):
        idx = range(start, min(start + batch_size, N))
        for rec in generate_samples_batched(idx, p.num_nodes, cfg.random_bridge, cfg.L_max, node_positions,
                                            rollers, available, p.patience, device, params=p, config=cfg,
                                            rng=rng):
            if rec is None:
                continue
            for k in TRAINING_DATA_KEYS:
                training_data[k].append(rec[k])
    return training_data


def sample_cases_packed(cfg: "GeneratorConfig", count: int, seed: int) -> "sampling.PackedCases":
    """``count`` generate_sample() draws after ``random.seed(seed)``, natively (csrc/sampler_host.cpp: a bit-exact
    replica of CPython's ``random`` in the reference's call order) and straight into the ABI arrays."""
    p = cfg.params
    rollers, available = sampling.fixed_bridge(p.num_nodes, cfg.roller_nodes)
    s = sampling.NativeSampler(seed)
    try:
        return s.draw_cases(count, p.num_nodes, cfg.random_bridge, cfg.L_max, rollers, available, L_max=cfg.L_max,
                            L_min=cfg.L_min, N_rollers_max=cfg.N_rollers_max, M_forces_max=cfg.M_forces_max,
                            max_force=cfg.max_force, min_force=cfg.min_force, num_cases=p.num_cases,
                            max_forces=p.max_forces)
    finally:
        s.close()


def generate_columnar(config: Optional[GeneratorConfig] = None, num_samples: Optional[int] = None,
                      seed: Optional[int] = 0, device="cuda", native_sampler: bool = True) -> dict:
    """``generate_dataset`` without the per-record Python objects: one array per key of the record
    (``dataset.columnar_from_run``), ready for ``dataset.save_npz`` / ``save_json``.  One launch.  With an integer seed
    the cases are drawn by the native sampler (same stream, same dataset, ~70x faster than the Python loop)."""
    from . import dataset as _dataset
    cfg = config or GeneratorConfig()
    p = cfg.params
    N = cfg.num_samples if num_samples is None else num_samples
    if native_sampler and isinstance(seed, int) and 0 <= seed < 2 ** 64:
        cases = sample_cases_packed(cfg, N * p.num_cases, seed)
    else:
        rollers, available = sampling.fixed_bridge(p.num_nodes, cfg.roller_nodes)
        rng = random.Random(seed) if seed is not None else random
        cases = [sampling.sample_case(p.num_nodes, cfg.random_bridge, cfg.L_max, rollers, available, L_max=cfg.L_max,
                                      L_min=cfg.L_min, N_rollers_max=cfg.N_rollers_max, M_forces_max=cfg.M_forces_max,
                                      max_force=cfg.max_force, min_force=cfg.min_force, rng=rng)
                 for _ in range(N * p.num_cases)]
    if isinstance(cases, sampling.PackedCases):
        from concurrent.futures import ThreadPoolExecutor
        with ThreadPoolExecutor(1) as pool:                    # the record keys that need no kernel output, beside the GPU run
            ahead = pool.submit(_dataset.case_columns, p, cases)
            out = optimise_cases(p, cases, device)
            return _dataset.columnar_from_run(p, cases, out, ahead.result())
    out = optimise_cases(p, cases, device)
    return _dataset.columnar_from_run(p, cases, out)


def stream_columnar(config: Optional[GeneratorConfig] = None, num_samples: Optional[int] = None, batch_size: int = 10000,
                    seed: int = 0, device="cuda", reuse_buffers: bool = False):
    """``main()`` of the generators as a PIPELINE (MultiCore:246-262 loops over batches too): yields the columnar record
    arrays of consecutive batches of one seeded stream; the cases of batch i + 1 are drawn (native sampler, its own
    thread -- ctypes releases the GIL -- together with the record keys that depend on the cases alone) while the GPU
    optimises batch i.  The concatenation of the batches is
    ``generate_columnar(config, num_samples, seed)``.  ``reuse_buffers=True``: the record arrays alias the session's pinned
    host buffers and are valid until the next batch is requested (write them out / consume them first)."""
    from concurrent.futures import ThreadPoolExecutor
    from . import dataset as _dataset
    cfg = config or GeneratorConfig()
    p = cfg.params
    N = cfg.num_samples if num_samples is None else num_samples
    rollers, available = sampling.fixed_bridge(p.num_nodes, cfg.roller_nodes)
    sampler = sampling.NativeSampler(seed)

    def draw(count):
        cases = sampler.draw_cases(count * p.num_cases, p.num_nodes, cfg.random_bridge, cfg.L_max, rollers, available,
                                   L_max=cfg.L_max, L_min=cfg.L_min, N_rollers_max=cfg.N_rollers_max,
                                   M_forces_max=cfg.M_forces_max, max_force=cfg.max_force, min_force=cfg.min_force,
                                   num_cases=p.num_cases, max_forces=p.max_forces)
        return cases, _dataset.case_columns(p, cases)              # (the record keys that need no kernel output)

    sizes = [min(batch_size, N - s) for s in range(0, N, batch_size)]
    with ThreadPoolExecutor(1) as pool:
        try:
            nxt = pool.submit(draw, sizes[0]) if sizes else None
            for i in range(len(sizes)):
                cases, case_cols = nxt.result()
                nxt = pool.submit(draw, sizes[i + 1]) if i + 1 < len(sizes) else None
                out = optimise_cases(p, cases, device, copy=not reuse_buffers)
                yield _dataset.columnar_from_run(p, cases, out, case_cols)
        finally:
            if nxt is not None:
                nxt.cancel()
            pool.shutdown(wait=True)
            sampler.close()


def save_training_data(training_data: dict, path: str = "training_data_PINN_mini.json") -> None:
    """json.dump of the dict-of-lists, the file the trainers json.load (SingleCore:263-264, PINN:192-206)."""
    def default(o):
        if isinstance(o, (np.floating, np.integer)):
            return o.item()
        raise TypeError(type(o))
    with open(path, "w") as f:
        json.dump(training_data, f, default=default)
