"""Sample sharding over ranks (one process per GPU) and the final dataset gather.

Beams are independent, so the path shards with NO per-iteration communication: rank r optimises
the contiguous block ``[r*ceil(B/W), (r+1)*ceil(B/W))`` of the global sample order (the order of the
reference's serial loop, SingleCore:256-257, which the trainers' consecutive-record grouping relies
on, PINN:237-258) and one ``all_gather_into_tensor`` per output tensor (NCCL over NVLink on GPUs)
reassembles the dataset.  Results are independent of the world size bit for bit.

On one NVLink box the gather can be fused into the kernel instead (``PeerDataset``): every rank maps
every other rank's dataset arrays (CUDA IPC) and the production kernel writes each finished beam's
record straight into all of them, so the transfer overlaps the optimisation of the beams still
running and NCCL only carries a 4-byte all_reduce as the closing barrier.

The reference has no counterpart (its parallelism is joblib/loky processes, MultiCore:258).
"""
from __future__ import annotations

from typing import Callable, Dict, Optional, Tuple

import torch
import torch.distributed as dist

OUTPUT_NAMES = ("I", "defl", "rot", "shear", "moment", "epochs", "loss", "status")


def shard_bounds(num_beams: int, rank: int, world: int) -> Tuple[int, int, int]:
    """(start, stop, per_rank): contiguous equal blocks of ceil(B/W); the last ranks may be short/empty."""
    per = (num_beams + world - 1) // world
    start = min(rank * per, num_beams)
    stop = min(start + per, num_beams)
    return start, stop, per


def shard_inputs(inputs: Dict[str, torch.Tensor], rank: int, world: int) -> Tuple[Dict[str, torch.Tensor], int]:
    """Slice beam-major tensors to this rank's block, padded (by repeating the last beam, or beam 0
    for an empty block) to exactly ceil(B/W) beams so that every rank gathers equal-sized buffers."""
    B = next(iter(inputs.values())).shape[0]
    start, stop, per = shard_bounds(B, rank, world)
    out = {}
    for k, t in inputs.items():
        s = t[start:stop]
        if s.shape[0] < per:
            filler = (s[-1:] if s.shape[0] > 0 else t[:1]).expand(per - s.shape[0], *t.shape[1:])
            s = torch.cat([s, filler], dim=0)
        out[k] = s.contiguous()
    return out, stop - start


_gather_buffers: Dict[tuple, torch.Tensor] = {}


def gather_outputs(local: Dict[str, torch.Tensor], num_beams: int, group=None) -> Dict[str, torch.Tensor]:
    """The dataset gather: every rank's block of every output in ONE collective.  The per-rank tensors are
    packed into one byte buffer (256-byte aligned segments), gathered with a single
    ``all_gather_into_tensor`` into a reused [world, bytes] buffer and handed back as typed views,
    trimmed to ``num_beams``."""
    world = dist.get_world_size(group)
    names = list(local)
    dev = local[names[0]].device
    segs, off = [], 0
    for k in names:
        t = local[k]
        nbytes = t.numel() * t.element_size()
        segs.append((k, off, nbytes, t.dtype, tuple(t.shape)))
        off += (nbytes + 255) // 256 * 256
    total = max(off, 256)
    key = (str(dev), world, total, id(group))
    bufs = _gather_buffers.get(key)
    if bufs is None:
        bufs = (torch.empty(total, dtype=torch.uint8, device=dev),
                torch.empty((world, total), dtype=torch.uint8, device=dev))
        _gather_buffers.clear()                      # one live shape at a time
        _gather_buffers[key] = bufs
    send, recv = bufs
    for k, o, nbytes, _, _ in segs:
        if nbytes:
            send[o:o + nbytes].copy_(local[k].contiguous().view(-1).view(torch.uint8))
    if dist.get_backend(group) == "gloo":
        dist.all_gather(list(recv.unbind(0)), send, group=group)
    else:
        dist.all_gather_into_tensor(recv, send, group=group)
    out = {}
    for k, o, nbytes, dtype, shape in segs:
        per = shape[0]
        if nbytes == 0:
            out[k] = torch.empty((0,) + shape[1:], dtype=dtype, device=dev)
            continue
        blocks = recv[:, o:o + nbytes].view(dtype).reshape((world * per,) + shape[1:]) if world == 1 else \
            recv[:, o:o + nbytes].contiguous().view(dtype).reshape((world * per,) + shape[1:])
        out[k] = blocks[:num_beams]
    return out


def run_sharded(compute: Callable[[Dict[str, torch.Tensor]], Dict[str, torch.Tensor]],
                inputs: Dict[str, torch.Tensor], group=None, gather: bool = True
                ) -> Dict[str, torch.Tensor]:
    """inputs: the GLOBAL beam-major tensors (identical on every rank, e.g. sampled from one seeded
    stream).  Each rank computes its block with ``compute`` and the blocks are gathered."""
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    B = next(iter(inputs.values())).shape[0]
    mine, _valid = shard_inputs(inputs, rank, world)
    local = compute(mine)
    if not gather:
        return local
    return gather_outputs(local, B, group)


def optimise_beams_sharded(params, inputs: Dict[str, torch.Tensor], group=None, gather: bool = True):
    """Product path: ``ops.optimise_beams`` on this rank's GPU + NCCL gather."""
    from . import ops as _ops

    def compute(shard):
        return _ops.optimise_beams(params, shard["fixed_uy"], shard["force_nodes"], shard["force_vals"],
                                   shard["L"])
    return run_sharded(compute, inputs, group, gather)


class PeerUnavailable(RuntimeError):
    """CUDA IPC / peer access between the ranks' GPUs does not work here (raised on every rank alike)."""


class PeerDataset:
    """The whole dataset's record arrays on THIS GPU, writable by every rank's kernel over NVLink.

    ``PeerDataset(params, num_beams, group)`` is collective: each rank allocates one device buffer
    (``ops_peer_alloc``), the 64-byte IPC handles are exchanged through ``torch.distributed`` and every
    rank maps the others' buffers (``ops_peer_open``).  ``optimise(shard_inputs, row0)`` launches the
    production kernel on this rank's beams with all ``world`` buffers as destinations and closes with a
    one-element all_reduce on the stream: when it has run, ``tensors()`` holds the complete dataset on
    every rank.  The tensors alias the peer buffer -- consume or clone them before any rank's next call.
    """

    NAMES = OUTPUT_NAMES

    def __init__(self, params, num_beams: int, group=None, device: Optional[torch.device] = None):
        import ctypes as C
        from . import _cabi
        self._cabi, self._C = _cabi, C
        self.params, self.num_beams, self.group = params, int(num_beams), group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        if self.world > 8:
            raise ValueError("PeerDataset: at most 8 ranks (one NVLink box)")
        # Decided from the parameter block alone, BEFORE any collective: every rank (also one whose shard will be
        # empty) raises alike for a configuration whose kernel has no scatter instance, and the caller's fall-back to
        # the NCCL gather is taken by all of them.
        if not _cabi.lib().ops_beamopt_scatter_supported(C.byref(_cabi.to_c_params(params))):
            raise PeerUnavailable("this configuration runs a kernel without the in-kernel dataset gather "
                                  "(solver, num_nodes > 169 or multi-case beyond 105 nodes)")
        self.device = device if device is not None else torch.device("cuda", torch.cuda.current_device())
        nn, Cc = params.num_nodes, params.num_cases
        n, Bt = nn - 1, self.num_beams
        spec = (("I", (Bt, n), torch.float32), ("defl", (Bt, Cc, nn), torch.float64), ("rot", (Bt, Cc, nn), torch.float64),
                ("shear", (Bt, Cc, n), torch.float32), ("moment", (Bt, Cc, n), torch.float32),
                ("epochs", (Bt,), torch.int32), ("loss", (Bt,), torch.float32), ("status", (Bt,), torch.int32))
        self._layout, off = [], 0
        for name, shape, dtype in spec:
            nbytes = int(torch.empty((), dtype=dtype).element_size())
            for d in shape:
                nbytes *= d
            self._layout.append((name, off, shape, dtype))
            off += (nbytes + 255) // 256 * 256
        self.nbytes = max(off, 256)
        lib = _cabi.lib()
        # Collective set-up that cannot leave a rank behind: every rank takes part in every exchange whatever
        # happened locally, and the outcome is agreed on with an all_reduce (PeerUnavailable on ALL ranks).
        err, self._own, self._bases = None, 0, []
        with torch.cuda.device(self.device):
            handle = C.create_string_buffer(64)
            try:
                _cabi.check(lib.ops_set_device(self.device.index), "ops_set_device")
                base = C.c_void_p()
                _cabi.check(lib.ops_peer_alloc(self.nbytes, C.byref(base), handle), "ops_peer_alloc")
                self._own = int(base.value)
            except Exception as ex:                      # noqa: BLE001
                err = ex
            handles = [None] * self.world
            dist.all_gather_object(handles, bytes(handle.raw) if err is None else None, group=group)
            if err is None and any(h is None for h in handles):
                err = RuntimeError("a peer could not allocate its dataset buffer")
            for r, h in enumerate(handles):
                if r == self.rank or err is not None:
                    self._bases.append(self._own if r == self.rank else 0)
                    continue
                try:
                    ptr = C.c_void_p()
                    _cabi.check(lib.ops_peer_open(C.create_string_buffer(h, 64), C.byref(ptr)), f"ops_peer_open(rank {r})")
                    self._bases.append(int(ptr.value))
                except Exception as ex:                  # noqa: BLE001
                    err = ex
                    self._bases.append(0)
            ok = torch.tensor([0 if err is not None else 1], dtype=torch.int32, device=self.device)
            dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=group)
            if int(ok.item()) == 0:
                for r, b in enumerate(self._bases):
                    if b and r != self.rank:
                        lib.ops_peer_close(b)
                dist.barrier(group=group)
                if self._own:
                    lib.ops_peer_free(self._own)
                self._bases = None
                raise PeerUnavailable(f"peer-visible dataset buffers are not available on this box: {err}")
        # destination 0 = this GPU's own arrays (the kernel writes the record there and copies the rows to the
        # others); the peers follow in ring order so that the ranks do not all store to the same GPU at once
        self._dests = (_cabi.OpsBeamOptRecordArrays * self.world)()
        fields = ("I_values", "deflections", "rotations", "shear", "moment", "epochs", "loss", "status")
        for i in range(self.world):
            b = self._bases[(self.rank + i) % self.world]
            for f, (_, o, _, _) in zip(fields, self._layout):
                setattr(self._dests[i], f, b + o)
        self._flag = torch.zeros(1, dtype=torch.int32, device=self.device)
        self._ok = torch.ones(1, dtype=torch.int32, device=self.device)     # closing collective: MIN over the ranks' launch status
        self._tensors = {name: self._wrap(self._own + o, shape, dtype) for name, o, shape, dtype in self._layout}

    def _wrap(self, ptr: int, shape, dtype) -> torch.Tensor:
        typestr = {torch.float32: "<f4", torch.float64: "<f8", torch.int32: "<i4"}[dtype]

        class _Arr:
            __cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (ptr, False), "version": 3,
                                        "strides": None}
        if 0 in shape:
            return torch.empty(shape, dtype=dtype, device=self.device)
        with torch.cuda.device(self.device):
            return torch.as_tensor(_Arr(), device=self.device)

    def tensors(self) -> Dict[str, torch.Tensor]:
        return dict(self._tensors)

    def optimise(self, shard: Dict[str, torch.Tensor], row0: int, check: bool = True) -> Dict[str, torch.Tensor]:
        """Optimise this rank's beams (``shard``: fixed_uy, force_nodes, force_vals, L on this device) into rows
        ``row0 ...`` of every rank's dataset, then the stream barrier.  The closing collective carries every rank's launch
        status (MIN): a rank whose launch failed still takes part in it and ALL ranks raise (``check=True`` reads it
        back, one host synchronisation; a caller that pipelines steps passes ``check=False`` and calls ``check()``
        itself)."""
        from . import ops as _ops
        C, _cabi, p = self._C, self._cabi, self.params
        B = int(shard["L"].shape[0])
        if row0 < 0 or row0 + B > self.num_beams:
            raise ValueError("PeerDataset.optimise: rows outside the dataset")
        ip, fp = _ops.pack_params(p)
        cp = _ops._c_params(ip, fp)
        sched = _ops.device_schedule(p, self.device)
        _ops._check_cuda(shard["fixed_uy"], shard["force_nodes"], shard["force_vals"], shard["L"], sched)
        lib = _cabi.lib()
        rc = 0
        with torch.cuda.device(self.device):
            # no rank may still be inside the kernels of its previous call when rows are rewritten
            dist.all_reduce(self._flag, group=self.group)
            if B > 0:
                ws_bytes = lib.ops_beamopt_workspace_bytes(C.byref(cp), B)
                ws = torch.empty((max(int(ws_bytes), 1),), dtype=torch.uint8, device=self.device)
                stream = torch.cuda.current_stream(self.device)
                rc = lib.ops_beamopt_launch_scatter(
                    C.byref(cp), B, shard["fixed_uy"].data_ptr(), shard["force_nodes"].data_ptr(),
                    shard["force_vals"].data_ptr(), shard["L"].data_ptr(), sched.data_ptr(),
                    self.world, self._dests, int(row0), ws.data_ptr(), ws.numel(), stream.cuda_stream)
                ws.record_stream(stream)
            self._ok.fill_(1 if rc == 0 else 0)
            self._rc = rc
            dist.all_reduce(self._ok, op=dist.ReduceOp.MIN, group=self.group)   # every rank's kernel, hence every record, has landed
        if check:
            self.check()
        return self.tensors()

    def check(self):
        """Raises on EVERY rank when any rank's last launch failed (reads the closing collective back: host sync)."""
        if int(self._ok.item()) == 0:
            self._cabi.check(getattr(self, "_rc", 0), "ops_beamopt_launch_scatter")
            raise self._cabi.CudaLibraryError("ops_beamopt_launch_scatter failed on another rank of the job")

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def close(self):
        """Collective: unmap the peers' buffers, then free this rank's (every rank must call it)."""
        if getattr(self, "_bases", None) is None:
            return
        lib = self._cabi.lib()
        torch.cuda.synchronize(self.device)
        dist.barrier(group=self.group)
        with torch.cuda.device(self.device):
            for r, b in enumerate(self._bases):
                if r != self.rank:
                    lib.ops_peer_close(b)
            dist.barrier(group=self.group)
            lib.ops_peer_free(self._own)
        self._bases, self._tensors = None, {}


def optimise_beams_scattered(params, inputs: Dict[str, torch.Tensor], dataset: Optional[PeerDataset] = None,
                             group=None) -> Tuple[Dict[str, torch.Tensor], PeerDataset]:
    """``optimise_beams_sharded`` with the gather fused into the kernel.  ``inputs``: the GLOBAL beam-major
    tensors on this rank's device; returns (dataset tensors on this rank, the PeerDataset to reuse / close)."""
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    B = next(iter(inputs.values())).shape[0]
    start, stop, _per = shard_bounds(B, rank, world)
    shard = {k: t[start:stop].contiguous() for k, t in inputs.items()}
    if dataset is None:
        dataset = PeerDataset(params, B, group, shard["L"].device)
    return dataset.optimise(shard, start), dataset


def init_from_env(backend: Optional[str] = None) -> Tuple[int, int, int]:
    """torchrun environment -> (rank, local_rank, world).  No-op when WORLD_SIZE is unset or 1."""
    import os
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local)
            dist.init_process_group(backend, device_id=torch.device("cuda", local))
        else:
            dist.init_process_group(backend)
    return rank, local, world
