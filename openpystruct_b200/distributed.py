"""Sample sharding over ranks (one process per GPU) and the final dataset gather.

Beams are independent, so the path shards with NO per-iteration communication: rank r optimises
the contiguous block ``[r*ceil(B/W), (r+1)*ceil(B/W))`` of the global sample order (the order of the
reference's serial loop, SingleCore:256-257, which the trainers' consecutive-record grouping relies
on, PINN:237-258) and one ``all_gather_into_tensor`` per output tensor (NCCL over NVLink on GPUs)
reassembles the dataset.  Results are independent of the world size bit for bit.

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
