"""The N>1 path on CPU: world_size 2 over gloo.  The sharding/gather plumbing is the product's
(openpystruct_b200.distributed); the per-rank compute is the CPU oracle standing in for the CUDA op."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from openpystruct_b200 import sampling
from openpystruct_b200.distributed import run_sharded
from openpystruct_b200.params import BeamOptParams
from tests.helpers import oracle_run, seeded_cases


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _inputs(B):
    p = BeamOptParams.for_script("MC").replace(max_e=30)
    cases = seeded_cases(p, B, seed=21)
    fixed, fn, fv, L = sampling.pack_cases(p.num_nodes, p.max_forces, cases)
    return p, {"fixed_uy": torch.from_numpy(fixed), "force_nodes": torch.from_numpy(fn),
               "force_vals": torch.from_numpy(fv), "L": torch.from_numpy(L)}


def _compute(p):
    def f(shard):
        out = oracle_run(p, shard["fixed_uy"].numpy(), shard["force_nodes"].numpy(),
                         shard["force_vals"].numpy(), shard["L"].numpy())
        return {k: torch.from_numpy(v) for k, v in out.items()}
    return f


def _worker(rank, world, port, B, path):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        p, inputs = _inputs(B)
        out = run_sharded(_compute(p), inputs)
        if rank == 0:
            torch.save(out, path)
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("B", [7, 8])
def test_two_ranks_equal_one_rank_bit_for_bit(tmp_path, B):
    path = str(tmp_path / "out.pt")
    mp.spawn(_worker, args=(2, _free_port(), B, path), nprocs=2, join=True)
    got = torch.load(path)
    p, inputs = _inputs(B)
    want = _compute(p)(inputs)
    assert set(got) == set(want)
    for k in want:
        assert got[k].shape == want[k].shape, k
        assert torch.equal(got[k], want[k]), k
