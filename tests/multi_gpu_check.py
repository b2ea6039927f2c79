"""Two or more GPUs of one box (not part of the single-GPU pytest run): the in-kernel dataset gather
(PeerDataset, ops_beamopt_launch_scatter) must produce, on EVERY rank, exactly the dataset the NCCL
all_gather of the per-rank blocks produces.  Launch:
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 \
        tests/multi_gpu_check.py
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from openpystruct_b200 import sampling                                   # noqa: E402
from openpystruct_b200.distributed import (PeerDataset, init_from_env, optimise_beams_scattered,   # noqa: E402
                                            optimise_beams_sharded)
from openpystruct_b200.params import BeamOptParams                       # noqa: E402
from tests.helpers import seeded_cases                                   # noqa: E402


def main():
    rank, local, world = init_from_env("nccl")
    dev = torch.device("cuda", local)
    # early-stopped batches: the record copy inside the kernel; fixed epoch counts (max_e given): chunks of whole rounds +
    # the peer-copy kernel on a side stream -- one chunk, several chunks with a ragged last one, four load cases
    for script, count, cases_per_beam, max_e in (("MC", 3001, 1, None), ("SC", 777, 1, None), ("MC", 402, 4, None),
                                                 ("MC", 1203, 1, 40), ("MC", 148 * 40 * world * 2 + 333, 1, 6),
                                                 ("MC", 395, 4, 25)):
        p = BeamOptParams.for_script(script)
        if cases_per_beam > 1:
            p = p.replace(num_cases=cases_per_beam)
        if max_e is not None:
            p = p.replace(early_stop=False, max_e=max_e)
        cases = seeded_cases(p, count * cases_per_beam, seed=17)
        fixed, fn, fv, L = sampling.pack_cases(p.num_nodes, p.max_forces, cases, cases_per_beam)
        t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)   # noqa: E731
        inputs = {"fixed_uy": t(fixed), "force_nodes": t(fn), "force_vals": t(fv), "L": t(L)}
        want = optimise_beams_sharded(p, inputs)
        got, ds = optimise_beams_scattered(p, inputs)
        torch.cuda.synchronize()
        for k in want:
            assert torch.equal(want[k], got[k][:count]), (rank, script, k)
        # a second call reuses the mappings (rows are rewritten)
        got2, _ = optimise_beams_scattered(p, inputs, ds)
        torch.cuda.synchronize()
        for k in want:
            assert torch.equal(want[k], got2[k][:count]), (rank, script, k, "second call")
        ds.close()
        if rank == 0:
            print(f"multi_gpu_check ok: {script} x{cases_per_beam} {count} beams on {world} GPUs, "
                  f"{'early stop' if max_e is None else 'fixed epochs (pipelined peer copy)'}, "
                  f"epochs {int(want['epochs'].min())}..{int(want['epochs'].max())}", flush=True)
    # the host-level entry under torchrun: every rank draws the same seeded cases, the beams are sharded, every rank
    # gets the whole dataset -- equal to the single-GPU run of all beams on this rank
    from openpystruct_b200 import generator
    from openpystruct_b200.generator import GeneratorConfig
    cfg = GeneratorConfig.multi_core()
    a = generator.generate_columnar(cfg, 1500, seed=3, device=dev)
    rollers, avail = sampling.fixed_bridge(cfg.params.num_nodes, cfg.roller_nodes)
    import random
    rng = random.Random(3)
    cases = [sampling.sample_case(cfg.params.num_nodes, cfg.random_bridge, cfg.L_max, rollers, avail, L_max=cfg.L_max,
                                  L_min=cfg.L_min, N_rollers_max=cfg.N_rollers_max, M_forces_max=cfg.M_forces_max,
                                  max_force=cfg.max_force, min_force=cfg.min_force, rng=rng) for _ in range(1500)]
    single = generator.optimise_cases(cfg.params, cases, dev, distributed=False)
    multi = generator.optimise_cases(cfg.params, cases, dev)
    for k in single:
        assert np.array_equal(single[k], multi[k]), (rank, "optimise_cases", k)
    assert len(a["I_values"]) == int((single["status"] == 0).sum())
    if rank == 0:
        print(f"multi_gpu_check ok: generate_columnar / optimise_cases under torchrun == single GPU ({world} GPUs)", flush=True)
    # a configuration whose kernel has no scatter instance (banded LDL^T solver) with an EMPTY shard on the last rank
    # (one beam): every rank must take the NCCL fall-back together -- no hang, no uninitialised rows
    p1 = cfg.params.replace(solver=1, max_e=20)
    one = generator.optimise_cases(p1, cases[:1], dev)
    ref1 = generator.optimise_cases(p1, cases[:1], dev, distributed=False)
    for k in ref1:
        assert np.array_equal(one[k], ref1[k]), (rank, "empty shard fall-back", k)
    few = generator.optimise_cases(cfg.params.replace(max_e=20), cases[:world - 1], dev)      # scatter path, last rank empty
    ref2 = generator.optimise_cases(cfg.params.replace(max_e=20), cases[:world - 1], dev, distributed=False)
    for k in ref2:
        assert np.array_equal(few[k], ref2[k]), (rank, "empty shard scatter", k)
    # ranks that sampled DIFFERENT cases (an unseeded `random` per process) are an error on every rank, not a silently
    # corrupted dataset
    rng_r = random.Random(100 + rank)
    mine = [sampling.sample_case(cfg.params.num_nodes, 0, cfg.L_max, rollers, avail, rng=rng_r) for _ in range(8)]
    try:
        generator.optimise_cases(cfg.params.replace(max_e=5), mine, dev)
        raise AssertionError("differing cases were accepted")
    except RuntimeError as ex:
        assert "sampled different cases" in str(ex), ex
    if rank == 0:
        print(f"multi_gpu_check ok: empty shards, collective fall-back, differing-cases check ({world} GPUs)", flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
