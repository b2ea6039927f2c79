"""Host-side logic: sampling order, ABI packing, record schema, sharding arithmetic."""
import json
import random

import numpy as np
import pytest
import torch

from openpystruct_b200 import generator, sampling
from openpystruct_b200.distributed import shard_bounds, shard_inputs
from openpystruct_b200.params import BeamOptParams
from tests.helpers import oracle_run, seeded_cases


def test_fixed_bridge_matches_reference_geometry():
    rollers, avail = sampling.fixed_bridge(101)
    assert rollers == [10, 30, 70, 85, 100]
    assert len(avail) == 94 and 1 not in avail and 101 not in avail and not set(rollers) & set(avail)


def test_sample_case_ranges_and_rng_call_order():
    rollers, avail = sampling.fixed_bridge(101)

    class Spy(random.Random):
        calls = []

        def randint(self, a, b):
            self.calls.append("randint"); return super().randint(a, b)

        def sample(self, pop, k):
            self.calls.append("sample"); return super().sample(pop, k)

        def uniform(self, a, b):
            self.calls.append("uniform"); return super().uniform(a, b)

        def choice(self, seq):
            self.calls.append("choice"); return super().choice(seq)

    rng = Spy(3)
    L, r, fnodes, fvals = sampling.sample_case(101, 0, 200.0, rollers, avail, rng=rng)
    assert rng.calls == ["randint", "sample"] + ["uniform"] * len(fnodes)
    assert L == 200.0 and r == rollers and 1 <= len(fnodes) <= 4
    assert all(-355857 <= v <= -35585.7 for v in fvals) and set(fnodes) <= set(avail)
    rng.calls.clear()
    L, r, fnodes, fvals = sampling.sample_case(101, 1, 200.0, rollers, avail, rng=rng)
    assert rng.calls[0] == "uniform" and rng.calls[1] == "randint"
    assert rng.calls[2:2 + len(r)] == ["choice"] * len(r)
    assert 15.0 <= L <= 215.0 and 1 <= len(r) <= 4 and not set(r) & set(fnodes)


def test_beamopt_rollers_respect_min_spacing():
    rng = random.Random(0)
    done = 0
    for _ in range(40):
        try:
            L, rollers, fnodes, fvals = sampling.sample_beamopt_case(rng=rng)
        except RuntimeError:        # a draw on which the reference's own rejection loop never ends
            continue
        done += 1
        assert len(rollers) == 5 and len(fnodes) == 5
        assert all(abs(a - b) >= 15 for i, a in enumerate(rollers) for b in rollers[:i])
        assert all(-355857 <= v <= -0.5 * 355857 for v in fvals)
    assert done >= 10


def test_pack_cases_layout():
    cases = [(200.0, [10], [5, 7], [-1.0, -2.0]), (50.0, [3], [9], [-3.0])]
    fixed, fn, fv, L = sampling.pack_cases(11, 4, cases)
    assert fixed.dtype == np.uint8 and fn.dtype == np.int32 and fv.dtype == np.float64
    assert fixed.shape == (2, 11) and fn.shape == (2, 1, 4)
    assert fixed[0].nonzero()[0].tolist() == [0, 9]
    assert fixed[1].nonzero()[0].tolist() == [0, 2]
    assert fn[0, 0].tolist() == [4, 6, -1, -1] and fv[0, 0].tolist() == [-1.0, -2.0, 0.0, 0.0]
    assert L.tolist() == [200.0, 50.0]
    with pytest.raises(ValueError):
        sampling.pack_cases(11, 1, cases)


def test_records_have_the_reference_schema_and_json_roundtrip(tmp_path):
    p = BeamOptParams.for_script("SC").replace(max_e=5)
    cases = seeded_cases(p, 3, seed=4)
    fixed, fn, fv, L = sampling.pack_cases(p.num_nodes, p.max_forces, cases)
    out = oracle_run(p, fixed, fn, fv, L)          # stands in for the GPU op: host logic under test
    out["status"][1] = 1
    recs = generator.make_records(p, cases, out)
    assert recs[1] is None
    assert tuple(recs[0].keys()) == generator.TRAINING_DATA_KEYS
    r = recs[0]
    assert len(r["I_values"]) == 100 and len(r["deflections"]) == 101 and len(r["node_positions"]) == 101
    assert r["roller_nodes"] == [10, 30, 70, 85, 100] and r["num_nodes"] == 101 and r["L"] == 200.0
    assert r["roller_x_locations"] == [18.0, 58.0, 138.0, 168.0, 198.0]
    data = {k: [rec[k] for rec in recs if rec is not None] for k in generator.TRAINING_DATA_KEYS}
    path = tmp_path / "training_data_PINN_mini.json"
    generator.save_training_data(data, str(path))
    back = json.load(open(path))
    assert list(back) == list(generator.TRAINING_DATA_KEYS)
    assert np.allclose(back["I_values"][0], r["I_values"], rtol=0, atol=0)
    # what the trainers do with it (PINN:226-258): pad to float32 and group consecutive records
    arr = np.array(back["deflections"], dtype=np.float32)
    assert arr.shape == (2, 101)


def test_script_presets():
    assert BeamOptParams.for_script("SC").patience == 5
    mc = BeamOptParams.for_script("MC")
    assert mc.patience == 10 and mc.zero_last_node and mc.tolerance == 5e-3
    g = BeamOptParams.for_script("GPU")
    assert g.patience == 100 and g.tolerance == 1e-2
    bo = BeamOptParams.for_script("BO")
    assert bo.max_e == 1000 and bo.uniform_udl == -5000.0 and bo.max_forces == 5
    assert BeamOptParams().G == pytest.approx(200e9 / 2.6)


@pytest.mark.parametrize("B,W", [(10, 1), (10, 2), (10, 3), (10, 4), (7, 8), (1, 8), (1000000, 8)])
def test_shard_bounds_cover_everything_in_order(B, W):
    seen = []
    for r in range(W):
        a, b, per = shard_bounds(B, r, W)
        assert per == -(-B // W) and 0 <= a <= b <= B
        seen.extend(range(a, b) if B < 1000 else [a, b])
    if B < 1000:
        assert seen == list(range(B))


def test_shard_inputs_pads_to_equal_blocks():
    x = {"a": torch.arange(10).reshape(10, 1), "b": torch.arange(20).reshape(10, 2)}
    blocks = [shard_inputs(x, r, 4) for r in range(4)]
    assert [v for _, v in blocks] == [3, 3, 3, 1]
    assert all(s["a"].shape == (3, 1) and s["b"].shape == (3, 2) for s, _ in blocks)
    assert blocks[3][0]["a"].flatten().tolist() == [9, 9, 9]
    empty = shard_inputs({"a": torch.arange(2).reshape(2, 1)}, 3, 4)
    assert empty[1] == 0 and empty[0]["a"].shape == (1, 1)


def test_native_sampler_is_cpythons_random_bit_for_bit():
    """csrc/sampler_host.cpp: MT19937 seeded like random.seed(int), randint / sample / choice / uniform in the
    reference's call order (SingleCore:133-160).  Against `random` itself: raw streams (random(), randint over many
    ranges), then > 10^5 sampled beams for both bridge modes, several seeds (incl. one beyond 32 bits), single and
    multi-case packing -- the ABI arrays, the record metadata, and the stream position afterwards must be identical."""
    import random
    s = sampling.NativeSampler(12345)
    r = random.Random(12345)
    for i in range(200000):
        if i % 3 == 0:
            assert s.random() == r.random()
        else:
            hi = 1 + (i * 7919) % 5000
            assert s.randint(1, hi) == r.randint(1, hi)
    rollers, avail = sampling.fixed_bridge(101)
    for flag, seed, nc, count in ((0, 0, 1, 60000), (1, 7, 1, 40000), (0, 2 ** 40 + 5, 4, 8000), (1, 99, 2, 8000)):
        rng = random.Random(seed)
        cases = [sampling.sample_case(101, flag, 200.0, rollers, avail, rng=rng) for _ in range(count)]
        want = sampling.pack_cases(101, 4, cases, nc)
        nat = sampling.NativeSampler(seed)
        pc = nat.draw_cases(count, 101, flag, 200.0, rollers, avail, num_cases=nc)
        for a, b in zip(want, pc.abi_arrays()):
            assert np.array_equal(a, b), (flag, seed, nc)
        assert pc.cases()[:500] == [(float(c[0]), list(c[1]), list(c[2]), list(c[3])) for c in cases[:500]]
        assert rng.random() == nat.random()
    # the reference goldens' draws (the reference's own statements consumed the stream of random.seed(seed))
    from tests.helpers import goldens
    for m, _ in goldens():
        if m["script"] == "BO":
            continue
        pc = sampling.NativeSampler(m["seed"]).draw_cases(1, 101, m["flag"], 200.0, rollers, avail)
        L, rl, ft, fvv = pc.cases()[0]
        assert (rl, ft, fvv) == (m["roller_nodes"], m["force_nodes"], m["force_values"]) and L == m["L"]
