"""Dataset writer / loader adapter and the trainers' pre-processing against the
reference's own formulas restated with numpy + sklearn (PINN:66-92, 226-368).  CPU only: the kernel
outputs are stood in for by the CPU oracle (tests only)."""
import json

import numpy as np
import torch
from sklearn.preprocessing import StandardScaler

from openpystruct_b200 import dataset, generator, sampling
from openpystruct_b200.params import BeamOptParams
from tests.helpers import oracle_run, seeded_cases


def _run(num_cases=1, count=40, flag=0, max_e=30):
    p = BeamOptParams.for_script("MC").replace(num_cases=num_cases, max_e=max_e)
    cases = seeded_cases(p, count * num_cases, seed=5, flag=flag)
    fixed, fn, fv, L = sampling.pack_cases(p.num_nodes, p.max_forces, cases, num_cases)
    return p, cases, oracle_run(p, fixed, fn, fv, L)


def test_columnar_equals_the_per_record_dicts(tmp_path):
    for num_cases, flag in ((1, 0), (1, 1), (4, 0)):
        p, cases, out = _run(num_cases, flag=flag)
        out["status"][3] = 1                                     # a failed beam is dropped (MultiCore:265)
        recs = [r for r in generator.make_records(p, cases, out) if r is not None]
        want = {k: [r[k] for r in recs] for k in generator.TRAINING_DATA_KEYS}
        col = dataset.columnar_from_run(p, cases, out)
        got = dataset.to_training_data(col)
        assert list(got) == list(generator.TRAINING_DATA_KEYS)
        for k in want:
            assert json.loads(json.dumps(got[k])) == json.loads(json.dumps(want[k], default=float)), k
        dataset.save_json(col, tmp_path / "d.json")
        dataset.save_npz(col, tmp_path / "d.npz")
        a, b = dataset.load_training_data(tmp_path / "d.json"), dataset.load_training_data(str(tmp_path / "d.npz"))
        assert a == b == got


def test_case_columns_computed_ahead_give_the_same_dataset():
    """stream_columnar computes the record keys that depend on the sampled cases alone on the sampler's thread
    (dataset.case_columns); the result must be the one-step columnar_from_run, and a batch with a failed beam must fall
    back to the filtering path."""
    rollers, avail = sampling.fixed_bridge(101)
    for num_cases, flag in ((1, 0), (1, 1), (4, 0)):
        p = BeamOptParams.for_script("MC").replace(num_cases=num_cases, max_e=12)
        pc = sampling.NativeSampler(17).draw_cases(24 * num_cases, 101, flag, 200.0, rollers, avail, num_cases=num_cases)
        out = oracle_run(p, *pc.abi_arrays())
        ahead = dataset.case_columns(p, pc)
        for drop in (False, True):
            o = {k: np.array(v) for k, v in out.items()}
            if drop:
                o["status"][5] = 1
            want = dataset.columnar_from_run(p, pc, o)
            got = dataset.columnar_from_run(p, pc, o, ahead)
            assert list(want) == list(got)
            for k in want:
                assert np.array_equal(want[k], got[k], equal_nan=True), (num_cases, flag, drop, k)
            assert len(got["L"]) == (24 - drop) * num_cases
        # and the packed cases produce the records of the per-case tuples
        listed = dataset.columnar_from_run(p, pc.cases(), out)
        for k in listed:
            assert np.array_equal(listed[k], dataset.columnar_from_run(p, pc, out, ahead)[k], equal_nan=True), k


def _reference_preprocess(data, n_cases, c, train_split, seed):
    """The trainers' block, restated with the reference's own tools (numpy + sklearn)."""
    def pad(rows):
        width = max(len(r) for r in rows)
        out = np.full((len(rows), width), 0.0, np.float32)
        for i, r in enumerate(rows):
            a = np.array(r, dtype=np.float32)
            out[i, :len(a)] = a
        return out
    total = len(data["I_values"]) // n_cases
    g = {k: pad(data[k])[:total * n_cases].reshape(total, n_cases, -1) for k in
         ("roller_x_locations", "force_x_locations", "force_values", "node_positions", "I_values", "deflections",
          "rotations")}
    idx = np.random.RandomState(seed).permutation(total)
    n_tr = int(train_split * total)
    tr, va = idx[:n_tr], idx[n_tr:]
    xs_tr, xs_va = [], []
    for k in ("roller_x_locations", "force_x_locations", "force_values", "node_positions"):
        s = StandardScaler()
        a = g[k][tr]
        xs_tr.append(s.fit_transform(a.reshape(-1, a.shape[-1])).reshape(a.shape))
        b = g[k][va]
        xs_va.append(s.transform(b.reshape(-1, b.shape[-1])).reshape(b.shape))
    unify = lambda x: x.mean(axis=1) + c * x.std(axis=1)        # noqa: E731
    ys_tr, ys_va = [], []
    for k in ("I_values", "deflections", "rotations"):
        s = StandardScaler().fit(unify(g[k][tr]))
        ys_tr.append(s.transform(unify(g[k][tr])))
        ys_va.append(s.transform(unify(g[k][va])))
    X_tr = np.concatenate(xs_tr, axis=2)
    X_va = np.concatenate(xs_va, axis=2)
    return (X_tr.reshape(X_tr.shape[0], -1), X_va.reshape(X_va.shape[0], -1), np.concatenate(ys_tr, axis=1),
            np.concatenate(ys_va, axis=1))


def _reference_labels(data, n_cases, c, train_split, seed):
    total = len(data["I_values"]) // n_cases
    idx = np.random.RandomState(seed).permutation(total)[:int(train_split * total)]
    cols = []
    for k in ("I_values", "deflections", "rotations"):
        a = np.asarray(data[k], np.float64)[:total * n_cases].reshape(total, n_cases, -1)[idx]
        cols.append(a.mean(axis=1) + c * a.std(axis=1))
    return np.concatenate(cols, axis=1)


def test_trainer_preprocess_matches_the_reference_block():
    p, cases, out = _run(1, count=60, flag=1)
    data = dataset.to_training_data(dataset.columnar_from_run(p, cases, out))
    for n_cases, c in ((4, 0.0), (6, 0.5)):
        want = _reference_preprocess(data, n_cases, c, 0.8, seed=3)
        got = dataset.trainer_preprocess(data, n_cases, c=c, train_split=0.8, seed=3)
        # columns whose labels are constant up to fp32 rounding (e.g. an inertia the optimiser never moved in
        # 30 epochs) are standardised by a ~1e-7 scale: there the reference's own numbers are rounding noise
        lab = _reference_labels(data, n_cases, c, 0.8, seed=3)
        ok = lab.std(axis=0) > 1e-4 * (np.abs(lab).mean(axis=0) + 1e-30)
        for w, key in zip(want, ("X_train", "X_val", "Y_train", "Y_val")):
            g = got[key].numpy()
            assert g.shape == w.shape and g.dtype == np.float32 and np.isfinite(g).all()
            cols = ok if key.startswith("Y") else slice(None)
            assert np.allclose(g[:, cols], w[:, cols], rtol=1e-4, atol=1e-4), key
        assert ok.mean() > 0.9
    assert isinstance(got["X_train"], torch.Tensor)


# --------------------------------------------------------------------------------------------------
# the reference trainer's OWN data block (OpenPyStruct_PINN_MultiCase.py:1-369 executed verbatim on a dataset the
# product's writer emitted; frozen by tests/golden/make_trainer_golden.py) against dataset.trainer_preprocess
# --------------------------------------------------------------------------------------------------
import os          # noqa: E402
import pytest      # noqa: E402

_TG = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "trainer_goldens.npz"))


def _trainer_golden_case(i):
    col = {k[4:]: _TG[k] for k in _TG.files if k.startswith("col:")}
    n_cases, c, seed, split = _TG[f"{i}:config"]
    ref = {k.split(":", 1)[1]: _TG[k] for k in _TG.files if k.startswith(f"{i}:")}
    return dataset.to_training_data(col), int(n_cases), float(c), int(seed), float(split), ref


def _check_trainer_block(device):
    for i in (0, 1):
        data, n_cases, c, seed, split, ref = _trainer_golden_case(i)
        got = dataset.trainer_preprocess(data, n_cases, c=c, train_split=split, seed=seed, device=device)
        assert got["X_train"].device.type == device
        # the loader half: json.load -> pad -> trim -> reshape by n_cases (PINN:190-258) and the permutation split (:260-264)
        assert np.array_equal(got["train_idx"].cpu().numpy(), ref["train_idx"])
        assert np.array_equal(got["val_idx"].cpu().numpy(), ref["val_idx"])
        # labels whose training column is constant up to fp32 rounding are standardised by a ~1e-7 scale: the
        # reference's own numbers are rounding noise there (sklearn divides by that scale); compared where the scale is real
        Yt = ref["Y_train_std"]
        I_g = ref["I_grouped"][ref["train_idx"]]
        lab = np.concatenate([I_g.mean(axis=1) + c * I_g.std(axis=1)], axis=1)
        real = np.ones(Yt.shape[1], bool)
        real[:lab.shape[1]] = lab.std(axis=0) > 1e-4 * (np.abs(lab).mean(axis=0) + 1e-30)
        for key, want, cols in (("X_train", ref["X_train_flat"], slice(None)), ("X_val", ref["X_val_flat"], slice(None)),
                                ("Y_train", Yt, real), ("Y_val", ref["Y_val_std"], real)):
            g = got[key].cpu().numpy()
            assert g.shape == want.shape and np.isfinite(g).all(), key
            assert np.allclose(g[:, cols], want[:, cols], rtol=2e-4, atol=2e-4), (i, key, np.abs(g[:, cols] - want[:, cols]).max())
        assert real.mean() > 0.9


def test_trainer_block_equals_the_reference_trainers_own_code_cpu():
    _check_trainer_block("cpu")


@pytest.mark.gpu
def test_trainer_block_equals_the_reference_trainers_own_code_on_the_device():
    _check_trainer_block("cuda")
