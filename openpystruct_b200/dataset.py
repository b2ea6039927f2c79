"""What sits on either side of the hot path (SURVEY.md 8f "next" rows 1 and 3).

* Dataset writer / loader: the generators end with ``json.dump(training_data, f)`` of a dict of
  13 lists-of-lists (SingleCore:73-87, 263-264) and every trainer starts with ``json.load`` of that file
  (PINN:190-206, FNN:185-197, TFD:238-250).  At 1M beams the JSON text is ~12 GB, so besides the exact
  JSON schema (``save_json``) there is a binary fast path (``save_npz``) and ONE loader,
  ``load_training_data(path)``, that returns the same dict-of-lists for either file -- the trainers'
  ``data = json.load(f)`` line becomes ``data = load_training_data(path)`` and nothing else changes.
* Trainer pre-processing on the device: pad -> trim to a multiple of n_cases -> group consecutive
  records -> train/validation split -> StandardScaler per feature block -> labels ``mean + c * std``
  over the case axis -> scaled targets ``[I, deflections, rotations]`` (PINN:66-92, 226-368; FNN and TFD
  use the same block without the displacement targets).  Plain torch tensor ops on whatever device the
  data is on; no custom kernel (a few MB of elementwise work).
"""
from __future__ import annotations

import json
from typing import Dict, List, Optional, Sequence

import numpy as np
import torch

from .generator import TRAINING_DATA_KEYS
from .params import BeamOptParams

_RAGGED = ("roller_x_locations", "force_x_locations", "force_values", "roller_nodes", "force_nodes")
_DENSE = ("I_values", "shear_forces", "bending_moments", "node_positions", "rotations", "deflections")
_SCALAR = ("num_nodes", "L")


def _ragged_packed(tags, values, dtype, pad):
    """tags: i32 [records, width0] 1-based, 0 = unused (used slots first); values: same shape or None."""
    lens = np.count_nonzero(tags, axis=1).astype(np.int32)
    width = int(lens.max()) if len(lens) else 0
    src = tags if values is None else values
    vals = np.where(np.arange(width)[None, :] < lens[:, None], src[:, :width], pad).astype(dtype)
    return vals, lens


def _x_locations(col, node_positions, nn):
    """roller / force x positions = node_positions[node - 1] (SingleCore:234-235), NaN-padded like the node lists."""
    n_rec = len(col["roller_nodes"])

    def locate(nodes):
        if not n_rec:
            return np.zeros((0, 0))
        return np.where(nodes >= 0, np.take_along_axis(node_positions, np.clip(nodes - 1, 0, nn - 1), axis=1), np.nan)
    col["roller_x_locations"], col["roller_x_locations_len"] = locate(col["roller_nodes"]), col["roller_nodes_len"]
    col["force_x_locations"], col["force_x_locations_len"] = locate(col["force_nodes"]), col["force_nodes_len"]


def case_columns(params: BeamOptParams, cases, keep_b=None) -> Dict[str, np.ndarray]:
    """The record keys that depend on the SAMPLED CASES alone (geometry, supports, loads) of a
    ``sampling.PackedCases``: everything of ``columnar_from_run`` but the kernel outputs.  ``keep_b``: the beams to
    keep (None = all).  ``stream_columnar`` computes this on the sampler's thread while the GPU runs the previous
    batch, so that nothing but views of the kernel outputs is left for the critical path."""
    C = params.num_cases
    nn = params.num_nodes
    B = len(cases) // C
    all_ok = keep_b is None
    if all_ok:
        keep_b = np.arange(B)
    rec = (keep_b[:, None] * C + np.arange(C)[None, :]).reshape(-1)          # record index = beam * C + case
    pick_b = (lambda a: np.asarray(a)[:B]) if all_ok else (lambda a: np.asarray(a)[keep_b])
    pick_r = (lambda a: a) if all_ok else (lambda a: a[rec])
    rep = (lambda a: a) if C == 1 else (lambda a: np.repeat(a, C, axis=0))
    L = pick_b(np.asarray(cases.L, np.float64))
    node_positions = np.zeros((0, nn))
    if len(L) and np.all(L == L[0]):                                         # fixed bridge: one row, broadcast (read-only view)
        node_positions = np.broadcast_to(np.linspace(0, L[0], nn), (len(L), nn))
    elif len(L):                                                             # np.linspace(0, L, nn) bit for bit
        step = L / (nn - 1)
        node_positions = np.arange(nn)[None, :] * step[:, None]
        node_positions[:, -1] = L
    col = {"node_positions": rep(node_positions), "num_nodes": np.full(len(rec), nn, np.int32), "L": rep(L)}
    first = (rec // C) * C                                                   # supports shared by the cases of a beam
    col["roller_nodes"], col["roller_nodes_len"] = _ragged_packed(
        cases.roller_tags if (all_ok and C == 1) else cases.roller_tags[first], None, np.int32, -1)
    ft = pick_r(cases.force_tags)
    col["force_nodes"], col["force_nodes_len"] = _ragged_packed(ft, None, np.int32, -1)
    fv = pick_r(cases.force_vals.reshape(len(cases), -1))
    col["force_values"], col["force_values_len"] = _ragged_packed(ft, fv, np.float64, np.nan)
    _x_locations(col, col["node_positions"], nn)
    return col


def columnar_from_run(params: BeamOptParams, cases, out: Dict[str, np.ndarray],
                      case_cols: Optional[Dict[str, np.ndarray]] = None) -> Dict[str, np.ndarray]:
    """Kernel outputs + the sampled cases -> one array per key of the reference record (failed beams
    dropped, MultiCore:265).  Ragged keys are (values, lengths) pairs padded with NaN / -1.
    ``cases``: the list of ``sampling.sample_case`` tuples, or a ``sampling.PackedCases`` (native sampler), which is
    consumed as arrays -- no per-record Python objects.  ``case_cols``: ``case_columns(params, cases)`` computed ahead
    (used when no beam has to be dropped)."""
    from .sampling import PackedCases
    C = params.num_cases
    packed = isinstance(cases, PackedCases)
    B = len(cases) // C
    keep_b = np.flatnonzero(np.asarray(out["status"][:B]) == 0)
    all_ok = keep_b.size == B                                                # the common case: nothing to drop, no gather copies
    rec = (keep_b[:, None] * C + np.arange(C)[None, :]).reshape(-1)          # record index = beam * C + case
    pick_b = (lambda a: np.asarray(a)[:B]) if all_ok else (lambda a: np.asarray(a)[keep_b])
    rep = (lambda a: a) if C == 1 else (lambda a: np.repeat(a, C, axis=0))
    nn = params.num_nodes
    col = {
        "I_values": rep(pick_b(out["I"])),
        "shear_forces": pick_b(out["shear"]).reshape(len(rec), -1),
        "bending_moments": pick_b(out["moment"]).reshape(len(rec), -1),
        "rotations": pick_b(out["rot"]).reshape(len(rec), -1),
        "deflections": pick_b(out["defl"]).reshape(len(rec), -1),
    }
    if packed:
        col.update(case_cols if (case_cols is not None and all_ok) else case_columns(params, cases, None if all_ok else keep_b))
        return col

    L = np.array([cases[b * C][0] for b in keep_b], np.float64)
    node_positions = np.stack([np.linspace(0, l_, nn) for l_ in L]) if len(L) else np.zeros((0, nn))

    def ragged(get, dtype, pad):
        rows = [get(i) for i in rec]
        width = max((len(r) for r in rows), default=0)
        vals = np.full((len(rows), width), pad, dtype)
        lens = np.zeros(len(rows), np.int32)
        for i, r in enumerate(rows):
            vals[i, :len(r)] = r
            lens[i] = len(r)
        return vals, lens

    col.update({"node_positions": rep(node_positions), "num_nodes": np.full(len(rec), nn, np.int32), "L": rep(L)})
    rollers = lambda i: cases[(i // C) * C][1]                               # noqa: E731  (supports shared by the cases)
    col["roller_nodes"], col["roller_nodes_len"] = ragged(rollers, np.int32, -1)
    col["force_nodes"], col["force_nodes_len"] = ragged(lambda i: cases[i][2], np.int32, -1)
    col["force_values"], col["force_values_len"] = ragged(lambda i: cases[i][3], np.float64, np.nan)
    _x_locations(col, col["node_positions"], nn)
    return col


def to_training_data(col: Dict[str, np.ndarray]) -> Dict[str, list]:
    """Columnar arrays -> the reference's dict of lists (SingleCore:73-87), key order included."""
    data = {}
    for k in TRAINING_DATA_KEYS:
        if k in _RAGGED:
            vals, lens = col[k], col[k + "_len"]
            data[k] = [row[:n].tolist() for row, n in zip(vals, lens)]
        else:
            data[k] = col[k].tolist()
    return data


def save_json(col: Dict[str, np.ndarray], path: str) -> None:
    with open(path, "w") as f:
        json.dump(to_training_data(col), f)


def save_npz(col: Dict[str, np.ndarray], path: str) -> None:
    np.savez(path, **col)


def load_training_data(path: str) -> Dict[str, list]:
    """The dict the trainers expect from ``json.load`` (PINN:192-206), from either file format."""
    if str(path).endswith(".npz"):
        with np.load(path) as z:
            return to_training_data({k: z[k] for k in z.files})
    with open(path) as f:
        return json.load(f)


# --------------------------------------------------------------------------------------------------
# trainer pre-processing (PINN:66-92, 226-368)
# --------------------------------------------------------------------------------------------------
def pad_sequences(rows: Sequence[Sequence[float]], max_length: int, pad_val: float = 0.0) -> np.ndarray:
    """PINN:66-76: float32 [num_samples, max_length], zero padded."""
    out = np.full((len(rows), max_length), pad_val, np.float32)
    for i, r in enumerate(rows):
        a = np.asarray(r, np.float32)[:max_length]
        out[i, :len(a)] = a
    return out


class _Scaler:
    """sklearn.preprocessing.StandardScaler semantics: population variance, zero scale -> 1."""

    def fit(self, x2d: torch.Tensor) -> "_Scaler":
        x = x2d.to(torch.float64)
        self.mean = x.mean(dim=0)
        var = x.var(dim=0, unbiased=False)
        scale = var.sqrt()
        self.scale = torch.where(scale < 10 * torch.finfo(torch.float64).eps, torch.ones_like(scale), scale)
        return self

    def transform(self, x: torch.Tensor) -> torch.Tensor:
        shp = x.shape
        y = (x.reshape(-1, shp[-1]).to(torch.float64) - self.mean) / self.scale
        return y.to(torch.float32).reshape(shp)


def unify_label_with_c(x3d: torch.Tensor, c: float) -> torch.Tensor:
    """PINN:79-92: mean over the case axis + c * population std."""
    return x3d.mean(dim=1) + c * x3d.std(dim=1, unbiased=False)


def trainer_preprocess(data: Dict[str, list], n_cases: int, c: float = 0.0, train_split: float = 0.8,
                       seed: Optional[int] = None, device="cpu", with_displacements: bool = True) -> Dict[str, object]:
    """The trainers' data block up to the tensors handed to the DataLoader.  ``seed`` seeds the
    ``np.random.permutation`` of the groups (the reference relies on the global numpy state)."""
    feats = {"roller_x": "roller_x_locations", "force_x": "force_x_locations", "force_values": "force_values",
             "node_positions": "node_positions"}
    targets = {"I": "I_values"}
    if with_displacements:
        targets.update({"deflections": "deflections", "rotations": "rotations"})
    num_samples = len(data["I_values"])
    total_grouped = num_samples // n_cases
    if total_grouped == 0:
        raise ValueError(f"n_cases={n_cases} > total samples={num_samples}.")
    trim = total_grouped * n_cases
    dev = torch.device(device)

    def grouped(key):
        rows = data[key]
        width = max((len(r) for r in rows), default=0)
        pad = pad_sequences(rows, width)[:trim]
        return torch.from_numpy(pad).to(dev).reshape(total_grouped, n_cases, -1)

    rng = np.random.RandomState(seed) if seed is not None else np.random
    idx = torch.from_numpy(rng.permutation(total_grouped)).to(dev)
    n_train = int(train_split * total_grouped)
    tr, va = idx[:n_train], idx[n_train:]
    scalers, x_tr, x_va = {}, [], []
    for name, key in feats.items():
        g = grouped(key)
        s = _Scaler().fit(g[tr].reshape(-1, g.shape[-1]))
        scalers[name] = s
        x_tr.append(s.transform(g[tr]))
        x_va.append(s.transform(g[va]))
    X_train = torch.cat(x_tr, dim=2)
    X_val = torch.cat(x_va, dim=2)
    y_tr, y_va = [], []
    for name, key in targets.items():
        g = grouped(key)
        lab_tr, lab_va = unify_label_with_c(g[tr], c), unify_label_with_c(g[va], c)
        s = _Scaler().fit(lab_tr)
        scalers[name] = s
        y_tr.append(s.transform(lab_tr))
        y_va.append(s.transform(lab_va))
    return {
        "X_train": X_train.reshape(X_train.shape[0], -1), "X_val": X_val.reshape(X_val.shape[0], -1),
        "X_train_3d": X_train, "X_val_3d": X_val,
        "Y_train": torch.cat(y_tr, dim=1), "Y_val": torch.cat(y_va, dim=1),
        "train_idx": tr, "val_idx": va, "scalers": scalers,
    }
