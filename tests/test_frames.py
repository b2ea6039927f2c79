"""Frame optimiser.  CPU: host mirror of the script's constants and draw order, the fp32 recipe of
the kernel against real torch autograd (bit for bit), the C ABI's host-side entries.  GPU (`-m gpu`): the CUDA path
through the C ABI against the fixtures frozen from the reference's own source (tests/golden/frame_goldens.npz) and
against the torch-path restatement (oracle/frameopt_port.py) on other frames."""
import ctypes as C
import os
import random

import numpy as np
import pytest
import torch

from openpystruct_b200 import _cabi, frames
from oracle import c_oracle
from oracle import frameopt_port as fp

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = np.load(os.path.join(HERE, "golden", "frame_goldens.npz"))
f32 = np.float32


def golden_params(key, **kw):
    c = GOLD[key + "_consts"]
    names = ("E", "G", "A", "I0", "alpha_moment", "alpha_shear", "k", "lateral_load", "vertical_load", "lr", "tolerance",
             "bay_width", "story_height")
    d = dict(zip(names, (float(x) for x in c)))
    d.pop("G")
    return frames.FrameOptParams(patience=int(GOLD[key + "_patience"][0]), **d).replace(**kw)


# ------------------------------------------------------------------------------------------------ CPU
def test_constants_and_draw_order_are_the_scripts():
    p = frames.FrameOptParams()
    assert (p.max_bays, p.max_stories, p.num_epochs, p.lr, p.tolerance, p.patience) == (10, 10, 5000, 0.005, 1e-3, 10)
    assert p.G == pytest.approx(200e9 / 2.6)
    random.seed(11)
    want = (random.randint(1, 10), random.randint(1, 10))           # :50-51
    random.seed(11)
    assert frames.draw_frame(p) == want == tuple(int(x) for x in GOLD["s11_shape"][:2])
    assert frames.frame_counts(8, 9) == (81, 72) and frames.max_elements(p) == 210


def test_struct_layout_and_host_entries():
    cp = frames.to_c_params(frames.FrameOptParams())
    assert C.sizeof(cp) == 6 * 4 + 18 * 8 and cp.struct_size == C.sizeof(cp)
    assert _cabi.lib().ops_frameopt_max_elements(C.byref(cp)) == 210
    p = frames.FrameOptParams(num_epochs=50)
    tab = frames.fill_schedule(p)
    opt = torch.optim.Adam([torch.zeros(1, requires_grad=True)], lr=p.lr)
    for t in range(1, 51):                                          # torch's _single_tensor_adam scalars, in double
        bc1, bc2 = 1 - 0.9 ** t, 1 - 0.999 ** t
        assert tab[t - 1, 0] == f32(-(opt.param_groups[0]["lr"] / bc1)) and tab[t - 1, 1] == f32(bc2 ** 0.5)


def test_compute_entries_refuse_without_a_device():
    if torch.cuda.is_available():
        pytest.skip("CUDA device present")
    with pytest.raises(RuntimeError):
        frames.optimise_frames([(1, 1)])


def kernel_recipe(I0, M, V, p=frames.FrameOptParams()):
    """The fp32 half of csrc/frameopt.cu in numpy, one rounded operation per line (header comment of the kernel)."""
    I = I0.astype(f32)
    E2, eps, Gf, kf, am, as_ = f32(2 * p.E), f32(p.bending_eps), f32(p.G), f32(p.k), f32(p.alpha_moment), f32(p.alpha_shear)
    be, se, g = f32(0), f32(0), np.zeros(len(I), f32)
    for e in range(len(I)):
        b = f32(f32(E2 * I[e]) + eps); rb = f32(f32(1) / b); c = f32(M[e] ** 2)
        s = f32(np.sqrt(I[e])); gg = f32(Gf * f32(kf * s)); rg = f32(f32(1) / gg); h = f32(V[e] ** 2)
        be = f32(be + f32(rb * c)); se = f32(se + f32(rg * h))
        gb = f32(f32(f32(-f32(am * c)) * f32(rb * rb)) * E2)
        gs = f32(f32(f32(f32(f32(-f32(as_ * h)) * f32(rg * rg)) * Gf) * kf) * f32(f32(0.5) * f32(f32(1) / s)))
        g[e] = f32(f32(1) + f32(gs + gb))
    total = f32(f32(c_oracle.torch_sum(I) + f32(am * be)) + f32(as_ * se))
    return total, g


def test_kernel_fp32_recipe_is_torch_autograd_bit_for_bit():
    """compute_combined_loss + backward of the reference (:141-160, 184) on random members against the recipe the kernel
    implements.  torch's `** 0.5` is not correctly rounded on ~0.7 % of its inputs (SURVEY finding 5), which the IEEE
    square root of the kernel cannot reproduce: the loss must match always, the gradient on >= 99 % of the members."""
    torch.set_num_threads(1)
    p = frames.FrameOptParams()
    rng = np.random.default_rng(0)
    same, total, losses = 0, 0, 0
    for _ in range(25):
        n = int(rng.integers(3, 211))
        I0 = np.exp(rng.uniform(np.log(1e-4), np.log(0.1), n)).astype(f32)
        M, V = rng.normal(0, 3e5, n), rng.normal(0, 3e5, n)
        I = torch.tensor(I0, requires_grad=True)
        be, se = 0.0, 0.0
        for e in range(n):                                          # the reference's statements, verbatim semantics
            I_val = I[e]
            be += (M[e] ** 2) / (2 * p.E * I_val + 1e-8)
            A_local = p.k * (I_val ** 0.5)
            se += (V[e] ** 2) / (p.G * A_local)
        tot = torch.sum(I) + p.alpha_moment * be + p.alpha_shear * se
        tot.backward()
        t, g = kernel_recipe(I0, M, V, p)
        losses += int(f32(tot.item()) == t)
        same += int((I.grad.numpy() == g).sum()); total += n
    assert losses == 25
    assert same / total > 0.99, same / total


# ------------------------------------------------------------------------------------------------ GPU
def run(frames_list, p):
    return frames.optimise_frames(frames_list, p, device="cuda")


@pytest.mark.gpu
@pytest.mark.parametrize("key", ["s2", "s4", "s11"])
def test_reference_runs(key):
    """The three frozen runs of the reference script (1x2 early-stopped after 228 of at most 400 epochs; 4x5 and 8x9
    capped at 250 epochs): epoch count, the whole loss history, opt_I and the last analysis' member forces."""
    bays, stories, cap, ran = (int(x) for x in GOLD[key + "_shape"])
    p = golden_params(key, num_epochs=cap)
    r = run([(bays, stories)], p)[0]
    assert r["status"] == 0 and r["epochs"] == ran
    want = GOLD[key + "_loss"]
    assert np.max(np.abs(r["loss_history"] - want) / want) < 2e-6
    assert (r["loss_history"].astype(f32) == want.astype(f32)).mean() > 0.5      # (torch's `** 0.5` is not IEEE on ~0.7 % of inputs)
    assert np.max(np.abs(r["opt_I"] - GOLD[key + "_I"]) / GOLD[key + "_I"]) < 1e-5
    assert np.max(np.abs(r["bending_moments"] - GOLD[key + "_M_trace"][-1])) / np.max(np.abs(GOLD[key + "_M_trace"][-1])) < 1e-6
    assert np.max(np.abs(r["shear_forces"] - GOLD[key + "_V_trace"][-1])) / np.max(np.abs(GOLD[key + "_V_trace"][-1])) < 1e-6


@pytest.mark.gpu
def test_first_analysis_forces_1e9():
    """One epoch from I0: the member forces of the FP64 solve against the reference's first analysis (shim, LU)."""
    for key in ("s2", "s4", "s11"):
        bays, stories = (int(x) for x in GOLD[key + "_shape"][:2])
        r = run([(bays, stories)], golden_params(key, num_epochs=1))[0]
        M0, V0 = GOLD[key + "_M_trace"][0], GOLD[key + "_V_trace"][0]
        assert np.max(np.abs(r["bending_moments"] - M0)) / np.max(np.abs(M0)) < 1e-9
        assert np.max(np.abs(r["shear_forces"] - V0)) / np.max(np.abs(V0)) < 1e-9


@pytest.mark.gpu
@pytest.mark.parametrize("frame,epochs", [((1, 1), 120), ((3, 2), 80), ((2, 7), 60), ((10, 10), 30)])
def test_against_the_torch_path_restatement(frame, epochs):
    p = frames.FrameOptParams(num_epochs=epochs, early_stop=False)
    o = fp.frame_optimise(frame[0], frame[1], fp.FrameParams(num_epochs=epochs, patience=10 ** 9))
    r = run([frame], p)[0]
    assert r["epochs"] == epochs == o["epochs"] and r["status"] == 0
    assert np.max(np.abs(r["loss_history"] - o["loss"]) / o["loss"]) < 2e-6
    assert np.max(np.abs(r["opt_I"] - o["I"]) / o["I"]) < 1e-5


@pytest.mark.gpu
def test_batch_properties_and_edge_cases():
    p = frames.FrameOptParams(num_epochs=300)
    rng = random.Random(3)
    batch = [frames.draw_frame(p, rng) for _ in range(300)] + [(10, 10), (1, 1), (1, 10), (10, 1)]
    a, b = run(batch, p), run(batch, p)
    for x, y, fr in zip(a, b, batch):
        assert x["status"] == 0 and x["epochs"] == y["epochs"] and np.array_equal(x["opt_I"], y["opt_I"])     # deterministic
        assert np.array_equal(x["loss_history"], y["loss_history"])
        n = sum(frames.frame_counts(*fr))
        assert x["opt_I"].shape == (n,) and (x["opt_I"] >= f32(1e-8)).all() and np.isfinite(x["loss_history"]).all()
        assert x["loss_history"][-1] < x["loss_history"][0] and x["best_loss"] <= x["loss_history"][0]
    one = run([batch[7]], p)[0]                                       # any split of the batch gives the same bytes
    assert np.array_equal(one["opt_I"], a[7]["opt_I"]) and one["epochs"] == a[7]["epochs"]
    assert run([], p) == []
    bad = run([(11, 3), (3, 0)], p)
    assert [r["status"] for r in bad] == [2, 2] and [r["epochs"] for r in bad] == [0, 0]


@pytest.mark.gpu
def test_run_host_entry_matches_the_device_pointer_entry():
    p = frames.FrameOptParams(num_epochs=40)
    batch = [(2, 3), (5, 1), (4, 4)]
    want = run(batch, p)
    cp, me = frames.to_c_params(p), frames.max_elements(p)
    B = len(batch)
    nb = np.array([f[0] for f in batch], np.int32); ns = np.array([f[1] for f in batch], np.int32)
    I = np.zeros((B, me), f32); hist = np.zeros((B, 40), f32); M = np.zeros((B, me)); V = np.zeros((B, me))
    best = np.zeros(B); ep = np.zeros(B, np.int32); st = np.zeros(B, np.int32); ms = C.c_float()
    rc = _cabi.lib().ops_frameopt_run_host(C.byref(cp), B, nb.ctypes.data, ns.ctypes.data, I.ctypes.data, hist.ctypes.data,
                                          M.ctypes.data, V.ctypes.data, best.ctypes.data, ep.ctypes.data, st.ctypes.data,
                                          0, C.byref(ms))
    assert rc == 0 and ms.value > 0
    for i, w in enumerate(want):
        n = len(w["opt_I"])
        assert np.array_equal(I[i, :n], w["opt_I"]) and ep[i] == w["epochs"] and best[i] == w["best_loss"]
