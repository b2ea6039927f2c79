"""The C oracle's FP32 half against the REAL torch CPU ops it restates (SingleCore:195-208)."""
import numpy as np
import torch
from torch.optim.lr_scheduler import ExponentialLR

from oracle import c_oracle

E, G = 200e9, 200e9 / 2.6
f32 = np.float32


def random_state(rng, n=100):
    I = np.exp(rng.uniform(np.log(3e-3), np.log(0.9), n)).astype(f32)
    M = (rng.normal(size=n) * 3e6).astype(f32)
    V = (rng.normal(size=n) * 2e5).astype(f32)
    return I, M, V


def test_torch_sum_bitwise():
    rng = np.random.default_rng(0)
    for n in (8, 31, 32, 33, 100, 101, 127, 128, 129, 511, 512, 513, 1000, 1001, 2048, 4099):
        for _ in range(20):
            x = (rng.normal(size=n) * rng.choice([1.0, 1e3, 1e-3])).astype(f32)
            assert c_oracle.torch_sum(x) == torch.sum(torch.from_numpy(x)).item(), n


def test_loss_bitwise_and_grad_within_sqrt_ulp():
    rng = np.random.default_rng(1)
    p = c_oracle.make_params()
    exact = total = 0
    for _ in range(100):
        I, M, V = random_state(rng)
        It = torch.tensor(I, requires_grad=True)
        Mt, Vt = torch.tensor(M), torch.tensor(V)
        be = torch.sum((Mt ** 2) / (2 * E * It + 1e-6))
        A = 0.03 * It ** 0.5
        se = torch.sum(Vt ** 2 / (G * A))
        tot = torch.sum(It) + 1e-2 * be + 1e-2 * se
        tot.backward()
        loss, grad = c_oracle.loss_grad(p, I, M * M, V * V)
        assert loss == tot.item()
        g = It.grad.numpy()
        exact += int(np.sum(g == grad))
        total += g.size
        # the only admissible deviation: torch's MKL sqrt is 1 ulp off IEEE on <1 % of inputs
        # (measured against the magnitude of the summed terms: grad = 1 - shear term - bending term cancels)
        I64, M64, V64 = I.astype(np.float64), M.astype(np.float64), V.astype(np.float64)
        mag = 1 + 1e-2 * M64 ** 2 * 2 * E / (2 * E * I64) ** 2 + 1e-2 * V64 ** 2 * 0.5 / (G * 0.03 * I64 ** 1.5)
        assert np.max(np.abs(g - grad) / mag) < 3e-7
    assert exact / total > 0.99


def test_adam_schedule_matches_torch_scalars():
    p = c_oracle.make_params(max_epochs=600)
    tab = c_oracle.adam_schedule(p)
    x = torch.zeros(1, requires_grad=True)
    opt = torch.optim.Adam([x], lr=0.01)
    sch = ExponentialLR(opt, gamma=0.98)
    for t in range(1, 601):
        lr = opt.param_groups[0]["lr"]
        bc1, bc2 = 1 - 0.9 ** t, 1 - 0.999 ** t
        assert tab[t - 1, 0] == f32(-(lr / bc1))
        assert tab[t - 1, 1] == f32(bc2 ** 0.5)
        x.grad = torch.ones(1)
        opt.step()
        sch.step()


def test_adam_step_bitwise_m_v_and_I_within_sqrt_ulp():
    rng = np.random.default_rng(2)
    p = c_oracle.make_params(max_epochs=16)
    tab = c_oracle.adam_schedule(p)
    exact = total = 0
    for _ in range(20):
        I0, _, _ = random_state(rng)
        It = torch.tensor(I0.copy(), requires_grad=True)
        opt = torch.optim.Adam([It], lr=0.01)
        sch = ExponentialLR(opt, gamma=0.98)
        I, m, v = I0.copy(), np.zeros(100, f32), np.zeros(100, f32)
        for t in range(8):
            g = (rng.normal(size=100) * rng.choice([1e-3, 1.0, 30.0])).astype(f32)
            It.grad = torch.tensor(g.copy())
            opt.step(); sch.step()
            with torch.no_grad():
                It.clamp_(min=1e-8)
            c_oracle.adam_step(p, tab[t, 0], tab[t, 1], g, I, m, v)
            st = opt.state[It]
            assert np.array_equal(st["exp_avg"].numpy(), m)
            assert np.array_equal(st["exp_avg_sq"].numpy(), v)
            ref = It.detach().numpy()
            exact += int(np.sum(ref == I)); total += 100
            assert np.max(np.abs(ref - I) / np.abs(ref)) < 3e-7
            I[:] = ref     # resync so single-step deviations do not accumulate
    assert exact / total > 0.99
