"""GPU parity: sampler kernels (through the C ABI) vs the oracle restatement of Model/model.py:438-515, 596-600."""
import math

import numpy as np
import pytest
import torch

from conftest import rel_l2
from inputs import small_proj_input

pytestmark = pytest.mark.gpu


def _fields(shape, seed, dev):
    g = torch.Generator().manual_seed(seed)
    x0 = torch.rand(shape, generator=g) * 3
    xt = x0 + 0.08 * torch.randn(shape, generator=g)
    eps = 0.3 + 1.7 * torch.randn(shape, generator=g)
    nz = torch.randn(shape, generator=g)
    return [t.contiguous() for t in (xt, x0, eps, nz)]


@pytest.mark.parametrize("shape", [(1, 1, 100, 76), (3, 1, 64, 64), (2, 1, 2000, 912)])
@pytest.mark.parametrize("t,lam,clip", [(14, 0.999, False), (3, 0.45, True), (0, 0.0133, False)])
def test_step_scalar_lambda(cuda, shape, t, lam, clip):
    from ipdm_pytorch_b200 import engine
    from oracle import ipdm_oracle as O
    xt, x0, eps, nz = _fields(shape, 5 + t, cuda)
    tab = O.Tables(1000, 5)
    want = torch.stack([O.p_sample_condition(tab, eps[b:b + 1], xt[b:b + 1], x0[b:b + 1], t, lam, clip, nz[b:b + 1])[0] for b in range(shape[0])])
    coef = engine.step_coefficients(1000, 5, t)
    got = engine.sampler_step(xt.to(cuda), x0.to(cuda), eps.to(cuda), coef, lam, noise=nz.to(cuda), clip=clip, t_nonzero=t != 0)
    assert rel_l2(got.cpu().numpy(), want.numpy()) < 3e-6


@pytest.mark.parametrize("shape", [(1, 1, 100, 76), (2, 1, 2000, 912)])
def test_step_lambda_map(cuda, shape):
    from ipdm_pytorch_b200 import engine
    from oracle import ipdm_oracle as O
    xt, x0, eps, nz = _fields(shape, 11, cuda)
    b, _, h, w = shape
    lam = 0.05 + 0.94 * torch.rand(b, h // 4, w // 4, generator=torch.Generator().manual_seed(2))
    tab = O.Tables(1000, 5)
    up = torch.nn.functional.interpolate(lam[:, None], size=(h, w), mode="nearest")
    want = torch.stack([O.p_sample_condition(tab, eps[i:i + 1], xt[i:i + 1], x0[i:i + 1], 7, up[i:i + 1], False, nz[i:i + 1])[0] for i in range(b)])
    got = engine.sampler_step(xt.to(cuda), x0.to(cuda), eps.to(cuda), engine.step_coefficients(1000, 5, 7), lam.to(cuda).contiguous(),
                              noise=nz.to(cuda), clip=False, t_nonzero=True, ks=4)
    assert rel_l2(got.cpu().numpy(), want.numpy()) < 3e-6


def test_in_place_and_batch_independence(cuda):
    from ipdm_pytorch_b200 import engine
    xt, x0, eps, nz = [t.to(cuda) for t in _fields((2, 1, 64, 64), 3, cuda)]
    coef = engine.step_coefficients(1000, 1, 5)
    ref = engine.sampler_step(xt, x0, eps, coef, 0.45, noise=nz, clip=True)
    one = engine.sampler_step(xt[1:].contiguous(), x0[1:].contiguous(), eps[1:].contiguous(), coef, 0.45, noise=nz[1:].contiguous(), clip=True)
    assert torch.equal(ref[1:], one)                                   # per-slice statistics
    x_in = xt.clone()
    engine.sampler_step(x_in, x0, eps, coef, 0.45, noise=nz, clip=True, out=x_in)
    assert torch.equal(x_in, ref)


def test_q_sample_lincomb_clamp(cuda):
    from ipdm_pytorch_b200 import engine
    g = torch.Generator().manual_seed(0)
    x, y, z = [torch.randn(2, 1, 50, 36, generator=g) for _ in range(3)]
    a, b = np.float32(0.996896), np.float32(0.078734)
    got = engine.q_sample(x.to(cuda), a, b, noise=y.to(cuda)).cpu()
    assert torch.equal(got, float(a) * x + float(b) * y)
    got = engine.lincomb(0.7, x.to(cuda), 0.25, y.to(cuda), 0.05, z.to(cuda)).cpu()
    assert torch.equal(got, 0.7 * x + (0.95 - 0.7) * y + 0.05 * z)    # model.py:635 evaluation order
    c = engine.clamp_(x.to(cuda).clone(), 0.0, math.inf).cpu()
    assert torch.equal(c, x.clamp(min=0))


def test_philox_noise_statistics(cuda):
    from ipdm_pytorch_b200 import engine
    x = torch.zeros(2, 1, 512, 512, device=cuda)
    n1 = engine.q_sample(x, 1.0, 1.0, noise=None, seed=123, call_id=0)
    n2 = engine.q_sample(x, 1.0, 1.0, noise=None, seed=123, call_id=1)
    n1b = engine.q_sample(x, 1.0, 1.0, noise=None, seed=123, call_id=0)
    assert torch.equal(n1, n1b) and not torch.equal(n1, n2) and not torch.equal(n1[0], n1[1])
    v = n1.flatten().double()
    assert abs(float(v.mean())) < 5e-3 and abs(float(v.std()) - 1) < 5e-3
    assert abs(float((v ** 3).mean())) < 2e-2 and abs(float((v ** 4).mean()) - 3) < 5e-2
    assert abs(float((n1.flatten() * n2.flatten()).mean())) < 5e-3


@pytest.mark.parametrize("shape", [(1, 1, 100, 76), (2, 1, 2000, 912)])
def test_delta_lambda_map_and_median(cuda, shape):
    from ipdm_pytorch_b200 import engine
    from oracle import ipdm_oracle as O
    b, _, h, w = shape
    g = torch.Generator().manual_seed(4)
    img = torch.rand(shape, generator=g) * 3
    x = img + 0.05 * torch.randn(shape, generator=g) * (1 + 4 * (torch.rand(shape, generator=g) > 0.97))
    lam, med = engine.delta_lambda_map(x.to(cuda), img.to(cuda), ks=4, amplitude=7.0, kind="proj", return_median=True)
    for i in range(b):
        d = torch.abs(x[i:i + 1] - img[i:i + 1])
        m = torch.median(d)
        assert float(med[i]) == float(m)                               # exact: radix select returns the lower middle element
        dd = torch.nn.functional.avg_pool2d(d - m, 4)
        dd = torch.where(dd <= 0, torch.zeros_like(dd), dd)
        want = O.lambda_curve(torch.exp(7.0 * dd).numpy(), "proj")[0, 0]
        assert np.abs(lam[i].cpu().numpy() - want).max() < 3e-4 * max(1.0, np.abs(want).max())


def test_lambda_step_map(cuda):
    from ipdm_pytorch_b200 import engine
    from oracle import ipdm_oracle as O
    lam_exp = torch.linspace(-0.19, 19.95, 5000)
    for i, ts in ((14, 15), (0, 15), (2, 5), (7, 10)):
        got = engine.lambda_step_map(lam_exp.to(cuda), i, ts).cpu().numpy()
        want = O.condition_lambda_map(lam_exp.numpy(), i, ts)
        assert np.abs(got - want).max() <= 1.2e-7


def test_sharpen(cuda):
    from ipdm_pytorch_b200 import engine
    from oracle import ipdm_oracle as O
    x = torch.randn(3, 1, 512, 512, generator=torch.Generator().manual_seed(1))
    for N in (42, 70):
        got = engine.sharpen3x3(x.to(cuda), N).cpu()
        want = torch.cat([O.tensor_sharpen(x[i:i + 1], N) for i in range(3)])
        assert rel_l2(got.numpy(), want.numpy()) < 1e-6
    assert torch.equal(engine.sharpen3x3(x.to(cuda), -1).cpu(), x)
