"""GPU parity: FBP convertor (CUDA, through the C ABI) vs the C oracle and the reference golden."""
import numpy as np
import pytest
import torch

from conftest import golden, rel_l2

pytestmark = pytest.mark.gpu
FBP_TOL = 1e-4          # north_star: "FBP must match within 1e-4 relative L2"


@pytest.fixture(scope="module")
def plan(cuda):
    from ipdm_pytorch_b200 import engine
    return engine.FBPPlan(max_batch=4)


@pytest.fixture(scope="module")
def slice0():
    import ipdm_pytorch_b200.synthetic as S
    return S.make_slice(0)


def test_tables_equal_reference(plan):
    g = golden("fbp_slice0")
    t = plan.tables()
    np.testing.assert_array_equal(t["theta"], g["theta"])
    np.testing.assert_array_equal(t["nda"], g["nda"])
    assert np.abs(t["h_RL"] - g["h_RL"]).max() <= 1e-7 * np.abs(g["h_RL"]).max()
    from oracle import fbp_oracle
    assert np.abs(t["wcos"] - fbp_oracle.Tables().wcos).max() <= 8e-6      # 1 ulp of 59.5


def test_filter_matches_oracle(plan, slice0, cuda):
    from oracle import fbp_oracle
    ld = slice0[0]
    for flip in (True, False):
        w, q = fbp_oracle.weight_and_ramp(ld, flip=flip)
        got = plan.filter(torch.from_numpy(ld)[None].to(cuda), flip=flip).cpu().numpy()[0]
        assert rel_l2(got, q) < 5e-5, flip      # fp32 summation order under heavy cancellation (|h0 p| ~ 140 vs |q| ~ 0.1)


def test_convert_matches_oracle_and_reference_golden(plan, slice0, cuda):
    from oracle import fbp_oracle
    g = golden("fbp_slice0")
    ld = slice0[0]
    got = plan.forward(torch.from_numpy(ld)[None].to(cuda)).cpu().numpy()[0]
    assert np.isfinite(got).all()
    e_ref = rel_l2(got, g["ld"])
    e_orc = rel_l2(got, fbp_oracle.convert(ld[None])[0])
    print(f"FBP rel-L2: vs reference golden {e_ref:.3e}, vs C oracle {e_orc:.3e}")
    assert e_ref < FBP_TOL and e_orc < FBP_TOL


def test_batches_tile_variants_and_flip(plan, slice0, cuda):
    """B = 1, 2, 4 use different tile heights (PX = 1, 2, 4): every slice must equal the B = 1 result bit for bit."""
    import ipdm_pytorch_b200.synthetic as S
    ld0 = torch.from_numpy(slice0[0]).to(cuda)
    ld1 = torch.from_numpy(S.make_slice(1)[0]).to(cuda)
    one0, one1 = plan.forward(ld0[None]), plan.forward(ld1[None])
    two = plan.forward(torch.stack([ld0, ld1]))
    four = plan.forward(torch.stack([ld1, ld0, ld0, ld1]))
    assert torch.equal(two[0], one0[0]) and torch.equal(two[1], one1[0])
    assert torch.equal(four[0], one1[0]) and torch.equal(four[2], one0[0])
    nf = plan.forward(ld0[None], flip=False)
    both = plan.forward(torch.flip(ld0, dims=[1])[None].contiguous(), flip=False)
    assert rel_l2(torch.flip(both, dims=[2]).cpu().numpy(), one0.cpu().numpy()) < 1e-6   # flip == flip in, flip out
    assert not torch.equal(nf, one0)


def test_linearity_and_zero(plan, slice0, cuda):
    """Size-independent properties at full size: FBP is linear; zero in, zero out."""
    ld = torch.from_numpy(slice0[0]).to(cuda)[None]
    nd = torch.from_numpy(slice0[1]).to(cuda)[None]
    a = plan.forward(ld); b = plan.forward(nd); c = plan.forward((0.5 * ld + 2.0 * nd).contiguous())
    assert rel_l2(c.cpu().numpy(), (0.5 * a + 2.0 * b).cpu().numpy()) < 2e-5
    assert float(plan.forward(torch.zeros_like(ld)).abs().max()) == 0.0


def test_reconstructs_the_phantom(plan, slice0, cuda):
    ld, nd, img = slice0
    rec = plan.forward(torch.from_numpy(nd)[None].to(cuda)).cpu().numpy()[0]
    body = img > 0.05
    rmse = float(np.sqrt(((rec - img)[body] ** 2).mean()))
    assert rmse < 0.006, rmse                                          # reference itself: golden nd_rmse_in_body
    assert abs(rmse - float(golden("fbp_slice0")["nd_rmse_in_body"])) < 2e-5


def test_reference_facing_convert_api(slice0, cuda):
    """Recon.FBP_kernel.FBP keeps the reference contract: ndarray/Tensor in, same kind out, on the host."""
    from Recon.FBP_kernel import FBP
    g = golden("fbp_slice0")
    fbp = FBP(device="cuda:0")
    out = fbp.convert(slice0[0])                                       # [2000,912] ndarray
    assert isinstance(out, np.ndarray) and out.shape == (1, 512, 512) and out.dtype == np.float32
    assert rel_l2(out[0], g["ld"]) < FBP_TOL
    out_t = fbp.convert(torch.from_numpy(slice0[0])[None])
    assert isinstance(out_t, torch.Tensor) and not out_t.is_cuda and torch.equal(out_t, torch.from_numpy(out))
    with pytest.raises(ValueError):
        fbp.convert(np.zeros((3, 100, 912), np.float32))
    with pytest.raises(RuntimeError):
        FBP(device="cpu")
