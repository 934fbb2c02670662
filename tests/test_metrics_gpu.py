"""PSNR / SSIM on the device against the numpy restatement of the reference's skimage calls
(metric_calculate, Utils/train_test_utils.py:789-799)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("shape", [(1, 512, 512), (3, 67, 45), (2, 11, 11), (2, 130, 33)])
def test_psnr_ssim_match_the_oracle(cuda, shape):
    from ipdm_pytorch_b200 import engine
    from oracle import metrics_oracle as M
    rng = np.random.default_rng(shape[1])
    ref = rng.random(shape).astype(np.float32)
    test = np.clip(ref + 0.03 * rng.standard_normal(shape), 0, 1).astype(np.float32)
    test[0, 2, 3] = np.nan                                              # counts as 0.5
    got = engine.psnr_ssim(torch.from_numpy(test).to(cuda), torch.from_numpy(ref).to(cuda), win_size=11).cpu().numpy()
    for b in range(shape[0]):
        assert got[b, 0] == pytest.approx(M.psnr(ref[b], test[b]), abs=1e-9)
        assert got[b, 1] == pytest.approx(M.ssim(ref[b], test[b], win_size=11), abs=1e-9)
    same = engine.psnr_ssim(torch.from_numpy(ref).to(cuda), torch.from_numpy(ref).to(cuda)).cpu().numpy()
    assert np.isinf(same[:, 0]).all() and np.allclose(same[:, 1], 1.0, atol=1e-12)


def test_miu2pixel_matches_the_oracle(cuda):
    from ipdm_pytorch_b200 import engine
    from oracle import metrics_oracle as M
    rng = np.random.default_rng(1)
    mu = (1.2 * rng.random((3, 40, 50)) - 0.05).astype(np.float32)      # spans below -1024 HU and above 3072 HU
    mu[1, 2, 3] = np.nan
    got = engine.miu2pixel(torch.from_numpy(mu).to(cuda)).cpu().numpy()
    want = M.miu2pixel(mu)
    want[np.isnan(want)] = 0.5
    assert np.abs(got - want).max() <= 1.2e-7                            # one fp32 ulp of the division
    assert got.min() == 0.0 and got.max() == 1.0


def test_bad_arguments_are_rejected(cuda):
    from ipdm_pytorch_b200 import engine
    a = torch.zeros(1, 8, 8, device=cuda)
    with pytest.raises(RuntimeError):
        engine.psnr_ssim(a, a, win_size=11)                              # window larger than the image
    with pytest.raises(ValueError):
        engine.psnr_ssim(a, torch.zeros(1, 8, 9, device=cuda))


def test_metric_calculate_through_the_reference_api(cuda, tmp_path):
    """result_figure_save(mode="progressive") fills metric_instance with the reference's keys (LDCT.psnr_iter_0, deProg.ssim_iter_k)."""
    import os
    import sys
    pkg = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "ipdm-pytorch_b200")
    sys.path.insert(0, pkg)
    import ipdm_pytorch_b200.synthetic as S
    from oracle import metrics_oracle as M
    from test_progressive_gpu import _model
    model = _model(tmp_path, dict(t_start_proj=[2, 1], t_start_img=[2, 1], noise_seed=5))
    noisy, _, ndct = S.make_slice(0)
    ld = torch.from_numpy(noisy)[None, None]
    fbp_ld = torch.from_numpy(ndct)[None, None] * 1.02
    model.data_sample_load(ldct=fbp_ld, ldproj=ld, fdproj=None, fdct=torch.from_numpy(ndct)[None, None])
    model.progressive_denoiser()
    model.result_figure_save(mode="progressive", display=False, only_metric=True)
    mi = model.metric_instance
    assert "psnr_iter_0" in mi["LDCT"] and "ssim_iter_0" in mi["LDCT"]
    assert mi["LDCT"]["psnr_iter_0"] == pytest.approx(M.psnr(M.miu2pixel(ndct), M.miu2pixel(ndct * np.float32(1.02))), abs=1e-6)
    keys = [k for k in mi["deProg"] if k.startswith("ssim_iter_")]
    assert len(keys) >= 1 and all(np.isfinite(mi["deProg"][k]) for k in keys)
