"""CPU: pins the oracle restatement (oracle/) to golden vectors produced by the UNMODIFIED reference
(oracle/make_golden.py, run in the build container).  The reference ships no tests of its own."""
import numpy as np
import pytest
import torch

from conftest import golden, rel_l2
from inputs import IMG_CFG, PROJ_CFG, noise_tape, small_img_input, small_proj_input, unet_small_input
from oracle import fbp_oracle, ipdm_oracle as O


def test_schedule_tables_match_reference():
    g = golden("schedule")
    for name, p in (("proj", 5), ("img", 1)):
        tab = O.Tables(1000, p)
        for attr in ("betas", "alphas_cumprod", "sqrt_alphas_cumprod", "sqrt_one_minus_alphas_cumprod", "sqrt_recip_alphas_cumprod",
                     "sqrt_recipm1_alphas_cumprod", "posterior_variance", "posterior_log_variance_clipped", "posterior_mean_coef1",
                     "posterior_mean_coef2"):
            np.testing.assert_array_equal(getattr(tab, attr).numpy()[:64], g[f"{name}_{attr}"])
    np.testing.assert_array_equal(O.cosine_beta_schedule(15, schedule_power=1).numpy(), g["lambda_cosine_15"])
    np.testing.assert_array_equal(O.cosine_beta_schedule(5, schedule_power=10).numpy(), g["lambda_cosine_5_p10"])


def test_known_answer_constants():
    """SURVEY.md Appendix A table (computed from the reference formulas)."""
    p, i = O.Tables(1000, 5), O.Tables(1000, 1)
    assert abs(float(p.sqrt_one_minus_alphas_cumprod[15]) - 0.078734) < 5e-7
    assert abs(float(i.sqrt_one_minus_alphas_cumprod[15]) - 0.035255) < 5e-7
    assert abs(float(p.posterior_mean_coef1[5]) - 0.204554) < 5e-7
    assert abs(float(i.posterior_mean_coef2[14]) - 0.903209) < 5e-7
    lam = O.cosine_beta_schedule(15, schedule_power=1).numpy()
    assert abs(lam[0] - 0.0133) < 5e-5 and lam[14] == 0.999


def test_lambda_curves_match_reference():
    g = golden("curves")
    for kind in ("proj", "img"):
        np.testing.assert_allclose(O.lambda_curve(g["x"], kind), g[kind], rtol=2e-6, atol=1e-6)
    z1, z2 = O.curve_coefficients("proj")
    assert abs(np.polyval(z1, 1.0) - 19.953) < 2e-3 and abs(np.polyval(z2, 2.75) + 0.190) < 2e-3


@pytest.mark.parametrize("name,cfg", [("proj", PROJ_CFG), ("img", IMG_CFG)])
def test_unet_oracle_matches_reference(name, cfg):
    g = golden("unet_small")
    seed, x, t = unet_small_input(name)
    torch.manual_seed(seed)
    net = O.UNetOracle(**cfg).eval()
    assert sum(p.numel() for p in net.parameters()) == int(g[f"{name}_nparams"])
    assert abs(sum(p.double().sum().item() for p in net.parameters()) - float(g[f"{name}_wsum"])) < 1e-6
    y = net(x, torch.full((1,), t, dtype=torch.long)).numpy()
    assert rel_l2(y, g[f"{name}_y"]) < 1e-5


def test_guided_process_oracle_matches_reference_proj():
    g = golden("grp_small")
    torch.manual_seed(0)
    net = O.UNetOracle(**PROJ_CFG).eval()
    x = small_proj_input(200)
    res = O.guided_reverse_process(net, O.Tables(1000, 5), x, [15, 15, 15], clip=False, lambda_ratio=1, eta=0.5, mode="proj",
                                   constant_guidance=None, noise=iter(noise_tape(x.shape, 48, 300)), kernel_size=4, amplitude=7.0)
    got = np.stack([r.numpy()[0, 0] for r in res])
    assert got.shape == g["proj0"].shape
    assert rel_l2(got, g["proj0"]) < 2e-5


def test_guided_process_oracle_matches_reference_img():
    g = golden("grp_small")
    torch.manual_seed(1)
    net = O.UNetOracle(**IMG_CFG).eval()
    x = small_img_input(400)
    tape = noise_tape(x.shape, 66, 500)
    tab = O.Tables(1000, 1)
    res = O.guided_reverse_process(net, tab, x, [15, 15, 15], clip=True, lambda_ratio=10, eta=0.7, mode="img",
                                   constant_guidance=0.45, noise=iter(tape[:48]), ldct=x)
    res += O.guided_reverse_process(net, tab, res[-1], [5, 5, 5], clip=True, lambda_ratio=10, eta=0.6, mode="img",
                                    constant_guidance=0.6, noise=iter(tape[48:]), ldct=x)
    got = np.stack([r.numpy()[0, 0] for r in res])
    assert rel_l2(got, g["img0"]) < 2e-5


def test_fbp_oracle_tables_and_image_match_reference():
    import ipdm_pytorch_b200.synthetic as S
    g = golden("fbp_slice0")
    T = fbp_oracle.Tables()
    np.testing.assert_array_equal(T.h, g["h_RL"])
    np.testing.assert_array_equal(T.nda, g["nda"])
    np.testing.assert_array_equal(T.theta, g["theta"])
    np.testing.assert_array_equal(T.r.reshape(512, 512)[1::4, 2::4], g["r_sub"])
    np.testing.assert_array_equal(T.phi.reshape(512, 512)[1::4, 2::4], g["phi_sub"])
    assert float(g["simens_theta_maxabs_rad"]) < 1e-6          # Simens_theta.txt == arange(0, 360, .18) deg
    ld, nd, img = S.make_slice(0)
    rec = fbp_oracle.convert(ld[None])[0]
    assert rel_l2(rec, g["ld"]) < 1e-5                         # measured 2.8e-6 (libm / summation-order differences)
    # domain property: FBP of the clean synthetic sinogram reproduces the phantom inside the body
    assert float(g["nd_rmse_in_body"]) < 0.006


def test_sharpen_and_conversions():
    x = small_img_input(7)
    y = O.tensor_sharpen(x, 42)
    k = torch.full((3, 3), -2.0); k[1, 1] = 42
    ref = torch.nn.functional.conv2d(x, (k / 26)[None, None], padding=1)
    assert torch.allclose(y, ref)
    assert O.tensor_sharpen(x, -1) is x
    mu = torch.tensor([0.0, 0.183, 0.5, 1.0])
    pix = O.miu2pixel(mu)
    assert pix[0] == 0 and abs(float(pix[1]) - (1024 - 24) / 4096) < 1e-6 and pix[3] == 1


# ---- N1: image-quality metrics (published skimage definitions; skimage itself is not installable here) --------------------
def test_metrics_oracle_matches_the_definition_and_its_properties():
    from oracle import metrics_oracle as M
    rng = np.random.default_rng(0)
    ref = rng.random((40, 37)).astype(np.float32)
    test = (ref + 0.05 * rng.standard_normal(ref.shape)).astype(np.float32)
    assert M.ssim(ref, ref) == pytest.approx(1.0, abs=1e-12)
    assert M.psnr(ref, test) == pytest.approx(10 * np.log10(1.0 / np.mean((ref.astype(np.float64) - test) ** 2)), abs=1e-12)
    # brute-force SSIM straight from the formula (explicit 11x11 windows over the cropped interior)
    win, n = 11, 121
    acc = []
    x, y = test.astype(np.float64), ref.astype(np.float64)
    for i in range(5, 40 - 5):
        for j in range(5, 37 - 5):
            a, b = x[i - 5:i + 6, j - 5:j + 6], y[i - 5:i + 6, j - 5:j + 6]
            ux, uy = a.mean(), b.mean()
            vx, vy, vxy = ((a - ux) ** 2).sum() / (n - 1), ((b - uy) ** 2).sum() / (n - 1), ((a - ux) * (b - uy)).sum() / (n - 1)
            acc.append((2 * ux * uy + 1e-4) * (2 * vxy + 9e-4) / ((ux * ux + uy * uy + 1e-4) * (vx + vy + 9e-4)))
    assert M.ssim(ref, test, win_size=win) == pytest.approx(np.mean(acc), abs=1e-10)
    # NaNs of the test image count as 0.5 (metric_calculate :792)
    t2 = test.copy(); t2[3, 4] = np.nan
    t3 = test.copy(); t3[3, 4] = 0.5
    assert M.psnr(ref, t2) == M.psnr(ref, t3) and M.ssim(ref, t2) == M.ssim(ref, t3)
    # unit conversion agrees with the host mirror of Dataset/npz_data_loader.py
    from Dataset.npz_data_loader import miu2pixel
    mu = (0.25 * rng.random((8, 9))).astype(np.float32)
    np.testing.assert_array_equal(M.miu2pixel(mu), miu2pixel(mu.copy()))


# ---- N3: sparse (DDIM) guided sampler ------------------------------------------------------------------------------------
def test_sparse_sampler_oracle_matches_reference():
    g = golden("sparse_small")
    torch.manual_seed(0)
    pnet = O.UNetOracle(**PROJ_CFG).eval()
    x = small_proj_input(200)
    res = O.sparse_guided_reverse_process(pnet, O.Tables(1000, 5), x, [15, 15, 5], 0.49, 0.35, [1, 2, 3], eta=0.5, clip=False,
                                          noise=noise_tape(x.shape, 7, 700))
    assert rel_l2(np.stack([r.numpy()[0, 0] for r in res]), g["proj0"]) < 2e-5
    torch.manual_seed(1)
    inet = O.UNetOracle(**IMG_CFG).eval()
    x = small_img_input(400)
    res = O.sparse_guided_reverse_process(inet, O.Tables(1000, 1), x, [18, 18, 5], 0.5, 0.3, [1, 2, 3], eta=0.7, clip=True,
                                          noise=noise_tape(x.shape, 7, 800))
    assert rel_l2(np.stack([r.numpy()[0, 0] for r in res]), g["img0"]) < 2e-5


# ---- N4: image-domain adaptive lambda (constant_guidance_img=None) -----------------------------------------------------------
def test_img_adaptive_lambda_oracle_matches_reference():
    g = golden("img_adaptive_small")
    torch.manual_seed(1)
    net = O.UNetOracle(**IMG_CFG).eval()
    x = small_img_input(400)
    res = O.guided_reverse_process(net, O.Tables(1000, 1), x, [10, 9, 8], clip=True, lambda_ratio=10, eta=0.7, mode="img",
                                   constant_guidance=None, noise=iter(noise_tape(x.shape, 30, 900)), kernel_size=4, amplitude=20.0, ldct=x)
    assert rel_l2(np.stack([r.numpy()[0, 0] for r in res]), g["img0"]) < 2e-5


# ---- N3 (second half): adaptive t_start=None schedule selection ------------------------------------------------------------------
def test_adaptive_schedule_oracle_matches_reference():
    """t_start=None in both domains (reference Model/model.py:531-535, 582-613, 639-640): golden from the unmodified reference
    (oracle/make_golden.py adaptive), three projection slices that land in three classes, three image slices with three noise_strength values."""
    g = golden("adaptive_schedule_small")
    torch.manual_seed(0)
    pnet = O.UNetOracle(**PROJ_CFG).eval()
    for sid, amp in ((0, 7.0), (1, 3.0), (2, 15.0)):
        x = small_proj_input(200 + sid)
        info = {}
        res = O.guided_reverse_process(pnet, O.Tables(1000, 5), x, None, clip=False, lambda_ratio=1, eta=0.5, mode="proj", constant_guidance=None,
                                       noise=iter(noise_tape(x.shape, 99, 1100 + sid)), kernel_size=4, amplitude=amp, info=info)
        assert info["class"] == str(g[f"proj{sid}_class"]) and len(res) == 4
        assert rel_l2(np.stack([r.numpy()[0, 0] for r in res]), g[f"proj{sid}"]) < 2e-5, sid
    assert {str(g[f"proj{s}_class"]) for s in range(3)} == {"low", "mid", "high"}
    torch.manual_seed(1)
    inet = O.UNetOracle(**IMG_CFG).eval()
    for sid, ns in ((0, "mid"), (1, None), (2, "high")):
        x = small_img_input(400 + sid)
        res = O.guided_reverse_process(inet, O.Tables(1000, 1), x, None, clip=True, lambda_ratio=10, eta=0.7, mode="img", constant_guidance=None,
                                       noise=iter(noise_tape(x.shape, 69, 1200 + sid)), kernel_size=4, amplitude=20.0, ldct=x, noise_strength=ns)
        assert len(res) == 4 and rel_l2(np.stack([r.numpy()[0, 0] for r in res]), g[f"img{sid}"]) < 2e-5, sid


def test_ssim_oracle_matches_the_skimage_code_path_on_scipy():
    """skimage is absent, scipy is not.  `ssim_skimage_path` replays skimage 0.19's structural_similarity on the
    scipy.ndimage.uniform_filter it calls; the cumulative-sum oracle (what the device metric is tested against) must agree with it,
    for the reference's arguments (win_size=11, data_range=1) on a CT-like image pair and on noise, and for an even/odd size mix."""
    from oracle import metrics_oracle as M
    import ipdm_pytorch_b200.synthetic as S
    nd = M.miu2pixel(S.rasterize(S.phantom_ellipses(0)))
    rng = np.random.default_rng(3)
    pairs = [(nd, np.clip(nd + 0.02 * rng.standard_normal(nd.shape).astype(np.float32), 0, 1)),
             (rng.random((97, 130)).astype(np.float32), rng.random((97, 130)).astype(np.float32))]
    for ref, test in pairs:
        for win in (11, 7):
            a, b = M.ssim(ref, test, win_size=win), M.ssim_skimage_path(ref, test, win_size=win)
            assert a == pytest.approx(b, abs=1e-12), (win, a, b)
