"""GPU parity: guided partial reverse process (ipdm_guided_process) vs goldens from the unmodified reference."""
import numpy as np
import pytest
import torch

from conftest import golden, rel_l2
from inputs import IMG_CFG, PROJ_CFG, noise_tape, small_img_input, small_proj_input

pytestmark = pytest.mark.gpu
GRP_TOL_FP32 = 3e-4  # precision="fp32" (3xTF32)
GRP_TOL = 6e-3      # tf32 mode: rel-L2 of every recorded iterate after 45 (proj) / 60 (img) UNet calls (1e-3 per forward, re-fed)


def _tape(shape, count, seed, dev):
    return torch.stack(noise_tape(shape, count, seed)).to(dev).contiguous()


@pytest.mark.parametrize("prec,tol", [("tf32", GRP_TOL), ("fp32", GRP_TOL_FP32)])
def test_proj_domain_adaptive_lambda_two_slices(cuda, prec, tol):
    from Model.model import GaussianDiffusion, UNetModel
    g = golden("grp_small")
    torch.manual_seed(0)
    net = UNetModel(**PROJ_CFG).to(cuda).eval()
    net.set_precision(prec)
    gd = GaussianDiffusion(1000, "cosine", schedule_power=5)
    xs = [small_proj_input(200 + s) for s in (0, 1)]
    tapes = [_tape(xs[0].shape, 48, 300 + s, cuda) for s in (0, 1)]
    kw = dict(t_start=[15, 15, 15], clip=False, lambda_ratio=1, eta=0.5, mode="proj", constant_guidance=None,
              kernel_size_proj=4, amplitude_proj=7, only_convertor=False, normal=False)
    # batch of two different slices == the reference run per slice at B = 1
    x = torch.cat(xs).to(cuda)
    noise = torch.cat(tapes, dim=1).contiguous()                        # [48, 2, 1, H, W]
    res, _, ns = gd.guided_reverse_process(net, x, noise=noise, **kw)
    assert ns is None and len(res) == 4
    for s in (0, 1):
        got = np.stack([r[s, 0].cpu().numpy() for r in res])
        err = [rel_l2(got[k], g[f"proj{s}"][k]) for k in range(4)]
        print(f"proj slice {s} ({prec}): rel-L2 per iterate {['%.2e' % e for e in err]}")
        assert max(err) < tol
    one, _, _ = gd.guided_reverse_process(net, xs[1].to(cuda), noise=tapes[1], **kw)
    assert rel_l2(one[-1].cpu().numpy(), res[-1][1:2].cpu().numpy()) < 1e-5


@pytest.mark.parametrize("prec,tol", [("tf32", GRP_TOL), ("fp32", GRP_TOL_FP32), ("bf16", 5e-2)])
def test_img_domain_constant_guidance_and_ultra(cuda, prec, tol):
    from Model.model import GaussianDiffusion, UNetModel
    g = golden("grp_small")
    torch.manual_seed(1)
    net = UNetModel(**IMG_CFG).to(cuda).eval()
    net.set_precision(prec)
    gd = GaussianDiffusion(1000, "cosine", schedule_power=1)
    for s in (0, 1):
        x = small_img_input(400 + s).to(cuda)
        tape = _tape(x.shape, 66, 500 + s, cuda)
        common = dict(model=net, clip=True, lambda_ratio=10, mode="img", ldct=x, kernel_size_img=4, amplitude_img=30,
                      only_convertor=False, normal=False)
        res, _, _ = gd.guided_reverse_process(img=x, t_start=[15, 15, 15], eta=0.7, constant_guidance=0.45, noise=tape[:48].contiguous(), **common)
        res2, _, _ = gd.guided_reverse_process(img=res[-1], t_start=[5, 5, 5], eta=0.6, constant_guidance=0.6, noise=tape[48:].contiguous(), **common)
        got = np.stack([r[0, 0].cpu().numpy() for r in res + res2])
        err = [rel_l2(got[k], g[f"img{s}"][k]) for k in range(8)]
        print(f"img slice {s} ({prec}): rel-L2 per iterate {['%.2e' % e for e in err]}")
        assert max(err) < tol
        assert got.min() >= 0 and got.max() <= 1                       # clip_img clamps every iterate to [0,1]


def test_philox_path_is_deterministic_and_unsupported_modes_fail_loudly(cuda):
    from Model.model import GaussianDiffusion, UNetModel
    torch.manual_seed(1)
    net = UNetModel(**IMG_CFG).to(cuda).eval()
    gd = GaussianDiffusion(1000, "cosine", schedule_power=1)
    x = small_img_input(3, n=32).to(cuda)
    kw = dict(model=net, img=x, t_start=[3, 2], clip=True, eta=0.7, mode="img", constant_guidance=0.45, ldct=x, only_convertor=False, normal=False)
    a, _, _ = gd.guided_reverse_process(seed=7, **kw)
    b, _, _ = gd.guided_reverse_process(seed=7, **kw)
    c, _, _ = gd.guided_reverse_process(seed=8, **kw)
    assert len(a) == 3 and torch.equal(a[-1], b[-1]) and not torch.equal(a[-1], c[-1])
    assert torch.allclose(a[2], (a[0] + a[1]) / 2)
    with pytest.raises(ValueError, match="empty list"):                 # t_start=None + constant guidance: the reference returns []
        gd.guided_reverse_process(net, x, t_start=None, mode="img", constant_guidance=0.45, ldct=x, only_convertor=False)
    with pytest.raises(RuntimeError):                                   # img mode blends with ldct: it must be given
        gd.guided_reverse_process(net, x, t_start=[2], mode="img", constant_guidance=0.45, ldct=None, only_convertor=False)
    out, _, _ = gd.guided_reverse_process(net, x, t_start=[2], mode="img", constant_guidance=0.45, ldct=x, only_convertor=True)
    assert out[0] is x


# ---- sparse (DDIM) guided sampler, notebook cell 3 -----------------------------------------------------------
@pytest.mark.parametrize("prec,tol", [("tf32", GRP_TOL), ("fp32", GRP_TOL_FP32)])
def test_sparse_sampler_matches_reference_golden(cuda, prec, tol):
    """sparse_guided_reverse_process with the arguments proj_denoiser / img_denoiser pass for sample_method='sparse'
    (reference train_test_utils.py:445-453, :505-514); goldens from the unmodified reference (make_golden.py sparse)."""
    from Model.model import GaussianDiffusion, UNetModel
    g = golden("sparse_small")
    torch.manual_seed(0)
    pnet = UNetModel(**PROJ_CFG).to(cuda).eval()
    pnet.set_precision(prec)
    pgd = GaussianDiffusion(1000, "cosine", schedule_power=5)
    xs = [small_proj_input(200 + s) for s in (0, 1)]
    noise = torch.cat([_tape(xs[0].shape, 7, 700 + s, cuda) for s in (0, 1)], dim=1).contiguous()      # two slices in one batch
    res = pgd.sparse_guided_reverse_process(model=pnet, condition=torch.cat(xs).to(cuda), t_start=[15, 15, 5], condition_lambda_max=0.49,
                                            condition_lambda_min=0.35, clip_denoised=False, ddim_timesteps=[1, 2, 3], eta=0.5, noise=noise)
    assert len(res) == 3
    for s in (0, 1):
        err = [rel_l2(res[k][s, 0].cpu().numpy(), g[f"proj{s}"][k]) for k in range(3)]
        print(f"sparse proj slice {s} ({prec}): rel-L2 per iterate {['%.2e' % e for e in err]}")
        assert max(err) < tol
    torch.manual_seed(1)
    inet = UNetModel(**IMG_CFG).to(cuda).eval()
    inet.set_precision(prec)
    igd = GaussianDiffusion(1000, "cosine", schedule_power=1)
    for s in (0, 1):
        x = small_img_input(400 + s).to(cuda)
        res = igd.sparse_guided_reverse_process(model=inet, condition=x, t_start=[18, 18, 5], condition_lambda_max=0.5, condition_lambda_min=0.3,
                                                clip_denoised=True, ddim_timesteps=[1, 2, 3], eta=0.7, noise=_tape(x.shape, 7, 800 + s, cuda))
        err = [rel_l2(res[k][0, 0].cpu().numpy(), g[f"img{s}"][k]) for k in range(3)]
        print(f"sparse img slice {s} ({prec}): rel-L2 per iterate {['%.2e' % e for e in err]}")
        assert max(err) < tol


def test_notebook_cell3_sparse_progressive_runs(cuda, tmp_path):
    """test_sample.ipynb cell 3: update_opt(sample_method_proj='sparse', sample_method_img='sparse', ddim_timesteps=[1,2,3]) then
    progressive_denoiser(); Philox noise, two slices, result layout as in the dense mode."""
    import ipdm_pytorch_b200.synthetic as S
    from test_progressive_gpu import _model
    model = _model(tmp_path, dict(t_start_proj=[15, 15, 5], sample_method_proj="sparse", ddim_timesteps_proj=[1, 2, 3], t_start_img=[18, 18, 5],
                                  ddim_timesteps_img=[1, 2, 3], sample_method_img="sparse", noise_seed=3))
    ld = torch.from_numpy(np.stack([S.make_slice(s)[0] for s in (0, 1)]))[:, None]
    model.data_sample_load(ldct=None, ldproj=ld, fdproj=None, fdct=None)
    out = model.progressive_denoiser(sharpen_num=70, save_proj_state=True)
    assert out.shape == (2, 1, 512, 512) and torch.isfinite(out).all()
    assert len(model.proj_denoise_result) == 3                            # one entry per sparse proj iteration


# ---- image-domain adaptive lambda (constant_guidance_img=None, the argparse default) ---------------------------
@pytest.mark.parametrize("prec,tol", [("tf32", GRP_TOL), ("fp32", GRP_TOL_FP32)])
def test_img_domain_adaptive_lambda_matches_reference_golden(cuda, prec, tol):
    from Model.model import GaussianDiffusion, UNetModel
    g = golden("img_adaptive_small")
    torch.manual_seed(1)
    net = UNetModel(**IMG_CFG).to(cuda).eval()
    net.set_precision(prec)
    gd = GaussianDiffusion(1000, "cosine", schedule_power=1)
    for s in (0, 1):
        x = small_img_input(400 + s).to(cuda)
        res, _, _ = gd.guided_reverse_process(model=net, img=x, t_start=[10, 9, 8], clip=True, lambda_ratio=10, eta=0.7, mode="img",
                                              constant_guidance=None, ldct=x, kernel_size_img=4, amplitude_img=20, only_convertor=False,
                                              normal=False, noise=_tape(x.shape, 30, 900 + s, cuda))
        err = [rel_l2(res[k][0, 0].cpu().numpy(), g[f"img{s}"][k]) for k in range(4)]
        print(f"img adaptive slice {s} ({prec}): rel-L2 per iterate {['%.2e' % e for e in err]}")
        assert max(err) < tol


# ---- SURVEY N3 (second half): adaptive t_start=None schedule selection -----------------------------------------------------
@pytest.mark.parametrize("prec,tol", [("tf32", 1.2e-2), ("fp32", 5e-4)])
def test_adaptive_schedule_proj_matches_reference_golden(cuda, prec, tol):
    """t_start=None in the projection domain (reference Model/model.py:531-535, 596-613, 639-640): probing iteration with t_start = 20,
    max(exp(amplitude * delta-map)) reduced on the device (ipdm_delta_exp_max), ONE host read, then the [30,25,20] / [20,18,15] /
    [15,15,15] continuation.  Goldens from the unmodified reference (make_golden.py adaptive): amplitude 7 / 3 / 15 land in mid / low / high."""
    from Model.model import GaussianDiffusion, UNetModel
    g = golden("adaptive_schedule_small")
    torch.manual_seed(0)
    net = UNetModel(**PROJ_CFG).to(cuda).eval()
    net.set_precision(prec)
    gd = GaussianDiffusion(1000, "cosine", schedule_power=5)
    seen = set()
    for sid, amp in ((0, 7), (1, 3), (2, 15)):
        x = small_proj_input(200 + sid).to(cuda)
        tape = _tape(x.shape, 99, 1100 + sid, cuda)
        res, _, ns = gd.guided_reverse_process(net, x, t_start=None, clip=False, lambda_ratio=1, eta=0.5, mode="proj", constant_guidance=None,
                                               kernel_size_proj=4, amplitude_proj=amp, only_convertor=False, normal=False, noise=tape)
        assert ns == str(g[f"proj{sid}_class"]) and len(res) == 4
        seen.add(ns)
        err = [rel_l2(res[k][0, 0].cpu().numpy(), g[f"proj{sid}"][k]) for k in range(4)]
        print(f"adaptive proj slice {sid} ({prec}, amplitude {amp}): class {ns}, rel-L2 per iterate {['%.2e' % e for e in err]}")
        assert max(err) < tol
    assert seen == {"low", "mid", "high"}
    # a batch of three slices: per-slice decision, one continuation per class, results scattered back in slice order
    xs = torch.cat([small_proj_input(200 + s) for s in range(3)]).to(cuda)
    tapes = torch.cat([_tape(xs[:1].shape, 99, 1100 + s, cuda) for s in range(3)], dim=1).contiguous()
    res, _, ns = gd.guided_reverse_process(net, xs, t_start=None, clip=False, lambda_ratio=1, eta=0.5, mode="proj", constant_guidance=None,
                                           kernel_size_proj=4, amplitude_proj=7, only_convertor=False, normal=False, noise=tapes)
    assert ns == "mid" and tuple(res[-1].shape) == tuple(xs.shape)
    assert rel_l2(res[-1][0, 0].cpu().numpy(), g["proj0"][3]) < tol


@pytest.mark.parametrize("prec,tol", [("tf32", 1.2e-2), ("fp32", 5e-4)])
def test_adaptive_schedule_img_per_slice_classes(cuda, prec, tol):
    """t_start=None in the image domain (reference :582-594): the schedule follows the noise_strength reported by the projection stage.
    Three slices with three different classes run as ONE call (per-slice list): three continuations, scattered back."""
    from Model.model import GaussianDiffusion, UNetModel
    g = golden("adaptive_schedule_small")
    torch.manual_seed(1)
    net = UNetModel(**IMG_CFG).to(cuda).eval()
    net.set_precision(prec)
    gd = GaussianDiffusion(1000, "cosine", schedule_power=1)
    kw = dict(t_start=None, clip=True, lambda_ratio=10, eta=0.7, mode="img", constant_guidance=None, kernel_size_img=4, amplitude_img=20,
              only_convertor=False, normal=False)
    classes = ["mid", None, "high"]
    xs = torch.cat([small_img_input(400 + s) for s in range(3)]).to(cuda)
    tapes = torch.cat([_tape(xs[:1].shape, 69, 1200 + s, cuda) for s in range(3)], dim=1).contiguous()
    res, _, _ = gd.guided_reverse_process(net, xs, ldct=xs, noise_strength=classes, noise=tapes, **kw)
    assert len(res) == 4
    for s in range(3):
        err = [rel_l2(res[k][s, 0].cpu().numpy(), g[f"img{s}"][k]) for k in range(4)]
        print(f"adaptive img slice {s} ({prec}, noise_strength {classes[s]}): rel-L2 per iterate {['%.2e' % e for e in err]}")
        assert max(err) < tol
    one, _, _ = gd.guided_reverse_process(net, xs[2:3].contiguous(), ldct=xs[2:3].contiguous(), noise_strength="high", noise=tapes[:, 2:3].contiguous(), **kw)
    assert rel_l2(one[-1].cpu().numpy(), res[-1][2:3].cpu().numpy()) < 1e-5


def test_adaptive_schedule_through_the_drop_in_api(cuda, tmp_path):
    """update_opt(dict(t_start_proj=None, t_start_img=None)) (the argparse defaults, train_img_option.json): the projection stage picks the
    schedule, reports noise_strength, the image stage follows it (constant_guidance_img=None), Philox noise."""
    import os
    import ipdm_pytorch_b200.synthetic as S
    from conftest import PKG
    from Config.default_config import default_cfg
    from Utils.train_test_utils import progressive_domain_denoiser
    opt = default_cfg(["--load_option_path", os.path.join(PKG, "Config/Mayo-Config/test_progressive_option.json"), "--device", "cuda:0"])
    opt.load_img_model_path = opt.load_proj_model_path = None
    opt.test_dataset_path_FD_img = opt.test_dataset_path_LD_img = opt.test_dataset_path_FD_proj = opt.test_dataset_path_LD_proj = None
    torch.manual_seed(0)
    model = progressive_domain_denoiser(opt, result_save_path=str(tmp_path))
    model.update_opt(dict(convertor="FBP", ultra_img_denoise=False, constant_guidance_img=None, noise_seed=3, precision="bf16",
                          t_start_proj=None, t_start_img=None))
    ld = np.stack([S.make_slice(s)[0] for s in (0, 1)])
    model.data_sample_load(ldct=None, ldproj=torch.from_numpy(ld)[:, None], fdproj=None, fdct=None)
    out = model.progressive_denoiser()
    assert tuple(out.shape) == (2, 1, 512, 512) and torch.isfinite(out).all()
    ns = model.noise_strength
    assert (ns in ("low", "mid", "high")) or (isinstance(ns, list) and all(c in ("low", "mid", "high") for c in ns))
