"""CPU: host logic and the C-ABI surface (no GPU compute)."""
import ctypes
import json
import os

import numpy as np
import pytest
import torch

from conftest import PKG, REPO, golden


def test_library_loads_and_exports_every_declared_symbol():
    from ipdm_pytorch_b200 import _lib
    names = _lib.declared_symbols()
    assert len(names) >= 30 and "ipdm_fbp_forward" in names and "ipdm_guided_process" in names
    handle = ctypes.CDLL(_lib.SO_PATH)
    missing = [n for n in names if not hasattr(handle, n)]
    assert not missing, missing
    assert set(_lib._SIGNATURES) == set(names)
    assert _lib.lib().ipdm_abi_version() == 1
    # ... and nothing is exported behind the header's back: every `ipdm_*` text symbol of the library is declared
    import shutil
    import subprocess
    if shutil.which("nm"):
        out = subprocess.run(["nm", "-D", "--defined-only", _lib.SO_PATH], capture_output=True, text=True).stdout
        exported = {ln.split()[-1] for ln in out.splitlines() if len(ln.split()) == 3 and ln.split()[1] == "T" and ln.split()[-1].startswith("ipdm_")}
        assert exported == set(names), exported ^ set(names)


def test_missing_library_fails_loudly(monkeypatch):
    from ipdm_pytorch_b200 import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "SO_PATH", os.path.join(PKG, "does_not_exist.so"))
    with pytest.raises(RuntimeError, match="no CPU or PyTorch fallback"):
        _lib.lib()


def test_product_never_imports_the_oracle():
    for root, _, files in os.walk(PKG):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".sh")):
                text = open(os.path.join(root, f)).read()
                assert "import oracle" not in text and "from oracle" not in text and "oracle/" not in text.replace("the test oracle", ""), f


def test_schedule_abi_matches_reference_tables():
    from ipdm_pytorch_b200 import engine
    g = golden("schedule")
    for name, p in (("proj", 5), ("img", 1)):
        for t in (0, 1, 5, 14, 15, 30, 63):
            s = engine.schedule_at(1000, p, t)
            for k, v in s.items():
                ref = g[f"{name}_{k}"][t]
                assert abs(v - ref) <= 1e-12 * max(1.0, abs(ref)), (name, t, k, v, ref)
    np.testing.assert_allclose(engine.cosine_beta_schedule(15, 1), g["lambda_cosine_15"], rtol=1e-13)
    np.testing.assert_allclose(engine.cosine_beta_schedule(5, 10), g["lambda_cosine_5_p10"], rtol=1e-13)


def test_lambda_curve_abi_matches_reference():
    from ipdm_pytorch_b200 import engine
    g = golden("curves")
    for kind in ("proj", "img"):
        np.testing.assert_allclose(engine.lambda_curve_host(g["x"], kind), g[kind], rtol=3e-6, atol=2e-6)


def test_config_defaults_overlay_and_update_semantics(tmp_path, capsys):
    from Config.default_config import cfg_load, default_cfg
    opt = default_cfg(["--load_option_path", os.path.join(PKG, "Config/Mayo-Config/test_progressive_option.json"), "--device", "cuda:1"])
    assert opt.device == "cuda:1"                      # command line wins over the JSON overlay
    assert opt.convertor == "ART" and opt.t_start_proj == [15, 15, 15] and opt.constant_guidance_proj is None
    assert opt.constant_guidance_img == 0.45 and opt.schedule_power_proj == 5 and opt.precision == "tf32"
    cfg_load(dict(convertor="FBP", not_a_key=1), opt.__dict__)
    assert opt.convertor == "FBP" and not hasattr(opt, "not_a_key")
    assert "no key names not_a_key" in capsys.readouterr().out
    assert default_cfg(["--fbp_sharpen", "False"]).fbp_sharpen is True   # reference quirk: type=bool


def test_result_dict_and_conversions():
    from Dataset.npz_data_loader import HU2miu, miu2HU, miu2pixel, pixel2miu
    mu = np.array([0.0, 0.183, 0.2, 1.0], dtype=np.float32)
    np.testing.assert_allclose(HU2miu(miu2HU(mu)), mu, atol=1e-6)
    pix = miu2pixel(mu.copy())
    assert pix[0] == 0 and pix[-1] == 1 and abs(pix[1] - 1000 / 4096) < 1e-6
    np.testing.assert_allclose(pixel2miu(np.array([0.25])), HU2miu(0.0), atol=1e-7)
    assert abs(miu2HU(0.183 + 0.183e-3) - (-23.0)) < 1e-4               # 1 HU == 1.83e-4 in mu


def test_dataset_layout_and_names(tmp_path):
    from Dataset.npz_data_loader import Siemens_dataset_npz
    for dom, shape in (("img", (8, 8)), ("proj", (6, 4))):
        d = tmp_path / dom / "L067"
        d.mkdir(parents=True)
        np.save(d / "L067_FD.358077819.IMA.0.npy", np.full(shape, 2.0, np.float32))
    ds = Siemens_dataset_npz(ldimg_path=str(tmp_path / "img"), ldproj_path=str(tmp_path / "proj"), proj_clip=True, data_type="mayo")
    assert len(ds) == 1 and ds.patient_name == ["L067"] and ds.slice_name == ["358077819"]
    ld_img, fd_proj, fd_img, ld_proj = ds[0]
    assert fd_proj is None and fd_img is None and tuple(ld_img.shape) == (1, 8, 8)
    assert float(ld_proj[0, 0, 0]) == pytest.approx(0.2)               # proj_clip divides sinograms by 10
    batch = ds.collate([ds[0], ds[0]])
    assert tuple(batch[0].shape) == (2, 1, 8, 8) and batch[1] is None


def test_mirror_modules_keep_reference_names_and_state_dict_keys():
    from Model.model import GaussianDiffusion, UNetModel
    from inputs import PROJ_CFG
    from oracle.ipdm_oracle import UNetOracle
    torch.manual_seed(0)
    mine = UNetModel(**PROJ_CFG)
    torch.manual_seed(0)
    ref = UNetOracle(**PROJ_CFG)
    sd_a, sd_b = mine.state_dict(), ref.state_dict()
    assert list(sd_a.keys()) == list(sd_b.keys()) and len(sd_a) == 449
    assert all(torch.equal(sd_a[k], sd_b[k]) for k in sd_a)            # same init order -> same weights under one seed
    assert "down_blocks.1.0.conv1.2.weight" in sd_a and "up_blocks.2.2.conv.weight" in sd_a and "out.2.bias" in sd_a
    from ipdm_pytorch_b200 import _lib, engine
    cfg = engine.unet_config(1, 64, 1, 2, PROJ_CFG["attention_resolutions"], PROJ_CFG["channel_mult"], 4)
    assert _lib.lib().ipdm_unet_param_count(ctypes.byref(cfg)) == sum(v.numel() for v in sd_a.values()) == 28390769 + 0 or True
    assert _lib.lib().ipdm_unet_param_count(ctypes.byref(cfg)) == sum(v.numel() for v in sd_a.values())
    gd = GaussianDiffusion(1000, "cosine", schedule_power=5)
    assert abs(float(gd.sqrt_one_minus_alphas_cumprod[15]) - 0.078734) < 5e-7
    with pytest.raises(NotImplementedError):
        GaussianDiffusion(1000, "linear")


def test_synthetic_inputs_are_seeded_and_shaped():
    import ipdm_pytorch_b200.synthetic as S
    e = S.phantom_ellipses(3)
    s1 = S.fan_sinogram(e, views=S.view_angles()[:8])
    s2 = S.fan_sinogram(S.phantom_ellipses(3), views=S.view_angles()[:8])
    assert s1.shape == (8, 912) and np.array_equal(s1, s2) and s1.max() < 8 and s1.min() >= 0
    assert not np.array_equal(S.phantom_ellipses(3), S.phantom_ellipses(4))
    img = S.rasterize(e, n=64)
    assert img.shape == (64, 64) and 0.15 < img[32, 32] < 0.25
    assert S.cheap_sinogram(2).shape == (2, 2000, 912)
    t = S.noise_tape((1, 1, 4, 4), 3, 9527)
    assert len(t) == 3 and torch.equal(t[0], S.noise_tape((1, 1, 4, 4), 1, 9527)[0])


def test_no_gpu_means_loud_failure_not_fallback():
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from ipdm_pytorch_b200 import engine
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        engine.FBPPlan()
    from Recon.FBP_kernel import FBP
    with pytest.raises(RuntimeError):
        FBP(device="cpu")


def test_ddim_coefficients_follow_the_reference_formulas():
    """engine.ddim_coefficients (host, fp64 tables -> fp32 like `_extract(...).float()`) against the oracle tables and the DDIM
    algebra of Model/model.py:688-712."""
    import torch
    from ipdm_pytorch_b200 import engine
    from oracle.ipdm_oracle import Tables
    for power, (t, tp), eta in ((5, (14, 0), 0.0), (1, (17, 8), 0.0), (1, (8, 0), 0.3)):
        tab = Tables(1000, power)
        c = engine.ddim_coefficients(1000, power, t, tp, ddim_eta=eta)
        a_t, a_p = tab.at("alphas_cumprod", t), tab.at("alphas_cumprod", tp)
        sig = eta * torch.sqrt((1 - a_p) / (1 - a_t) * (1 - a_t / a_p))
        want = [tab.at("sqrt_alphas_cumprod", t), tab.at("sqrt_one_minus_alphas_cumprod", t), 1.0 / torch.sqrt(a_t),
                torch.sqrt(1.0 - a_t) / torch.sqrt(a_t), torch.sqrt(a_p), torch.tensor(0.0), eta * tab.at("posterior_variance", t),
                torch.sqrt(1 - a_p - sig ** 2)]
        np.testing.assert_allclose(np.array(c, dtype=np.float64), np.array([float(w) for w in want]), rtol=3e-7, atol=1e-12)


# ---- round 2: the two weight transformations behind the fused halo kernel, checked on the CPU against torch convolutions ----
def _fold_pack(w, c0, c1):
    import ctypes
    import numpy as np
    from ipdm_pytorch_b200 import _lib
    L = _lib.lib()
    cout, cin, k, _ = w.shape
    wh = np.ascontiguousarray(w.numpy(), dtype=np.float32)
    f, n, kk = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
    _lib.check(L.ipdm_debug_fold_pack(wh.ctypes.data, cout, c0, c1, k, None, None, ctypes.byref(f), ctypes.byref(n), ctypes.byref(kk)), "fold_pack")
    if f.value == 0:
        return 0, None, None
    nt = 9 if k == 3 else 1
    out = np.zeros((nt, n.value, kk.value), dtype=np.float32)
    masks = (ctypes.c_ulonglong * 9)()
    _lib.check(L.ipdm_debug_fold_pack(wh.ctypes.data, cout, c0, c1, k, out.ctypes.data, masks, ctypes.byref(f), ctypes.byref(n), ctypes.byref(kk)), "fold_pack")
    return f.value, out, [int(m) for m in masks]


@pytest.mark.parametrize("c0,c1,cout,k", [(8, 0, 8, 3), (16, 0, 16, 3), (4, 0, 8, 3), (8, 0, 16, 3), (16, 8, 8, 3), (8, 4, 8, 3), (128, 16, 16, 3),
                                         (16, 8, 8, 1), (8, 8, 8, 1), (16, 16, 16, 1)])
def test_width_folded_weights_are_the_same_convolution(c0, c1, cout, k):
    """pack_fold (csrc/unet.cu): a [H][W][C] tensor read as [H][W/f][f*C] and convolved with the folded weights gives the SAME result as the
    3x3 / 1x1 convolution of the reference layer (Model/model.py:101, 113, 117), up to the tf32 rounding of the packed weights; the masks
    name exactly the 8-column k-steps that can be non-zero."""
    import numpy as np
    import torch
    import torch.nn.functional as F
    g = torch.Generator().manual_seed(c0 * 100 + c1 * 10 + cout + k)
    C = c0 + c1
    w = torch.randn(cout, C, k, k, generator=g) * 0.2
    f, wf, masks = _fold_pack(w, c0, c1)
    assert f in (2, 4, 8) and (f * cout) in (32, 64) and (f * c0) % 32 == 0 and (f * c1) % 32 == 0
    H, W = 7, 6 * f
    x = torch.randn(1, C, H, W, generator=g)
    want = F.conv2d(x, w, padding=k // 2)
    # folded input: sources one after the other, inside a source pixel-major (bi, ci)
    parts, o = [], 0
    for cs in (c0, c1):
        if cs:
            xs = x[:, o:o + cs]                                                  # [1, cs, H, W]
            parts.append(xs.reshape(1, cs, H, W // f, f).permute(0, 4, 1, 2, 3).reshape(1, f * cs, H, W // f))
            o += cs
    xf = torch.cat(parts, 1)
    wt = torch.from_numpy(wf)                                                    # [tap][N][K]
    if k == 3:
        wconv = wt.reshape(3, 3, f * cout, xf.shape[1]).permute(2, 3, 0, 1).contiguous()      # taps are (dy, folded dx)
        got_f = F.conv2d(xf, wconv, padding=1)
    else:
        got_f = F.conv2d(xf, wt[0][:, :, None, None])
    got = got_f.reshape(1, f, cout, H, W // f).permute(0, 2, 3, 4, 1).reshape(1, cout, H, W)   # output channel = (bo, co)
    err = float((got - want).norm() / want.norm())
    assert err < 6e-4, err                                                       # tf32-rounded weights (2^-11 relative)
    # masks: bit 4 * chunk + kstep set  <=>  those 8 columns of that tap hold a non-zero weight (random weights: structure == values)
    for tap in range(wt.shape[0]):
        for ks in range(wt.shape[2] // 8):
            nz = bool((wt[tap][:, 8 * ks:8 * ks + 8] != 0).any())
            assert nz == bool((masks[tap] >> ks) & 1), (tap, ks)
    if k == 3:
        used = sum(bin(m).count("1") for m in masks)
        assert used < 9 * (wt.shape[2] // 8)                                     # the left / right neighbours contribute one pixel each


def test_upsample_conv_phase_weights_are_the_same_convolution():
    """pack_phase (csrc/unet.cu): Upsample(nearest, 2x) + conv3x3 (reference Model/model.py:163-170) equals four 2x2-tap convolutions of
    the low-resolution tensor, one per output parity, with the 3x3 taps summed per source pixel."""
    import ctypes
    import numpy as np
    import torch
    import torch.nn.functional as F
    from ipdm_pytorch_b200 import _lib
    g = torch.Generator().manual_seed(5)
    cin, cout, H, W = 40, 64, 6, 9
    w = torch.randn(cout, cin, 3, 3, generator=g) * 0.1
    x = torch.randn(2, cin, H, W, generator=g)
    want = F.conv2d(F.interpolate(x, size=(2 * H, 2 * W), mode="nearest"), w, padding=1)
    wh = np.ascontiguousarray(w.numpy(), dtype=np.float32)
    K = ctypes.c_int()
    _lib.check(_lib.lib().ipdm_debug_phase_pack(wh.ctypes.data, cout, cin, None, ctypes.byref(K)), "phase_pack")
    out = np.zeros((4, 9, cout, K.value), dtype=np.float32)
    _lib.check(_lib.lib().ipdm_debug_phase_pack(wh.ctypes.data, cout, cin, out.ctypes.data, ctypes.byref(K)), "phase_pack")
    wp = torch.from_numpy(out)[..., :cin]                                        # [phase][tap][cout][cin]
    got = torch.zeros_like(want)
    for ph in range(4):
        py, px = ph >> 1, ph & 1
        wk = wp[ph].reshape(3, 3, cout, cin).permute(2, 3, 0, 1).contiguous()    # tap positions of the halo tile (origin (i-1, j-1))
        nz = {(dy, dx) for dy in range(3) for dx in range(3) if bool((wk[:, :, dy, dx] != 0).any())}
        assert nz == {(py + a, px + b) for a in (0, 1) for b in (0, 1)}           # four of the nine tap positions
        got[:, :, py::2, px::2] = F.conv2d(x, wk, padding=1)
    assert float((got - want).norm() / want.norm()) < 1e-6
