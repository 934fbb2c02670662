"""TEST INFRASTRUCTURE ONLY -- drives the UNMODIFIED reference (/root/reference) on CPU.

Used by oracle/make_golden.py in the build container to (a) pin oracle/ipdm_oracle.py
and oracle/fbp_oracle.c against the reference itself and (b) write the committed
fixtures under tests/golden/.  /root/reference does not exist on the GPU box, so
nothing reachable from `pytest -m gpu`, smoke() or bench.py imports this module.

Recipe: stub the reporting-only imports the container
lacks (matplotlib, skimage, piq, the Windows .pyd), chdir into the reference so
its relative `Recon/Simens_*.txt` reads resolve, replace the numba.cuda lambda
kernel (Model/model.py:328-351, no CPU branch) by a numpy restatement with the
same fp64 arithmetic, and replace torch.randn_like by a pre-generated noise tape.
No reference file is edited.
"""
import math
import os
import sys
import types

import numpy as np

REF_ROOT = os.environ.get("IPDM_REFERENCE_ROOT", "/root/reference")


class _LambdaKernelOnHost:
    """Stand-in for `condition_lambda_ratio_cuda[grid, block](I, idx, B, H, W, ts, lam)`."""

    def __getitem__(self, _launch_cfg):
        def launch(I, idx, B, H, W, timesteps, lam):
            s = 0.008
            f = [math.cos(((float(i) / timesteps) + s) / (1 + s) * math.pi * 0.5) ** 2 for i in idx]
            lam64 = lam.astype(np.float64)
            a0, a1, a2 = f[0] ** lam64, f[1] ** lam64, f[2] ** lam64
            I[...] = (1 - ((a2 / a0) / (a1 / a0))).astype(I.dtype)
        return launch


def install_stubs():
    def mod(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
        return m

    for name in ("matplotlib", "skimage", "piq"):
        try:
            __import__(name)
        except Exception:
            mod(name)
    if "matplotlib.pyplot" not in sys.modules:
        try:
            __import__("matplotlib.pyplot")
        except Exception:
            sys.modules["matplotlib"].pyplot = mod("matplotlib.pyplot")
    try:
        __import__("skimage.metrics")
    except Exception:
        mod("skimage.metrics", structural_similarity=lambda *a, **k: 0.0,
            peak_signal_noise_ratio=lambda *a, **k: 0.0)
    if not hasattr(sys.modules["piq"], "vif_p"):
        sys.modules["piq"].vif_p = lambda *a, **k: None
        sys.modules["piq"].fsim = lambda *a, **k: None

    def _no_art(*a, **k):
        raise RuntimeError("ART convertor (TASART2DNSL0.pyd) is out of scope")
    mod("Recon.TASART2DNSL0", recons_torch=_no_art, proj_torch=_no_art)


def load_reference():
    """Returns the reference's (Model.model, Recon.FBP_kernel, Utils.train_test_utils, Config.default_config)."""
    if not os.path.isdir(REF_ROOT):
        raise RuntimeError(f"reference not present at {REF_ROOT}")
    install_stubs()
    os.chdir(REF_ROOT)
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    sys.dont_write_bytecode = True
    import Model.model as MM
    import Recon.FBP_kernel as RF
    MM.condition_lambda_ratio_cuda = _LambdaKernelOnHost()
    import Utils.train_test_utils as TT
    import Config.default_config as CFG
    return MM, RF, TT, CFG


class NoiseTape:
    """Monkeypatch for torch.randn_like that pops pre-generated tensors in call order."""

    def __init__(self, tensors):
        self.tensors = list(tensors)
        self.used = 0

    def __call__(self, like, *a, **k):
        t = self.tensors[self.used]
        self.used += 1
        assert t.shape == like.shape, (t.shape, like.shape)
        return t.to(like.dtype)

    def __enter__(self):
        import torch
        self._orig = torch.randn_like
        torch.randn_like = self
        return self

    def __exit__(self, *exc):
        import torch
        torch.randn_like = self._orig


def build_denoiser(tmp_dir, seed=0, extra_opt=None):
    """The reference progressive_domain_denoiser on CPU with seeded random-init weights."""
    import numba
    import torch
    MM, RF, TT, CFG = load_reference()
    numba.set_num_threads(1)           # fbp_cpu races across views otherwise
    opt = CFG.default_cfg(argv=["--load_option_path", "Config/Mayo-Config/test_progressive_option.json",
                                "--device", "cpu"])
    opt.load_img_model_path = None
    opt.load_proj_model_path = None
    torch.manual_seed(seed)
    model = TT.progressive_domain_denoiser(opt, result_save_path=tmp_dir)
    cfg = dict(convertor="FBP", save_it_state_img=False, save_it_state_proj=False, ultra_img_denoise=True)
    cfg.update(extra_opt or {})
    model.update_opt(cfg)
    model.proj_model.eval()
    model.img_model.eval()
    return model
