"""Teacher-forced per-step parity of the guided reverse process at the BASELINE sizes (2000x912 and 512x512).

With random-init weights the UNet is not a denoiser: a per-forward difference is re-fed 45 + 60 times with gain > 1, so an
end-to-end comparison measures the amplification, not the kernels (tests/test_progressive_gpu.py reports that number next to
the reference's own GPU-vs-CPU distance).  Here every one of the 45 + 60 reverse steps is checked in isolation: the CUDA path
runs its own trajectory; at each step the oracle (oracle/ipdm_oracle.py, pinned to the unmodified reference) is handed the
CUDA path's x_t, guidance, lambda and noise, runs its own UNet forward and its own p_sample_condition (reference
Model/model.py:492-515) and must reproduce eps and x_{t-1}.  The oracle runs in torch fp32 on the GPU (TF32 disabled); a few
steps are repeated on the host CPU to tie that to the CPU oracle the goldens were made with.
"""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from conftest import golden

pytestmark = pytest.mark.gpu

# per-step bounds (max over the steps of a stage), stated per precision mode: (eps rel-L2, x_{t-1} rel-L2).
# Measured on B200 (round 2): eps fp32 1.0e-4 / 4.3e-5 (proj / img), tf32 4.9e-3 / 3.3e-3, bf16 4.3e-2 / 2.3e-2;
# x_{t-1} fp32 3.6e-7 / 4.3e-6, tf32 2.0e-5 / 3.3e-4, bf16 1.9e-4 / 2.2e-3 (the image-domain update is a larger share of x).
BOUNDS = {
    ("proj", "fp32"): (2e-4, 1e-5), ("img", "fp32"): (2e-4, 1e-5),
    ("proj", "tf32"): (6e-3, 6e-4), ("img", "tf32"): (6e-3, 6e-4),
    ("proj", "bf16"): (6e-2, 4e-3), ("img", "bf16"): (3e-2, 4e-3),
}


def _rel(a, b):
    return float((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30))


class _NoTF32:
    def __enter__(self):
        self.old = (torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32)
        torch.backends.cuda.matmul.allow_tf32 = torch.backends.cudnn.allow_tf32 = False

    def __exit__(self, *exc):
        torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = self.old


def _pair(kind, prec, cuda):
    """(CUDA UNet, oracle UNet on the GPU, oracle UNet on the CPU) with the weights of the golden runs (seed 0, proj then img)."""
    from inputs import IMG_CFG, PROJ_CFG
    from Model.model import UNetModel
    from oracle import ipdm_oracle as O
    torch.manual_seed(0)
    proj = UNetModel(**PROJ_CFG)
    img = UNetModel(**IMG_CFG)
    net, cfg = (proj, PROJ_CFG) if kind == "proj" else (img, IMG_CFG)
    ora_cpu = O.UNetOracle(**cfg).eval()
    ora_cpu.load_state_dict(net.state_dict())
    ora_gpu = O.UNetOracle(**cfg).eval()
    ora_gpu.load_state_dict(net.state_dict())
    net = net.to(cuda).eval()
    net.set_precision(prec)
    return net, ora_gpu.to(cuda), ora_cpu


def _stage_args(kind):
    if kind == "proj":
        return dict(t_start=[15, 15, 15], clip=False, lambda_ratio=1, eta=0.5, mode="proj", constant_guidance=None, schedule_power=5)
    return dict(t_start=[15, 15, 15], clip=True, lambda_ratio=10, eta=0.7, mode="img", constant_guidance=0.45, schedule_power=1)


@pytest.mark.parametrize("kind", ["proj", "img"])
def test_stepwise_twin_equals_the_one_call_process(cuda, kind):
    """The Python twin (tests/stepping.py) issues the same ABI calls as ipdm_guided_process: same iterates."""
    from inputs import small_img_input, small_proj_input
    from ipdm_pytorch_b200 import engine
    from stepping import guided_process_stepwise
    net, _, _ = _pair(kind, "tf32", cuda)
    x = (small_proj_input(200) if kind == "proj" else small_img_input(210)).to(cuda)
    a = _stage_args(kind)
    a["t_start"] = [3, 2, 2]
    g = torch.Generator().manual_seed(5)
    tape = torch.randn(10, *x.shape, generator=g).to(cuda)
    sp = a.pop("schedule_power")
    p = engine.guided_params(a["mode"], list(a["t_start"]), a["clip"], a["lambda_ratio"], a["eta"], a["constant_guidance"], 4, 7.0, sp)
    want = engine.guided_process(net.cuda_handle(), p, x, x if kind == "img" else None, tape)
    got = guided_process_stepwise(net.cuda_handle(), x, noise=tape, schedule_power=sp, ldct=x, **a)
    assert len(got) == want.shape[0]
    for k in range(len(got)):
        assert _rel(got[k], want[k]) < 1e-6, k


@pytest.mark.parametrize("prec", ["fp32", "tf32", "bf16"])
@pytest.mark.parametrize("kind", ["proj", "img"])
def test_teacher_forced_every_step_full_size(cuda, kind, prec):
    import ipdm_pytorch_b200.synthetic as S
    from inputs import noise_tape
    from oracle import ipdm_oracle as O
    from stepping import guided_process_stepwise
    net, ora, ora_cpu = _pair(kind, prec, cuda)
    a = _stage_args(kind)
    sp = a.pop("schedule_power")
    tab = O.Tables(1000, sp)
    if kind == "proj":
        x = torch.from_numpy(S.make_slice(0)[0])[None, None].to(cuda)
        tape = torch.stack(noise_tape((1, 1, 2000, 912), 48, 9527)).to(cuda)
    else:
        x = torch.from_numpy(golden("img_stage512")["x"])[None, None].to(cuda)          # the reference's own sharpened FBP image
        tape = torch.stack(noise_tape((1, 1, 512, 512), 66, 19527)).to(cuda)
    cpu_steps = {(0, 14), (1, 7), (2, 0)}
    rows, cpu_rows = [], []

    def check(s):
        t = torch.full((1,), s["i"], dtype=torch.long)
        lam = s["lam"]
        if isinstance(lam, torch.Tensor):
            lam = F.interpolate(lam[:, None], size=s["x_t"].shape[-2:], mode="nearest")
        with _NoTF32():
            eps_ref = ora(s["x_t"], t)
            x_ref = O.p_sample_condition(tab, eps_ref, s["x_t"], s["guide"], s["i"], lam, a["clip"], s["noise"])
            x_mix = O.p_sample_condition(tab, s["eps"], s["x_t"], s["guide"], s["i"], lam, a["clip"], s["noise"])   # oracle step on OUR eps
        upd = (x_ref - s["x_t"]).double().norm().clamp_min(1e-30)
        rows.append((s["it"], s["i"], _rel(s["eps"], eps_ref), _rel(s["x_next"], x_ref), float((s["x_next"] - x_ref).double().norm() / upd),
                     _rel(s["x_next"], x_mix)))
        if (s["it"], s["i"]) in cpu_steps:
            eps_cpu = ora_cpu(s["x_t"].cpu(), t)
            cpu_rows.append((s["it"], s["i"], _rel(eps_ref.cpu(), eps_cpu), _rel(s["eps"].cpu(), eps_cpu)))

    res = guided_process_stepwise(net.cuda_handle(), x, noise=tape[:48], schedule_power=sp, ldct=x, on_step=check, **a)
    if kind == "img":                                          # the "ultra" pass (reference train_test_utils.py:515-536): 15 more steps
        a = dict(a, t_start=[5, 5, 5], eta=0.6, constant_guidance=0.6)
        cpu_steps = {(1, 2)}
        guided_process_stepwise(net.cuda_handle(), res[-1], noise=tape[48:], schedule_power=sp, ldct=x, on_step=check, **a)
    r = np.array(rows)
    assert len(rows) == (45 if kind == "proj" else 60)
    print(f"teacher-forced {kind} stage ({prec}), {len(rows)} steps: eps rel-L2 max {r[:, 2].max():.2e} (median {np.median(r[:, 2]):.2e}); "
          f"x_(t-1) rel-L2 max {r[:, 3].max():.2e}; error / size of the update max {r[:, 4].max():.2e}; "
          f"sampler step alone (oracle step on our eps) max {r[:, 5].max():.2e}")
    for it, i, eg, ec in cpu_rows:
        print(f"    step (it {it}, t {i}): GPU-fp32 oracle vs CPU oracle eps {eg:.2e}; CUDA path vs CPU oracle eps {ec:.2e}")
    e_eps, e_x = BOUNDS[(kind, prec)]
    assert r[:, 2].max() <= e_eps, r[:, 2].max()
    assert r[:, 3].max() <= e_x, r[:, 3].max()
    assert r[:, 5].max() <= 5e-6                              # the fused reduce -> apply kernel itself, every precision mode
    assert max(c[2] for c in cpu_rows) <= 5e-5                # the GPU-hosted oracle is the CPU oracle to fp32 summation order
    assert max(c[3] for c in cpu_rows) <= 1.25 * e_eps
