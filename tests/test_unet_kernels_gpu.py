"""GPU parity, one UNet building block at a time (ipdm_debug_* entry points) vs plain PyTorch fp32 on the CPU.

Tolerances: CUDA-core kernels are fp32 (1e-5 rel-L2); tcgen05 kind::tf32 contractions carry a 10-bit
mantissa on both operands (<= 2e-3 rel-L2 for these K, the error the reference itself has on an
Ampere+ GPU with its default allow_tf32=True)."""
import ctypes

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from conftest import rel_l2

pytestmark = pytest.mark.gpu
TF32_TOL = 2e-3
FP32_TOL = 1e-5
SPLIT_TOL = 2e-5        # tcgen05 3xTF32 ("fp32" precision mode): dropped lo*lo terms ~2^-22, tensor-core fp32 accumulation


def tf32_rn(x):
    """Round to nearest tf32 (ties away from zero, like cvt.rna.tf32.f32)."""
    b = x.contiguous().view(torch.int32)
    return ((b + 0x1000) & ~0x1FFF).view(torch.float32)


def nhwc(x, cs=None):
    n, c, h, w = x.shape
    cs = c if cs is None else cs
    out = torch.zeros(n, h, w, cs, dtype=torch.float32)
    out[..., :c] = x.permute(0, 2, 3, 1)
    return out.contiguous()


def nchw(y, c):
    return y[..., :c].permute(0, 3, 1, 2).contiguous()


def alloc_cs(c):
    return (c + 31) // 32 * 32 if c >= 16 else c


def _p(t):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def run_conv(cuda, x0, x1, weight, bias, k, stride, use_tc, res=None, up=None, norm=None, dense_out=False, dense_src=False):
    from ipdm_pytorch_b200 import _lib
    use_tc = int(use_tc)
    variant, use_tc = use_tc >> 8, use_tc & 0xFF                       # bits 8..: tensor-core kernel variant (0 auto)
    fused = norm is not None and use_tc in (1, 3)                      # GroupNorm-fused conv: raw fp32 sources whatever the operand type
    if use_tc == 3 and not fused:                                       # bf16 operand tensor: one source, stride multiple of 64
        x0 = x0 if x1 is None else torch.cat([x0, x1], 1)
        x1 = None
    n, c0, h, w = x0.shape
    c1 = 0 if x1 is None else x1.shape[1]
    cs0, cs1 = (alloc_cs(c0), alloc_cs(c1)) if (use_tc and not dense_src) else (c0, c1)
    if use_tc == 3 and not fused:
        cs0 = (c0 + 63) // 64 * 64
    cout = weight.shape[0]
    hin, win = (h, w) if up is None else up
    ho, wo = (hin, win) if stride == 1 else ((hin + 1) // 2, (win + 1) // 2)
    ocs = cout if dense_out else alloc_cs(cout)                      # the planner stores <= 16-channel activations dense
    a0 = nhwc(x0, cs0).to(cuda)
    if use_tc == 3 and not fused:
        a0 = a0.to(torch.bfloat16).contiguous()
    a1 = None if x1 is None else nhwc(x1, cs1).to(cuda)
    r = None if res is None else nhwc(res, ocs).to(cuda)
    out = torch.full((n, ho, wo, ocs), float("nan"), device=cuda)
    wh = weight.contiguous().float()
    bh = None if bias is None else bias.contiguous().float()
    sc = sh = None
    if norm is not None:
        sc, sh = [t.to(cuda).contiguous() for t in norm]
    rc = _lib.lib().ipdm_debug_conv(_p(a0), c0, cs0, _p(a1), c1, cs1, n, h, w, _p(wh), _p(bh), cout, k, stride,
                                    0 if up is None else up[0], 0 if up is None else up[1], _p(sc), _p(sh), _p(r), ocs, _p(out), ocs,
                                    int(use_tc) | (variant << 8), None)
    _lib.check(rc, "ipdm_debug_conv")
    torch.cuda.synchronize()
    full = out.cpu()
    if ocs > cout:
        assert float(full[..., cout:].abs().max()) == 0.0              # channel padding stays zero
    return nchw(full, cout)


def ref_conv(x0, x1, weight, bias, k, stride, res=None, up=None, norm=None):
    x = x0 if x1 is None else torch.cat([x0, x1], 1)
    if norm is not None:
        x = F.silu(x * norm[0][:, :, None, None] + norm[1][:, :, None, None])
    if up is not None:
        x = F.interpolate(x, size=up, mode="nearest")
    y = F.conv2d(x, weight, bias, stride=stride, padding=k // 2)
    return y if res is None else y + res


def rnd(*shape, seed=0, scale=1.0):
    return scale * torch.randn(*shape, generator=torch.Generator().manual_seed(seed))


# ---- direct (CUDA core) path: the thin full-resolution layers of the proj net --------------------
@pytest.mark.parametrize("c0,c1,cout,k,stride,hw", [
    (1, 0, 4, 3, 1, (40, 70)), (4, 0, 8, 3, 1, (33, 65)), (4, 0, 8, 1, 1, (33, 65)), (8, 0, 8, 3, 2, (25, 57)),
    (16, 8, 16, 3, 1, (21, 47)), (8, 4, 8, 1, 1, (21, 47)), (8, 0, 1, 3, 1, (40, 70)), (16, 0, 16, 3, 2, (50, 38)),
    (1, 0, 64, 3, 1, (32, 32)), (64, 0, 1, 3, 1, (24, 40)),
    (16, 8, 8, 1, 1, (37, 95)), (16, 16, 16, 1, 1, (21, 47)), (8, 0, 16, 1, 1, (64, 31)), (8, 0, 8, 3, 2, (64, 96)), (1, 0, 64, 3, 1, (47, 53)),
])
def test_direct_conv(cuda, c0, c1, cout, k, stride, hw):
    x0 = rnd(2, c0, *hw, seed=1)
    x1 = rnd(2, c1, *hw, seed=2) if c1 else None
    w = rnd(cout, c0 + c1, k, k, seed=3, scale=0.2)
    b = rnd(cout, seed=4)
    ho, wo = hw if stride == 1 else ((hw[0] + 1) // 2, (hw[1] + 1) // 2)
    res = rnd(2, cout, ho, wo, seed=5) if (stride == 1 and k == 3) else None
    got = run_conv(cuda, x0, x1, w, b, k, stride, False, res=res)
    assert rel_l2(got.numpy(), ref_conv(x0, x1, w, b, k, stride, res=res).numpy()) < FP32_TOL
    if cout % 4 == 0:                                                   # dense output rows: the fully unrolled streaming kernels
        res2 = rnd(2, cout, ho, wo, seed=6) if k == 1 else None
        got = run_conv(cuda, x0, x1, w, b, k, stride, False, res=res2, dense_out=True)
        assert rel_l2(got.numpy(), ref_conv(x0, x1, w, b, k, stride, res=res2).numpy()) < FP32_TOL


def test_direct_conv_fused_norm_and_upsample(cuda):
    x0, x1 = rnd(2, 16, 20, 30, seed=1), rnd(2, 8, 20, 30, seed=2)
    sc, sh = 1 + 0.3 * rnd(2, 24, seed=6), 0.2 * rnd(2, 24, seed=7)
    w, b = rnd(16, 24, 3, 3, seed=3, scale=0.2), rnd(16, seed=4)
    got = run_conv(cuda, x0, x1, w, b, 3, 1, False, norm=(sc, sh))
    assert rel_l2(got.numpy(), ref_conv(x0, x1, w, b, 3, 1, norm=(sc, sh)).numpy()) < 2e-5
    xs = rnd(2, 16, 13, 19, seed=8)                                    # 13x19 -> 25x38 (odd target, like 63 -> 125)
    w2 = rnd(16, 16, 3, 3, seed=9, scale=0.2)
    got = run_conv(cuda, xs, None, w2, b, 3, 1, False, up=(25, 38))
    assert rel_l2(got.numpy(), ref_conv(xs, None, w2, b, 3, 1, up=(25, 38)).numpy()) < FP32_TOL


# ---- GroupNorm ------------------------------------------------------------------------------------
@pytest.mark.parametrize("c0,c1,hw", [(8, 0, (40, 70)), (16, 8, (21, 47)), (128, 16, (13, 10)), (256, 128, (7, 5)), (64, 0, (32, 32)), (12, 0, (9, 11))])
def test_groupnorm_stats_and_apply(cuda, c0, c1, hw):
    from ipdm_pytorch_b200 import _lib
    from oracle.ipdm_oracle import gn_groups
    C = c0 + c1
    x0 = 0.5 + 2 * rnd(2, c0, *hw, seed=1)
    x1 = -1 + rnd(2, c1, *hw, seed=2) if c1 else None
    gamma, beta = 1 + 0.2 * rnd(C, seed=3), 0.1 * rnd(C, seed=4)
    a0 = nhwc(x0, alloc_cs(c0)).to(cuda)
    a1 = None if x1 is None else nhwc(x1, alloc_cs(c1)).to(cuda)
    ocs = (C + 31) // 32 * 32
    sc, sh = torch.empty(2, C, device=cuda), torch.empty(2, C, device=cuda)
    out = torch.full((2, hw[0], hw[1], ocs), float("nan"), device=cuda)
    for act in (1, 0):
        rc = _lib.lib().ipdm_debug_groupnorm(_p(a0), c0, alloc_cs(c0), _p(a1), c1, alloc_cs(c1), 2, hw[0], hw[1], _p(gamma.contiguous()),
                                             _p(beta.contiguous()), act, _p(sc), _p(sh), _p(out), ocs, None)
        _lib.check(rc, "ipdm_debug_groupnorm")
        x = x0 if x1 is None else torch.cat([x0, x1], 1)
        want = F.group_norm(x, gn_groups(C), gamma, beta, eps=1e-5)
        want = F.silu(want) if act else want
        full = out.cpu()
        assert rel_l2(nchw(full, C).numpy(), want.numpy()) < 4e-4     # the applied tensor is rounded to tf32 (MMA operand)
        assert ocs == C or float(full[..., C:].abs().max()) == 0.0
    lin = x * sc.cpu()[:, :, None, None] + sh.cpu()[:, :, None, None]
    assert rel_l2(lin.numpy(), F.group_norm(x, gn_groups(C), gamma, beta, eps=1e-5).numpy()) < 2e-5


# ---- tensor-core implicit GEMM ----------------------------------------------------------------------
@pytest.mark.parametrize("c0,c1,cout,k,stride,hw", [
    (128, 0, 128, 3, 1, (25, 19)),      # ragged tile edges
    (128, 0, 128, 3, 1, (16, 8)),       # exactly one 128-pixel tile
    (256, 0, 256, 3, 1, (7, 5)),        # two N tiles, tiny image
    (128, 128, 128, 3, 1, (13, 10)),    # virtual concat
    (128, 16, 128, 3, 1, (25, 19)),     # 144 -> 128 (skip padded to 32)
    (128, 16, 16, 3, 1, (25, 19)),      # 144 -> 16, N = 16 tile
    (16, 0, 128, 3, 1, (25, 19)),       # 16 -> 128, K padded to 32
    (128, 0, 128, 3, 2, (25, 19)),      # stride 2, odd size
    (256, 0, 256, 3, 2, (14, 10)),      # stride 2, even size
    (256, 0, 768, 1, 1, (7, 5)),        # qkv-like 1x1
    (256, 128, 256, 1, 1, (13, 10)),    # 1x1 shortcut over a concat
    (64, 0, 64, 3, 1, (32, 32)),        # img net, N = 64 tile
    (128, 64, 64, 3, 1, (32, 32)),
])
def test_tc_conv(cuda, c0, c1, cout, k, stride, hw):
    x0 = rnd(2, c0, *hw, seed=1)
    x1 = rnd(2, c1, *hw, seed=2) if c1 else None
    w = rnd(cout, c0 + c1, k, k, seed=3, scale=(1.0 / ((c0 + c1) * k * k)) ** 0.5)
    b = rnd(cout, seed=4)
    ho, wo = hw if stride == 1 else ((hw[0] + 1) // 2, (hw[1] + 1) // 2)
    res = rnd(2, cout, ho, wo, seed=5) if stride == 1 else None
    want = ref_conv(x0, x1, w, b, k, stride, res=res)
    for mode, tol in ((1, TF32_TOL), (2, SPLIT_TOL)):
        got = run_conv(cuda, x0, x1, w, b, k, stride, mode, res=res)
        assert torch.isfinite(got).all()
        err = rel_l2(got.numpy(), want.numpy())
        print(f"tc conv {c0}+{c1}->{cout} k{k} s{stride} {hw} mode {'tf32' if mode == 1 else '3xtf32'}: rel-L2 {err:.2e}")
        assert err < tol
    if stride == 1:                                                     # bf16 operands (GroupNorm-apply / upsample outputs only: stride 1)
        got = run_conv(cuda, x0, x1, w, b, k, stride, 3, res=res)
        xb = (x0 if x1 is None else torch.cat([x0, x1], 1)).to(torch.bfloat16).float()
        exact = ref_conv(xb, None, w.to(torch.bfloat16).float(), b, k, stride, res=res)
        e_fp32, e_bf16 = rel_l2(got.numpy(), want.numpy()), rel_l2(got.numpy(), exact.numpy())
        print(f"tc conv {c0}+{c1}->{cout} k{k} {hw} mode bf16: rel-L2 {e_fp32:.2e} vs fp32, {e_bf16:.2e} vs bf16-rounded operands")
        assert e_fp32 < 8e-3 and e_bf16 < 2e-5


@pytest.mark.parametrize("c0,c1,cout,k,stride,hw", [(128, 0, 128, 3, 1, (176, 228)), (64, 0, 64, 3, 1, (256, 160)), (128, 16, 16, 3, 1, (260, 300)),
                                                  (128, 0, 128, 3, 2, (353, 457)), (128, 128, 256, 3, 1, (203, 331)), (256, 0, 128, 3, 1, (500, 228))])
def test_tc_conv_halo_reuse_variant(cuda, c0, c1, cout, k, stride, hw):
    """Halo-reuse kernels (nine taps read one staged halo tile through shifted UMMA descriptors): variant 2 one tile per CTA,
    variant 4 persistent with double-buffered accumulator pairs; ragged in both directions, concat, N = 16/64/128, tf32 and bf16
    operands.  The stride-2 case checks that an ineligible layer falls back to the per-tap kernels."""
    x0 = rnd(2, c0, *hw, seed=1)
    x1 = rnd(2, c1, *hw, seed=2) if c1 else None
    w = rnd(cout, c0 + c1, k, k, seed=3, scale=(1.0 / ((c0 + c1) * k * k)) ** 0.5)
    b = rnd(cout, seed=4)
    res = rnd(2, cout, *hw, seed=5) if stride == 1 else None
    want = ref_conv(x0, x1, w, b, k, stride, res=res)
    for variant in (2, 4):
        got = run_conv(cuda, x0, x1, w, b, k, stride, 1 | (variant << 8), res=res)
        assert rel_l2(got.numpy(), want.numpy()) < TF32_TOL, variant
        if stride == 1:
            got = run_conv(cuda, x0, x1, w, b, k, stride, 3 | (variant << 8), res=res)
            assert rel_l2(got.numpy(), want.numpy()) < 8e-3, variant


@pytest.mark.parametrize("c0,c1,cout,hw", [(64, 0, 64, (176, 120)), (128, 0, 128, (99, 115)), (128, 64, 64, (96, 150)), (128, 16, 128, (130, 89)),
                                          (256, 0, 256, (104, 90)), (256, 256, 128, (120, 120))])
def test_tc_conv_fused_groupnorm(cuda, c0, c1, cout, hw):
    """conv_halo_fused_kernel: GroupNorm affine + SiLU applied on the operand path of the persistent halo conv (raw fp32 sources,
    virtual concat, zero padding applied AFTER the activation, ragged edges, padded skip channels, one or two N tiles).
    tf32 operands: within the tf32 tolerance of torch; bf16 operands: within the bf16 tolerance of torch and within 3e-4 of a torch conv
    on bf16-rounded operands (the device SiLU is approximate: a few operands round to the neighbouring bf16 value)."""
    from ipdm_pytorch_b200 import _lib
    n, C = 2, c0 + c1
    x0 = rnd(n, c0, *hw, seed=1) + 0.3
    x1 = rnd(n, c1, *hw, seed=2) - 0.2 if c1 else None
    w = rnd(cout, C, 3, 3, seed=3, scale=(1.0 / (C * 9)) ** 0.5)
    b = rnd(cout, seed=4)
    res = rnd(n, cout, *hw, seed=5)
    scale, shift = 0.5 + torch.rand(n, C, generator=torch.Generator().manual_seed(6)), rnd(n, C, seed=7, scale=0.5)
    want = ref_conv(x0, x1, w, b, 3, 1, res=res, norm=(scale, shift))
    got = run_conv(cuda, x0, x1, w, b, 3, 1, 1, res=res, norm=(scale, shift))
    e_tf32 = rel_l2(got.numpy(), want.numpy())
    assert e_tf32 < TF32_TOL, e_tf32
    # the unfused pair on the same scale / shift: apply pass (identity statistics: gamma = scale, beta = shift on pre-normalised input is not
    # available through the debug ABI, so apply in torch with the kernel's rounding) is covered by the whole-UNet tests; here the fused
    # kernel is compared with the persistent halo kernel fed an operand tensor computed with the SAME device arithmetic
    xs = x0 if x1 is None else torch.cat([x0, x1], 1)
    got_bf = run_conv(cuda, x0, x1, w, b, 3, 1, 3, res=res, norm=(scale, shift))
    e_bf16 = rel_l2(got_bf.numpy(), want.numpy())
    act = F.silu(xs * scale[:, :, None, None] + shift[:, :, None, None])
    exact = ref_conv(act.to(torch.bfloat16).float(), None, w.to(torch.bfloat16).float(), b, 3, 1, res=res)
    e_exact = rel_l2(got_bf.numpy(), exact.numpy())
    print(f"fused GroupNorm conv {c0}+{c1}->{cout} {hw}: tf32 {e_tf32:.2e}, bf16 {e_bf16:.2e} vs fp32, {e_exact:.2e} vs bf16-rounded operands")
    assert e_bf16 < 8e-3 and e_exact < 3e-4          # (the device SiLU uses ex2.approx: a few operands round to the neighbouring bf16)


@pytest.mark.parametrize("cin,cout,hw", [(128, 128, (120, 121)), (64, 64, (97, 150)), (128, 64, (128, 96)), (256, 256, (64, 150))])
def test_upsample_conv_phases(cuda, cin, cout, hw):
    """Upsample(nearest, exactly 2x) + conv3x3 as four output-parity 2x2-tap convs on the low-resolution tensor (pack_phase,
    ConvTcDesc::phase_up): summed 3x3 taps per phase, zero padding of the UPSAMPLED image at all four borders, ragged tiles,
    strided output pixels.  bf16 and tf32 operands, fp32 accumulate."""
    n = 2
    x = rnd(n, cin, *hw, seed=1)
    w = rnd(cout, cin, 3, 3, seed=3, scale=0.1)
    b = rnd(cout, seed=4)
    up = (2 * hw[0], 2 * hw[1])
    want = ref_conv(x, None, w, b, 3, 1, up=up)
    for mode, tol in ((7, 8e-3), (8, TF32_TOL)):
        got = run_conv(cuda, x, None, w, b, 3, 1, mode, up=up)
        assert rel_l2(got.numpy(), want.numpy()) < tol, mode


@pytest.mark.parametrize("c0,c1,cout,k,hw", [
    (8, 0, 8, 3, (40, 72)), (16, 0, 16, 3, (33, 66)), (4, 0, 8, 3, (20, 80)), (8, 0, 16, 3, (25, 64)), (16, 8, 8, 3, (21, 68)),
    (8, 8, 8, 3, (19, 36)), (8, 4, 8, 3, (18, 48)), (16, 16, 16, 3, (23, 34)), (16, 8, 16, 3, (17, 44)), (128, 16, 16, 3, (20, 62)),
    (8, 0, 8, 3, (300, 912)),
    (4, 0, 8, 1, (20, 80)), (8, 0, 16, 1, (25, 64)), (16, 8, 8, 1, (21, 68)), (8, 8, 8, 1, (19, 36)), (16, 16, 16, 1, (23, 34)), (128, 16, 16, 1, (20, 62)),
])
def test_width_folded_conv(cuda, c0, c1, cout, k, hw):
    """Thin layers as width-folded tensor-core layers (pack_fold): [H][W][C] read as [H][W/f][f*C], folded weights with structurally
    zero k-steps skipped, GroupNorm + SiLU fused on the operand path, per-channel affine / bias indexed modulo the real channel
    count, virtual concat, residual, ragged tile edges.  tf32 operands, fp32 accumulate."""
    n, C = 2, c0 + c1
    x0 = rnd(n, c0, *hw, seed=1)
    x1 = rnd(n, c1, *hw, seed=2) if c1 else None
    w = rnd(cout, C, k, k, seed=3, scale=0.2)
    b = rnd(cout, seed=4)
    res = rnd(n, cout, *hw, seed=5)
    scale, shift = 0.5 + torch.rand(n, C, generator=torch.Generator().manual_seed(6)), rnd(n, C, seed=7, scale=0.5)
    norms = (None, (scale, shift)) if k == 3 else (None,)           # the 1x1 layers are the shortcuts: no GroupNorm
    for norm in norms:
        for r in (None, res):
            want = ref_conv(x0, x1, w, b, k, 1, res=r, norm=norm)
            got = run_conv(cuda, x0, x1, w, b, k, 1, 6, res=r, norm=norm, dense_out=True, dense_src=True)
            assert rel_l2(got.numpy(), want.numpy()) < TF32_TOL, (norm is not None, r is not None)


@pytest.mark.parametrize("cin,cs,cout,k,hw,batch", [(8, 8, 8, 3, (40, 70), 2), (4, 8, 8, 3, (33, 65), 1), (8, 8, 16, 1, (21, 47), 2), (16, 16, 16, 3, (50, 38), 2),
                                                    (12, 16, 8, 3, (25, 61), 1), (24, 32, 8, 3, (37, 95), 2), (32, 32, 16, 3, (64, 64), 1),
                                                    (8, 8, 8, 3, (300, 400), 2)])
def test_thin_tc_conv(cuda, cin, cs, cout, k, hw, batch):
    """Thin layers on tcgen05 with 32 / 64 / 128-byte swizzled operand rows, nine taps from one staged halo tile."""
    from ipdm_pytorch_b200 import _lib
    x = rnd(batch, cin, *hw, seed=1)
    w = rnd(cout, cin, k, k, seed=3, scale=(1.0 / (cin * k * k)) ** 0.5)
    b = rnd(cout, seed=4)
    res = rnd(batch, cout, *hw, seed=5)
    ocs = alloc_cs(cout)
    a0 = nhwc(x, cs).to(cuda)
    r = nhwc(res, ocs).to(cuda)
    out = torch.full((batch, hw[0], hw[1], ocs), float("nan"), device=cuda)
    rc = _lib.lib().ipdm_debug_conv(_p(a0), cin, cs, None, 0, 0, batch, hw[0], hw[1], _p(w.contiguous()), _p(b.contiguous()), cout, k, 1, 0, 0,
                                    None, None, _p(r), ocs, _p(out), ocs, 4, None)
    _lib.check(rc, "ipdm_debug_conv(thin)")
    full = out.cpu()
    assert torch.isfinite(full).all()
    if ocs > cout:
        assert float(full[..., cout:].abs().max()) == 0.0
    err = rel_l2(nchw(full, cout).numpy(), ref_conv(x, None, w, b, k, 1, res=res).numpy())
    print(f"thin conv {cin}({cs})->{cout} k{k} {hw}: rel-L2 {err:.2e}")
    assert err < TF32_TOL


def test_tc_conv_full_size_row_shapes(cuda):
    """The proj net's real row widths (228, 114, 57, 29) at reduced height, batch 1."""
    for hw, c in (((24, 228), 128), ((20, 114), 128), ((16, 57), 256), ((9, 29), 256)):
        x = rnd(1, c, *hw, seed=hw[1])
        w = rnd(c, c, 3, 3, seed=3, scale=(1.0 / (9 * c)) ** 0.5)
        got = run_conv(cuda, x, None, w, None, 3, 1, True)
        assert rel_l2(got.numpy(), ref_conv(x, None, w, None, 3, 1).numpy()) < TF32_TOL, hw


# ---- attention ----------------------------------------------------------------------------------------
@pytest.mark.parametrize("hw,batch", [((7, 5), 2), ((13, 10), 1), ((25, 19), 2), ((32, 32), 1), ((63, 29), 1), ((125, 57), 1)])
def test_attention(cuda, hw, batch):
    from ipdm_pytorch_b200 import _lib
    C, heads, d = 256, 4, 64
    T = hw[0] * hw[1]
    tpad = (T + 3) // 4 * 4
    qkv = rnd(batch, 3 * C, *hw, seed=T)
    q, k, v = qkv.reshape(batch * heads, 3 * d, T).chunk(3, dim=1)
    sc = 1.0 / (d ** 0.25)
    att = torch.einsum("bct,bcs->bts", q * sc, k * sc).softmax(dim=-1)
    want = torch.einsum("bts,bcs->bct", att, v).reshape(batch, C, *hw)
    qk_nhwc = nhwc(qkv).to(cuda)                                        # [B,H,W,3C]; v channels are ignored by the kernel
    vt = torch.zeros(batch, heads, d, tpad)
    vt[..., :T] = v.reshape(batch, heads, d, T)
    vt = vt.to(cuda).contiguous()
    for split, tol in ((False, TF32_TOL), (True, SPLIT_TOL)):
        out = torch.full((batch, hw[0], hw[1], C), float("nan"), device=cuda)
        if split:
            qh, vh = tf32_rn(qk_nhwc), tf32_rn(vt)
            ql, vl = tf32_rn(qk_nhwc - qh), tf32_rn(vt - vh)
            rc = _lib.lib().ipdm_debug_attention(_p(qh), _p(vh), _p(ql), _p(vl), _p(out), batch, T, tpad, heads, C, None)
        else:
            rc = _lib.lib().ipdm_debug_attention(_p(qk_nhwc), _p(vt), None, None, _p(out), batch, T, tpad, heads, C, None)
        _lib.check(rc, "ipdm_debug_attention")
        torch.cuda.synchronize()
        got = nchw(out.cpu(), C)
        assert torch.isfinite(got).all()
        err = rel_l2(got.numpy(), want.numpy())
        print(f"attention T={T} {'3xtf32' if split else 'tf32'}: rel-L2 {err:.2e}")
        assert err < tol
    # bf16 operand mode (kind::f16): q, k, v^T handed over as bf16, t_pad a multiple of 8
    tpad8 = (T + 7) // 8 * 8
    vt8 = torch.zeros(batch, heads, d, tpad8)
    vt8[..., :T] = v.reshape(batch, heads, d, T)
    out = torch.full((batch, hw[0], hw[1], C), float("nan"), device=cuda)
    qk16, vt16 = qk_nhwc.to(torch.bfloat16).contiguous(), vt8.to(cuda).to(torch.bfloat16).contiguous()
    _lib.check(_lib.lib().ipdm_debug_attention_bf16(_p(qk16), _p(vt16), _p(out), batch, T, tpad8, heads, C, None), "ipdm_debug_attention_bf16")
    torch.cuda.synchronize()
    got = nchw(out.cpu(), C)
    assert torch.isfinite(got).all()
    err = rel_l2(got.numpy(), want.numpy())
    print(f"attention T={T} bf16: rel-L2 {err:.2e}")
    assert err < 2e-2


def test_upsample_nearest_index_rule(cuda):
    from ipdm_pytorch_b200 import _lib
    for (hs, ws), (hd, wd) in (((63, 29), (125, 57)), ((4, 3), (7, 5)), ((16, 16), (32, 32))):
        x = torch.round(8 * rnd(2, 32, hs, ws, seed=hs))              # small integers: exact in tf32 (the kernel rounds its output)
        src = nhwc(x).to(cuda)
        dst = torch.empty(2, hd, wd, 32, device=cuda)
        _lib.check(_lib.lib().ipdm_debug_upsample(_p(src), 2, hs, ws, 32, _p(dst), hd, wd, None), "ipdm_debug_upsample")
        assert torch.equal(nchw(dst.cpu(), 32), F.interpolate(x, size=(hd, wd), mode="nearest"))
