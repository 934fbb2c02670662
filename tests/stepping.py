"""Step-by-step twin of `ipdm_guided_process` (csrc/engine.cu) built from the single-op entry points of the C ABI.

The product enqueues the whole guided partial reverse process (reference Model/model.py:517-642) in one C call; the
teacher-forced parity tests need to look at every reverse step, so this helper issues the SAME sequence of ABI calls
(ipdm_q_sample, ipdm_lambda_step_map, ipdm_unet_forward, ipdm_sampler_step, ipdm_clamp, ipdm_delta_lambda_map[_img],
ipdm_lincomb) from Python and hands each step to a callback.  `test_stepwise_twin_equals_the_one_call_process` checks
that the two produce the same iterates, so what the callback sees is what the product computes.
"""
import ctypes
import math

import torch

from ipdm_pytorch_b200 import _lib, engine


def _delta_map_img(x, img, ks, amplitude):
    b, h, w = x.shape[0], x.shape[-2], x.shape[-1]
    out = torch.empty(b, h // ks, w // ks, device=x.device, dtype=torch.float32)
    pooled = torch.empty_like(out)
    ws = engine._workspace(_lib.lib().ipdm_sampler_workspace_bytes(b, h, w), x.device)
    engine.check(_lib.lib().ipdm_delta_lambda_map_img(engine._dev(x), engine._dev(img), engine._dev(out), None, engine._dev(pooled), b, h, w,
                                                      int(ks), float(amplitude), 1, engine._dev(ws), engine._stream()), "ipdm_delta_lambda_map_img")
    return out


def guided_process_stepwise(unet, img, t_start, clip, lambda_ratio, eta, mode, constant_guidance, noise, schedule_power,
                            ks=4, amplitude=7.0, ldct=None, timesteps=1000, on_step=None):
    """Returns the list of iterates (+ the mean of the last two), like `GaussianDiffusion.guided_reverse_process`.

    unet: engine.UNetHandle; img [B,1,H,W] CUDA f32; noise: tensor [count,B,1,H,W] in the reference's randn_like order.
    on_step(info) is called after every reverse step with a dict: it, i (timestep), x_t, guide, eps, lam (float or the
    [B,H/ks,W/ks] map), noise, x_next, coef (the 7 step coefficients).  The state continues from the CUDA result."""
    adaptive = constant_guidance is None
    x = img.clone()
    guide = img
    lam_exp = None
    iters, call = [], 0
    for it, ts in enumerate(t_start):
        tab = engine.schedule_at(timesteps, schedule_power, ts)
        x = engine.q_sample(x, tab["sqrt_alphas_cumprod"], tab["sqrt_one_minus_alphas_cumprod"], noise=noise[call].contiguous())
        call += 1
        lam_cos = engine.cosine_beta_schedule(ts, lambda_ratio) if (adaptive and it == 0) else None
        for i in range(ts - 1, -1, -1):
            if adaptive:
                lam = float(lam_cos[i]) if it == 0 else engine.lambda_step_map(lam_exp, i, ts)
            else:
                lam = float(constant_guidance)
            eps = unet.forward(x, i)
            s = engine.schedule_at(timesteps, schedule_power, i)
            f32 = lambda v: ctypes.c_float(v).value
            coef = [f32(s["sqrt_alphas_cumprod"]), f32(s["sqrt_one_minus_alphas_cumprod"]), f32(s["sqrt_recip_alphas_cumprod"]),
                    f32(s["sqrt_recipm1_alphas_cumprod"]), f32(s["posterior_mean_coef1"]), f32(s["posterior_mean_coef2"]),
                    f32(math.exp(0.5 * f32(s["posterior_log_variance_clipped"])))]
            nz = noise[call].contiguous()
            x_next = engine.sampler_step(x, guide, eps, coef, lam, noise=nz, clip=clip, t_nonzero=i != 0, ks=ks)
            call += 1
            if on_step is not None:
                on_step(dict(it=it, i=i, ts=ts, x_t=x, guide=guide, eps=eps, lam=lam, noise=nz, x_next=x_next, coef=coef))
            x = x_next
        if clip:
            engine.clamp_(x, 0.0, 1.0 if mode == "img" else math.inf)
        out_it = x.clone()
        iters.append(out_it)
        blend = (lambda: engine.lincomb(eta, out_it, 1 - eta, img)) if mode == "proj" else \
                (lambda: engine.lincomb(eta, out_it, 0.95 - eta, img, 0.05, ldct))
        if adaptive and it == 0:
            lam_exp = engine.delta_lambda_map(x, img, ks, amplitude, "proj") if mode == "proj" else _delta_map_img(x, img, ks, amplitude)
            x = img.clone()
        else:
            guide = blend()
    if len(iters) > 1:
        m = engine.lincomb(1.0, iters[-1], 1.0, iters[-2])
        iters.append(engine.lincomb(0.5, m, 0.0, m))
    return iters
