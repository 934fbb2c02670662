"""Host-side mirror of the reference's Model/model.py for the progressive inference path.

Same public names, constructor arguments, state_dict keys and (under a fixed torch seed) the same
random initialisation as the reference (Model/model.py: UNetModel :190-310, GaussianDiffusion
:376-642, cosine_beta_schedule :366-372), but every forward computation is executed by
libipdm_b200.so:
  * `UNetModel` only *holds* parameters (so checkpoints `proj_model-N` / `img_model-N` load
    unchanged); `forward` hands x and t to the tcgen05 UNet plan.
  * `GaussianDiffusion.guided_reverse_process` hands the whole iteration structure to
    `ipdm_guided_process` (one stream, no host round trips); the extra keyword `noise=` carries a
    caller-supplied tape [count,B,1,H,W] in the reference's randn_like order, `seed=` keys the
    in-kernel Philox generator otherwise.
  * `ddim_sample` / `sparse_guided_reverse_process` (:654-759): host loop over UNet forward + fused DDIM step.
Out of scope here (reference lines): train_losses :645-652, Yeo-Johnson :762-807.
"""
import math
from copy import copy

import numpy as np
import torch
import torch.nn as nn

from _ipdm_boot import engine as _eng
from Dataset.npz_data_loader import miu2pixel  # noqa: F401  (re-export, as the reference does)


def _group_count(channels):
    if channels % 32 == 0:
        return 32
    if channels < 32:
        return channels
    divisors = []
    for i in range(1, int(math.sqrt(channels)) + 1):
        if channels % i == 0:
            divisors.append(i)
            if channels // i != i:
                divisors.append(channels // i)
    divisors = np.array(divisors)
    return int(divisors[np.argmin((divisors - 32) ** 2)])


def norm_layer(channels):
    return nn.GroupNorm(_group_count(channels), channels)


def timestep_embedding(timesteps, dim, max_period=10000, dtype=torch.float32):
    half = dim // 2
    freqs = torch.exp(-math.log(max_period) * torch.arange(start=0, end=half) / half).type(dtype).to(timesteps.device)
    args = timesteps[:, None] * freqs[None]
    emb = torch.cat([torch.cos(args), torch.sin(args)], dim=-1)
    if dim % 2:
        emb = torch.cat([emb, torch.zeros_like(emb[:, :1])], dim=-1)
    return emb


class TimestepBlock(nn.Module):
    pass


class TimestepEmbedSequential(nn.Sequential, TimestepBlock):
    """Parameter container only; execution order lives in the CUDA plan."""


def _conv_norm_act(cin, cout):
    return nn.Sequential(norm_layer(cin), nn.SiLU(), nn.Conv2d(cin, cout, kernel_size=3, padding=1))


class ResidualBlock(TimestepBlock):
    def __init__(self, in_channels, out_channels, time_channels, dropout):
        super().__init__()
        self.conv1 = _conv_norm_act(in_channels, out_channels)
        self.time_emb = nn.Sequential(nn.SiLU(), nn.Linear(time_channels, out_channels))
        self.conv2 = _conv_norm_act(out_channels, out_channels)
        self.shortcut = nn.Conv2d(in_channels, out_channels, kernel_size=1) if in_channels != out_channels else nn.Identity()


class AttentionBlock(nn.Module):
    def __init__(self, channels, num_heads=1):
        super().__init__()
        assert channels % num_heads == 0
        self.num_heads = num_heads
        self.norm = norm_layer(channels)
        self.qkv = nn.Conv2d(channels, channels * 3, kernel_size=1, bias=False)
        self.proj = nn.Conv2d(channels, channels, kernel_size=1)


class Upsample(nn.Module):
    def __init__(self, channels, use_conv):
        super().__init__()
        if not use_conv:
            raise NotImplementedError("conv_resample=False is not used by any shipped config")
        self.use_conv = use_conv
        self.conv = nn.Conv2d(channels, channels, kernel_size=3, padding=1)


class Downsample(nn.Module):
    def __init__(self, channels, use_conv):
        super().__init__()
        if not use_conv:
            raise NotImplementedError("conv_resample=False is not used by any shipped config")
        self.use_conv = use_conv
        self.op = nn.Conv2d(channels, channels, kernel_size=3, stride=2, padding=1)


class UNetModel(nn.Module):
    def __init__(self, in_channels=3, model_channels=128, out_channels=3, num_res_blocks=2, attention_resolutions=(8, 16),
                 dropout=0, channel_mult=(1, 2, 2, 2), conv_resample=True, num_heads=4, pre_downsample_times=1):
        super().__init__()
        self.in_channels, self.model_channels, self.out_channels = in_channels, model_channels, out_channels
        self.num_res_blocks, self.attention_resolutions, self.dropout = num_res_blocks, attention_resolutions, dropout
        self.channel_mult, self.conv_resample, self.num_heads = channel_mult, conv_resample, num_heads
        self.precision = "tf32"
        self.max_t = 64                      # rows of the per-ResBlock bias + time-embedding table (t in [0, max_t))
        self._handle = None
        self._handle_key = None

        tdim = model_channels * 4
        self.time_embed = nn.Sequential(nn.Linear(model_channels, tdim), nn.SiLU(), nn.Linear(tdim, tdim))
        width = lambda m: int(m * model_channels)
        ch = width(channel_mult[0])
        self.down_blocks = nn.ModuleList([TimestepEmbedSequential(nn.Conv2d(in_channels, ch, kernel_size=3, padding=1))])
        skip_chans, ds, mults = [ch], 1, list(channel_mult[1:])
        for level, mult in enumerate(mults):
            for _ in range(num_res_blocks):
                parts = [ResidualBlock(ch, width(mult), tdim, dropout)]
                ch = width(mult)
                if ds in attention_resolutions:
                    parts.append(AttentionBlock(ch, num_heads=num_heads))
                self.down_blocks.append(TimestepEmbedSequential(*parts))
                skip_chans.append(ch)
            if level != len(mults) - 1:
                self.down_blocks.append(TimestepEmbedSequential(Downsample(ch, conv_resample)))
                skip_chans.append(ch)
                ds *= 2
        self.middle_block = TimestepEmbedSequential(ResidualBlock(ch, ch, tdim, dropout), AttentionBlock(ch, num_heads=num_heads),
                                                    ResidualBlock(ch, ch, tdim, dropout))
        self.up_blocks = nn.ModuleList([])
        for level, mult in list(enumerate(mults))[::-1]:
            for i in range(num_res_blocks + 1):
                parts = [ResidualBlock(ch + skip_chans.pop(), width(mult), tdim, dropout)]
                ch = width(mult)
                if ds in attention_resolutions:
                    parts.append(AttentionBlock(ch, num_heads=num_heads))
                if level and i == num_res_blocks:
                    parts.append(Upsample(ch, conv_resample))
                    ds //= 2
                self.up_blocks.append(TimestepEmbedSequential(*parts))
        self.out = _conv_norm_act(ch, out_channels)

    # -- CUDA plan ---------------------------------------------------------------------------
    def set_precision(self, precision):
        if precision != self.precision:
            self.precision, self._handle = precision, None

    def _weights_key(self):
        return (self.precision, self.max_t) + tuple((p.data_ptr(), p._version) for p in self.parameters())

    def ensure_max_t(self, t_count):
        """Grows the precomputed time-embedding table when a schedule reaches past it (t_start >= 64)."""
        if t_count > self.max_t:
            self.max_t, self._handle = int(t_count), None

    def cuda_handle(self):
        """Packs the current parameters for the CUDA plan on the device that holds them (re-packed if any parameter changed or moved)."""
        key = self._weights_key()
        if self._handle is None or key != self._handle_key:
            dev = next(self.parameters()).device
            if dev.type != "cuda":
                raise RuntimeError("UNetModel: parameters must live on a CUDA device (the B200 build has no CPU path); call .to('cuda:N')")
            cfg = _eng.unet_config(self.in_channels, self.model_channels, self.out_channels, self.num_res_blocks,
                                   list(self.attention_resolutions), list(self.channel_mult), self.num_heads, self.precision, self.max_t)
            with torch.cuda.device(dev):
                self._handle = _eng.UNetHandle(cfg, self.state_dict(), device=dev)
            self._handle_key = key
        return self._handle

    def forward(self, x, timesteps):
        """x [N,C,H,W] on the GPU; `timesteps` must hold one value for the whole batch (model.py:564)."""
        t = timesteps.reshape(-1)
        t0 = int(t[0])
        if t.numel() > 1 and not bool((t == t0).all()):
            raise NotImplementedError("per-sample timesteps are a training feature; the inference path shares t across the batch")
        self.ensure_max_t(t0 + 1)
        return self.cuda_handle().forward(x.contiguous().float(), t0)


# ---------------------------------------------------------------------------------------------
def cosine_beta_schedule(timesteps, s=0.008, schedule_power=1):
    if s != 0.008:
        raise NotImplementedError("only the reference offset s=0.008 is built into the schedule kernel")
    return torch.from_numpy(_eng.cosine_beta_schedule(timesteps, schedule_power))


class GaussianDiffusion:
    def __init__(self, timesteps=1000, beta_schedule='linear', schedule_power=1):
        if beta_schedule != 'cosine':
            raise NotImplementedError("the progressive path uses beta_schedule='cosine' (train_test_utils.py:221-245)")
        self.timesteps, self.schedule_power = timesteps, schedule_power
        self.betas = cosine_beta_schedule(timesteps, schedule_power=schedule_power)
        self.alphas = 1. - self.betas
        self.alphas_cumprod = torch.cumprod(self.alphas, axis=0)
        self.alphas_cumprod_prev = torch.nn.functional.pad(self.alphas_cumprod[:-1], (1, 0), value=1.)
        self.sqrt_alphas_cumprod = torch.sqrt(self.alphas_cumprod)
        self.sqrt_one_minus_alphas_cumprod = torch.sqrt(1.0 - self.alphas_cumprod)
        self.log_one_minus_alphas_cumprod = torch.log(1.0 - self.alphas_cumprod)
        self.sqrt_recip_alphas_cumprod = torch.sqrt(1.0 / self.alphas_cumprod)
        self.sqrt_recipm1_alphas_cumprod = torch.sqrt(1.0 / self.alphas_cumprod - 1)
        self.posterior_variance = self.betas * (1.0 - self.alphas_cumprod_prev) / (1.0 - self.alphas_cumprod)
        self.posterior_log_variance_clipped = torch.log(self.posterior_variance.clamp(min=1e-20))
        self.posterior_mean_coef1 = self.betas * torch.sqrt(self.alphas_cumprod_prev) / (1.0 - self.alphas_cumprod)
        self.posterior_mean_coef2 = (1.0 - self.alphas_cumprod_prev) * torch.sqrt(self.alphas) / (1.0 - self.alphas_cumprod)

    def _extract(self, a, t, x_shape):
        out = a.to(t.device).gather(0, t).float()
        return out.reshape(t.shape[0], *((1,) * (len(x_shape) - 1)))

    def q_sample(self, x_start, t, noise=None, seed=0):
        ts = int(t.reshape(-1)[0])
        return _eng.q_sample(x_start.contiguous(), float(self.sqrt_alphas_cumprod[ts].float()),
                             float(self.sqrt_one_minus_alphas_cumprod[ts].float()), noise=noise, seed=seed)

    @torch.no_grad()
    def guided_reverse_process(self, model, img, t_start=None, clip=True, lambda_ratio=1, eta=0.5, save_states=False,
                               mode="img", constant_guidance=None, noise=None, seed=0, **kwargs):
        if kwargs.get("only_convertor", False):
            return [img], None, None
        if kwargs.get("normal", False):
            raise NotImplementedError("normal=True (Yeo-Johnson) is off in every shipped config and not on the B200 path")
        if save_states:
            raise NotImplementedError("save_states=True copies every reverse step to the host; not on the B200 path")
        ks = kwargs.get("kernel_size_proj" if mode == "proj" else "kernel_size_img", 4)
        amp = kwargs.get("amplitude_proj" if mode == "proj" else "amplitude_img", 7.0)
        img = img.contiguous().float()
        ldct = kwargs.get("ldct", None)
        ldct = None if ldct is None else ldct.contiguous().float()
        if t_start is None:
            return self._adaptive_schedule_process(model, img, clip, lambda_ratio, mode, constant_guidance, ks, amp, ldct,
                                                   kwargs.get("noise_strength", None), noise, seed)
        p = _eng.guided_params(mode, copy(t_start), clip, lambda_ratio, eta, constant_guidance, ks, amp, self.schedule_power,
                               self.timesteps, seed)
        model.ensure_max_t(max(t_start) + 1)
        out = _eng.guided_process(model.cuda_handle(), p, img, ldct, noise)
        return [out[k] for k in range(out.shape[0])], [], None

    # schedules of the adaptive branch (reference Model/model.py:582-613): (t_start list, eta) per noise-strength class
    _ADAPTIVE_PROJ = {"high": ([30, 25, 20], 0.6), "mid": ([20, 18, 15], 0.5), "low": ([15, 15, 15], 0.5)}
    _ADAPTIVE_IMG = {"high": ([15, 15, 15], 0.6), "mid": ([15, 12, 10], 0.55), "low": ([10, 10, 10], 0.5)}

    def _adaptive_schedule_process(self, model, img, clip, lambda_ratio, mode, constant_guidance, ks, amp, ldct, noise_strength, noise, seed):
        """t_start=None (reference :531-535, 582-613, 639-640): a probing iteration with t_start = 20 and the cosine lambda schedule,
        then the schedule and eta are picked PER SLICE -- projection domain: from max(exp(amplitude * delta-map)) (>= 30 high, >= 4.5 mid,
        else low); image domain: from the `noise_strength` the projection stage reported -- and the process continues from the input
        with the per-pixel lambda map.  The only host round trip is one read of B floats (projection domain); slices that picked the
        same schedule continue as one batch.  Returns (iterates without the probe + mean of the last two, [], noise_strength) where
        noise_strength is the reference's string for one slice (or when all slices agree) and a per-slice list otherwise."""
        if constant_guidance is not None:
            raise ValueError("t_start=None with a constant guidance returns an empty list in the reference (the schedule is only "
                             "selected in the adaptive-lambda branch, Model/model.py:575); pass constant_guidance=None or explicit lists")
        B = img.shape[0]
        model.ensure_max_t(31)
        handle = model.cuda_handle()
        probe_p = _eng.guided_params(mode, [20], clip, lambda_ratio, 0.5, None, ks, amp, self.schedule_power, self.timesteps, seed)
        probe = _eng.guided_process(handle, probe_p, img, ldct, None if noise is None else noise[:21])[0]
        if mode == "proj":
            lam_exp = _eng.delta_lambda_map(probe, img, ks, amp, "proj")
            dmax = _eng.delta_exp_max(probe, img, ks, amp).cpu().numpy()            # the one device -> host read of this branch
            classes = ["high" if v >= 30 else ("mid" if v >= 4.5 else "low") for v in dmax]
            table = self._ADAPTIVE_PROJ
        else:
            lam_exp = _eng.delta_lambda_map_img(probe, img, ks, amp)
            ns = noise_strength if isinstance(noise_strength, (list, tuple)) else [noise_strength] * B
            if len(ns) != B:
                raise ValueError(f"noise_strength holds {len(ns)} entries for a batch of {B} slices")
            classes = ["low" if v is None else v for v in ns]
            table = self._ADAPTIVE_IMG
        if any(c not in table for c in classes):
            raise ValueError(f"noise_strength must be 'high', 'mid', 'low' or None, got {sorted(set(classes))}")
        out = None
        for cls in sorted(set(classes)):
            idx = [j for j, c in enumerate(classes) if c == cls]
            whole = len(idx) == B
            sel = torch.tensor(idx, device=img.device)
            take = (lambda t: t) if whole else (lambda t: t.index_select(0, sel).contiguous())
            ts_list, eta = table[cls]
            p = _eng.guided_params(mode, list(ts_list), clip, lambda_ratio, eta, None, ks, amp, self.schedule_power, self.timesteps, seed)
            tape = None
            if noise is not None:
                tape = noise[21:] if whole else noise[21:].index_select(1, sel).contiguous()
            res = _eng.guided_process_resume(handle, p, take(img), take(lam_exp), 21, None if ldct is None else take(ldct), tape)
            if whole:
                out = res
            else:
                out = torch.empty((res.shape[0], B) + tuple(res.shape[2:]), device=img.device) if out is None else out
                out.index_copy_(1, sel, res)
        if mode == "img":
            strength = noise_strength
        else:
            strength = classes[0] if len(set(classes)) == 1 else classes
        return [out[k] for k in range(out.shape[0])], [], strength

    @torch.no_grad()
    def ddim_sample(self, sample_img, model, condition, t_start, condition_lambda=0.5, batch_size=1, ddim_timesteps=2,
                    ddim_discr_method="uniform", ddim_eta=0.0, clip_denoised=True, noise=None, seed=0, call_base=0):
        """Guided DDIM sub-sequence of the sparse sampler (reference model.py:654-720): the timestep sequence and the prediction /
        guidance / x_{t-1} algebra of the reference, each step = one UNet forward + one fused `ipdm_sampler_step_ddim`.
        `noise`: optional list of tensors, one per step (the reference draws randn_like every step even when ddim_eta = 0)."""
        if ddim_discr_method == 'uniform':
            seq = np.linspace(t_start - 1, 0, ddim_timesteps + 1).astype(int)[0:-1]
        elif ddim_discr_method == 'quad':
            seq = ((np.linspace(0, np.sqrt(self.timesteps * .8), ddim_timesteps)) ** 2).astype(int)
        else:
            raise NotImplementedError(f'There is no ddim discretization method called "{ddim_discr_method}"')
        prev = np.append(seq[1:], np.array([0]))
        f32 = lambda v: v.float()                                                   # _extract(...).float()
        for i in range(ddim_timesteps):
            t, tp = int(seq[i]), int(prev[i])
            a_t, a_p = f32(self.alphas_cumprod[t]), f32(self.alphas_cumprod[tp])
            sig_dir = ddim_eta * torch.sqrt((1 - a_p) / (1 - a_t) * (1 - a_t / a_p))
            coef8 = [f32(self.sqrt_alphas_cumprod[t]), f32(self.sqrt_one_minus_alphas_cumprod[t]), 1.0 / torch.sqrt(a_t),
                     torch.sqrt(1. - a_t) / torch.sqrt(a_t), torch.sqrt(a_p), 0.0, ddim_eta * f32(self.posterior_variance[t]),
                     torch.sqrt(1 - a_p - sig_dir ** 2)]
            eps = model(sample_img, torch.full((1,), t, device=sample_img.device, dtype=torch.long))
            sample_img = _eng.sampler_step_ddim(sample_img.contiguous(), condition.contiguous(), eps.contiguous(), [float(c) for c in coef8],
                                                float(condition_lambda), noise=None if noise is None else noise[i].contiguous(),
                                                clip=clip_denoised, with_noise=ddim_eta != 0.0, seed=seed, call_id=call_base + i)
        return sample_img

    @torch.no_grad()
    def sparse_guided_reverse_process(self, model, condition, t_start, condition_lambda_max=0.5, condition_lambda_min=0.25,
                                      batch_size=1, ddim_timesteps=[2], ddim_discr_method="uniform", ddim_eta=0.0, eta=0.5,
                                      clip_denoised=True, noise=None, seed=0):
        """Sparse (DDIM) guided sampler, reference model.py:726-759.  `noise`: optional tape in the reference's randn order:
        [q_sample, then one tensor per DDIM step]."""
        condition = condition.contiguous().float()
        ts0 = int(t_start[0])
        sample_img = _eng.q_sample(condition, float(self.sqrt_alphas_cumprod[ts0].float()), float(self.sqrt_one_minus_alphas_cumprod[ts0].float()),
                                   noise=None if noise is None else noise[0].contiguous(), seed=seed, call_id=0)
        condition_ = condition.clone()
        step = (condition_lambda_max - condition_lambda_min) / len(t_start)
        condition_lambda = np.arange(condition_lambda_max, condition_lambda_min - step, -step)
        result, used = [], 1
        for i, t in enumerate(t_start):
            n_i = None if noise is None else [noise[used + k] for k in range(ddim_timesteps[i])]
            sample_img = self.ddim_sample(sample_img=sample_img, model=model, condition=condition, t_start=t,
                                          condition_lambda=condition_lambda[i], batch_size=batch_size, ddim_timesteps=ddim_timesteps[i],
                                          ddim_discr_method=ddim_discr_method, ddim_eta=ddim_eta, clip_denoised=clip_denoised,
                                          noise=n_i, seed=seed, call_base=1 + used)
            used += ddim_timesteps[i]
            condition = _eng.lincomb(eta, sample_img, 1 - eta, condition_)
            result.append(sample_img.clone())
        return result


def yeo_johnson_transform(img_tensor):
    raise NotImplementedError("normal=True (Yeo-Johnson, model.py:762-807) is off in every shipped config")
