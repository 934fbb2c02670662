"""Seeded inputs shared by the golden generator (oracle/make_golden.py) and the tests."""
import numpy as np
import torch

PROJ_CFG = dict(in_channels=1, model_channels=64, out_channels=1, attention_resolutions=[16, 32],
                channel_mult=[0.0625, 0.125, 0.25, 2, 2, 4, 4])
IMG_CFG = dict(in_channels=1, model_channels=64, out_channels=1, attention_resolutions=[8, 16],
               channel_mult=[1, 1, 2, 2, 4, 4])


def small_proj_input(seed, h=100, w=76):
    g = torch.Generator().manual_seed(seed)
    d = torch.arange(w, dtype=torch.float32)[None, :]
    v = torch.arange(h, dtype=torch.float32)[:, None]
    base = 3.0 * torch.exp(-((d - w / 2) / (w / 3.5)) ** 2) * (1 + 0.1 * torch.sin(2 * np.pi * v / h))
    x = base + 0.08 * torch.randn(h, w, generator=g)
    return x.clamp(min=0)[None, None].contiguous()


def small_img_input(seed, n=64):
    g = torch.Generator().manual_seed(seed)
    yy, xx = torch.meshgrid(torch.linspace(-1, 1, n), torch.linspace(-1, 1, n), indexing="ij")
    body = ((xx / 0.8) ** 2 + (yy / 0.6) ** 2 <= 1).float()
    x = 0.19 * body + 0.012 * body * torch.sin(6 * xx) + 0.004 * torch.randn(n, n, generator=g)
    return x.clamp(min=0)[None, None].contiguous()


def unet_small_input(name):
    seed, shape, t = (0, (1, 1, 100, 76), 7) if name == "proj" else (1, (1, 1, 64, 64), 3)
    g = torch.Generator().manual_seed(100 + seed)
    return seed, torch.randn(shape, generator=g), t


def noise_tape(shape, count, seed):
    g = torch.Generator().manual_seed(int(seed))
    return [torch.randn(shape, generator=g, dtype=torch.float32) for _ in range(count)]
