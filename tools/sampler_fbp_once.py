"""One guided reverse step (scalar lambda and per-pixel map, Philox and tape noise) and one FBP on 16 slices, for ncu captures:
    ncu --set full --clock-control none -k regex:'moments|apply_kernel|fbp_' -o gpurun_out/prof_sampler_fbp python tools/sampler_fbp_once.py"""
import os
import sys
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO); sys.path.insert(0, os.path.join(REPO, "ipdm-pytorch_b200"))
import torch
from ipdm_pytorch_b200 import engine, synthetic
B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(0)
x = 3 * torch.rand(B, 1, 2000, 912, device=dev, generator=g)
guide = x + 0.05 * torch.randn(x.shape, device=dev, generator=g)
eps = torch.randn(x.shape, device=dev, generator=g)
noise = torch.randn(x.shape, device=dev, generator=g)
coef = engine.step_coefficients(1000, 5, 7)
lam_map = 0.05 + 0.9 * torch.rand(B, 500, 228, device=dev, generator=g)
out = torch.empty_like(x)
for rep in range(2):
    engine.sampler_step(x, guide, eps, coef, 0.4, noise=None, out=out)          # scalar lambda, in-kernel Philox (the bench path)
    engine.sampler_step(x, guide, eps, coef, 0.4, noise=noise, out=out)         # scalar lambda, noise tape (the parity path)
    engine.sampler_step(x, guide, eps, coef, lam_map, noise=None, out=out)      # per-pixel lambda map
sino = torch.from_numpy(synthetic.cheap_sinogram(4, seed=7)).to(dev).repeat(B // 4 if B >= 4 else 1, 1, 1).contiguous()
plan = engine.FBPPlan(max_batch=sino.shape[0])
for rep in range(2):
    img = plan.forward(sino)
torch.cuda.synchronize()
print("ok", float(out.std()), float(img.std()))
