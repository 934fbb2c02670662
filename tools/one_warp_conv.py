"""A few launches of ONE warp-MMA thin conv with fused GroupNorm (for timing / ncu): python tools/one_warp_conv.py c0 c1 cout k n h w [iters]"""
import ctypes
import os
import sys
import time
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import torch
from ipdm_pytorch_b200 import _lib
L = _lib.lib()
c0, c1, cout, k, n, h, w = (int(a) for a in sys.argv[1:8])
iters = int(sys.argv[8]) if len(sys.argv) > 8 else 5
dev = torch.device("cuda:0")
x0 = torch.randn(n, h, w, c0, device=dev)
x1 = torch.randn(n, h, w, c1, device=dev) if c1 else None
wt = (0.1 * torch.randn(cout, c0 + c1, k, k)).contiguous()
b = torch.randn(cout)
sc, sh = torch.rand(n, c0 + c1, device=dev) + 0.5, torch.randn(n, c0 + c1, device=dev)
res = torch.randn(n, h, w, cout, device=dev)
out = torch.empty(n, h, w, cout, device=dev)
p = lambda t: None if t is None else ctypes.c_void_p(t.data_ptr())
def run():
    _lib.check(L.ipdm_debug_conv(p(x0), c0, c0, p(x1), c1, c1, n, h, w, p(wt), p(b), cout, k, 1, 0, 0, p(sc), p(sh), p(res), cout, p(out), cout, 5, None), "conv")
run(); torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(iters):
    run()
torch.cuda.synchronize()
ms = (time.perf_counter() - t0) / iters * 1e3
gb = 4.0 * n * h * w * (c0 + c1 + 2 * cout) / 1e9
print(f"warp conv {c0}+{c1}->{cout} k{k} {n}x{h}x{w} dbg={os.environ.get('IPDM_WARP_DBG', '0')}: {ms:.3f} ms per call (incl. weight packing on the host), {gb / ms:.2f} TB/s algorithmic")
