"""A few launches of ONE tensor-core conv layer (for `ncu --set full -k regex:... --launch-skip 3 -c 1`):
python tools/one_conv.py c0 n h w cout [mode=3 bf16|1 tf32] [variant=0] [c1=0] [k=3] [with_res=1]
variant 5 = GroupNorm-fused halo kernel, 6 / 7 = width-folded thin layer with / without the fused GroupNorm"""
import ctypes
import os
import sys
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import torch
from ipdm_pytorch_b200 import _lib
L = _lib.lib()
torch.zeros(1, device="cuda")
c0, n, h, w, cout = (int(a) for a in sys.argv[1:6])
mode = int(sys.argv[6]) if len(sys.argv) > 6 else 3
variant = int(sys.argv[7]) if len(sys.argv) > 7 else 0
c1 = int(sys.argv[8]) if len(sys.argv) > 8 else 0
k = int(sys.argv[9]) if len(sys.argv) > 9 else 3
with_res = int(sys.argv[10]) if len(sys.argv) > 10 else 1
ms, fl = ctypes.c_float(), ctypes.c_double()
_lib.check(L.ipdm_debug_conv_time(c0, c1, n, h, w, cout, k, 1, mode, variant, with_res, 5, ctypes.byref(ms), ctypes.byref(fl)), "conv_time")
print(f"{c0}+{c1}->{cout} k{k} res{with_res} {n}x{h}x{w} mode {mode} variant {variant}: {ms.value:.3f} ms, {fl.value / ms.value / 1e9:.0f} TFLOP/s")
