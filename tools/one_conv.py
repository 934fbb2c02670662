"""A few launches of ONE tensor-core conv layer (for `ncu --set full -k regex:... --launch-skip 3 -c 1`):
python tools/one_conv.py c0 n h w cout [mode=3 bf16|1 tf32] [variant=0]"""
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
ms, fl = ctypes.c_float(), ctypes.c_double()
_lib.check(L.ipdm_debug_conv_time(c0, 0, n, h, w, cout, 3, 1, mode, variant, 1, 2, ctypes.byref(ms), ctypes.byref(fl)), "conv_time")
print(f"{c0}->{cout} {n}x{h}x{w} mode {mode} variant {variant}: {ms.value:.3f} ms, {fl.value / ms.value / 1e9:.0f} TFLOP/s")
