"""Per-shape timing of the tensor-core conv kernels: python tools/bench_conv.py [batch]  (prints a markdown table)."""
import ctypes
import os
import sys
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import torch
from ipdm_pytorch_b200 import _lib
L = _lib.lib()
torch.zeros(1, device="cuda")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4
SHAPES = [  # c0, c1, n, h, w, cout, k, stride
    (128, 0, B, 500, 228, 128, 3, 1), (256, 0, B, 500, 228, 128, 3, 1), (128, 0, B, 1000, 456, 128, 3, 1), (144, 0, B, 1000, 456, 16, 3, 1),
    (128, 0, B, 250, 114, 128, 3, 1), (256, 0, B, 125, 57, 256, 3, 1), (256, 0, B, 63, 29, 256, 3, 1), (64, 0, B, 512, 512, 64, 3, 1),
    (128, 0, B, 512, 512, 64, 3, 1), (128, 0, B, 512, 512, 128, 3, 1), (128, 0, B, 256, 256, 128, 3, 1), (256, 0, B, 64, 64, 256, 3, 1),
    (256, 0, B, 32, 32, 256, 3, 1), (128, 128, B, 500, 228, 128, 3, 1), (64, 64, B, 512, 512, 64, 3, 1),
]
NAMES = {1: "one-tile", 2: "halo", 3: "persistent", 4: "halo-persistent", 5: "halo-persistent + fused GroupNorm/SiLU (raw fp32 sources)"}
ONLY = [int(v) for v in sys.argv[2].split(",")] if len(sys.argv) > 2 else (3, 4, 5)
print("| shape | mode | kernel | ms | TFLOP/s |\n|---|---|---|---:|---:|")
for sh in SHAPES:
    c0, c1, n, h, w, cout, k, stride = sh
    for mode, mname in ((1, "tf32"), (3, "bf16")):
        for fg in ONLY:
            ms, fl = ctypes.c_float(), ctypes.c_double()
            rc = L.ipdm_debug_conv_time(c0, c1, n, h, w, cout, k, stride, mode, fg, 1, 10, ctypes.byref(ms), ctypes.byref(fl))
            if rc != 0 and fg == 5:
                continue                                               # shape not eligible for the fused kernel
            _lib.check(rc, "conv_time")
            print(f"| {c0}+{c1}->{cout} k{k} s{stride} {n}x{h}x{w} | {mname} | {NAMES[fg]} | {ms.value:.3f} | {fl.value / ms.value / 1e9:.0f} |")
