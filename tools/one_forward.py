"""One full-size UNet forward of each network (for ncu captures): python tools/one_forward.py [n_forwards] [precision] [batch] [which]"""
import os
import sys
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO); sys.path.insert(0, os.path.join(REPO, "ipdm-pytorch_b200"))
import torch
from Model.model import UNetModel
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1
prec = sys.argv[2] if len(sys.argv) > 2 else "tf32"
B = int(sys.argv[3]) if len(sys.argv) > 3 else 1
which = sys.argv[4] if len(sys.argv) > 4 else "proj"
torch.manual_seed(0)
if which in ("proj", "both"):
    net = UNetModel(in_channels=1, model_channels=64, out_channels=1, attention_resolutions=[16, 32],
                    channel_mult=[0.0625, 0.125, 0.25, 2, 2, 4, 4]).cuda().eval()
    net.set_precision(prec)
    x = 3 * torch.rand(B, 1, 2000, 912, device="cuda")
    for i in range(n):
        y = net(x, torch.full((1,), 7, device="cuda"))
    torch.cuda.synchronize()
    print("proj ok", float(y.std()), float(y.double().abs().sum()))
if which in ("img", "both"):
    net = UNetModel(in_channels=1, model_channels=64, out_channels=1, attention_resolutions=[8, 16], channel_mult=[1, 1, 2, 2, 4, 4]).cuda().eval()
    net.set_precision(prec)
    x = 0.2 * torch.rand(B, 1, 512, 512, device="cuda")
    for i in range(n):
        y = net(x, torch.full((1,), 7, device="cuda"))
    torch.cuda.synchronize()
    print("img ok", float(y.std()), float(y.double().abs().sum()))
