"""One full-size projection-UNet forward (for ncu captures): python tools/one_forward.py [n_forwards] [precision]"""
import os
import sys
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO); sys.path.insert(0, os.path.join(REPO, "ipdm-pytorch_b200"))
import torch
from Model.model import UNetModel
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1
torch.manual_seed(0)
net = UNetModel(in_channels=1, model_channels=64, out_channels=1, attention_resolutions=[16, 32],
                channel_mult=[0.0625, 0.125, 0.25, 2, 2, 4, 4]).cuda().eval()
if len(sys.argv) > 2:
    net.set_precision(sys.argv[2])
x = 3 * torch.rand(1, 1, 2000, 912, device="cuda")
for i in range(n):
    y = net(x, torch.full((1,), 7, device="cuda"))
torch.cuda.synchronize()
print("ok", float(y.std()))
