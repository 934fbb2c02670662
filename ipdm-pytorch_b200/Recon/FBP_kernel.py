"""Host-side mirror of the reference's Recon/FBP_kernel.py: class FBP with `convert`.

`FBP(device).convert(pj, flip=True)` keeps the reference contract (:86-122): ndarray or Tensor
[B,2000,912] / [2000,912] in, the same kind out, [B,512,512] float32 ON THE HOST.  The arithmetic
(cosine weighting, ramp filter, pixel-driven fan-beam backprojection) runs in libipdm_b200.so;
`convert_device` is the additive zero-copy entry the progressive pipeline uses.
"""
import numpy as np
import torch

from _ipdm_boot import engine as _eng


class FBP:
    def __init__(self, device="cuda:0"):
        if str(device) == "cpu":
            raise RuntimeError("FBP: the B200 build has no CPU path (the reference's numba CPU twin is the test oracle only)")
        self.device = torch.device(device)
        with torch.cuda.device(self.device):
            self._plan = _eng.FBPPlan()
        t = self._plan.tables()
        # geometry attributes of the reference object (FBP_kernel.py:32-56)
        self.os_, self.od, self.T, self.da = 59.5, 108.56 - 59.5, 0.0010125, 0.0010125
        self.D, self.N, self.M = 59.5, 912, 2000
        self.theta, self.nda, self.h_RL = t["theta"], t["nda"], t["h_RL"][:, None]

    def convert_device(self, pj, flip=True):
        """[B,2000,912] CUDA tensor -> [B,512,512] CUDA tensor, no host copy, current stream."""
        with torch.cuda.device(self.device):
            return self._plan.forward(pj.contiguous().float(), flip=flip)

    def convert(self, pj, flip=True):
        as_tensor = isinstance(pj, torch.Tensor)
        if as_tensor and pj.is_cuda:
            x = pj.detach().float()
            x = x[None] if x.dim() == 2 else x
            return self.convert_device(x.contiguous(), flip).cpu()
        arr = pj.detach().cpu().numpy() if as_tensor else np.asarray(pj)
        with torch.cuda.device(self.device):
            out = self._plan.convert_host(arr, flip)
        return torch.from_numpy(out) if as_tensor else out
