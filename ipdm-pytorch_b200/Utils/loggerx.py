"""Checkpoint / option I/O of the reference's Utils/loggerx.py (formats kept, plotting dropped).

Checkpoints are plain `torch.save(module.state_dict())` files named `<module_name>-<epoch>` under
`save_models/` (:62-80); `load_network` strips a DataParallel `module.` prefix (:131-140).
"""
import json
import os
import os.path as osp
import time
from collections import OrderedDict

import torch
import torch.distributed as dist


class LoggerX(object):
    def __init__(self, save_root, opt):
        self.models_save_dir = osp.join(save_root, 'save_models')
        self.curve_save_dir = osp.join(save_root, 'save_curve')
        os.makedirs(self.models_save_dir, exist_ok=True)
        self.modules, self.module_names = [], []
        self.world_size, self.local_rank = 1, 0
        self.curve_data = dict()

    def _named(self):
        return [(n, m) for n, m in zip(self.module_names, self.modules) if m is not None]

    def checkpoints(self, epoch):
        if self.local_rank != 0:
            return
        for name, module in self._named():
            torch.save(module.state_dict(), osp.join(self.models_save_dir, f'{name}-{epoch}'))

    def load_checkpoints(self, epoch, model_load_path):
        print("load model...")
        for name, module in self._named():
            path = osp.join(model_load_path, f'{name}-{epoch}')
            if osp.exists(path):
                module.load_state_dict(load_network(path))
                print(f"load {name} finished!")

    def save_option(self, opt):
        with open(osp.join(self.models_save_dir, 'option.json'), 'w') as f:
            f.write(json.dumps(opt.__dict__, sort_keys=False, indent=4, separators=(',', ': ')))

    def msg(self, stats, step):
        items = stats.items() if isinstance(stats, dict) else [(f"v{i}", v) for i, v in enumerate(stats)]
        parts = []
        for name, var in items:
            if isinstance(var, torch.Tensor):
                var = reduce_tensor(var.detach().mean()).item()
            parts.append('{} {:2.5f}'.format(name, var))
        if self.local_rank == 0:
            print('[{}] {:05d}, {}'.format(time.strftime("%Y-%m-%d %H:%M:%S", time.localtime()), step, ', '.join(parts)))


def load_network(state_dict):
    if isinstance(state_dict, str):
        state_dict = torch.load(state_dict, map_location='cpu')
    return OrderedDict((k.replace('module.', ''), v) for k, v in state_dict.items())


def reduce_tensor(tensor, world_size=None):
    rt = tensor.clone()
    if dist.is_available() and dist.is_initialized():
        dist.all_reduce(rt, op=dist.ReduceOp.SUM)
    else:
        world_size = 1
    if world_size is not None:
        rt /= world_size
    return rt
