"""Training-only index sampler kept for import compatibility (reference Utils/sampler.py:6-50):
an endless, deterministically shuffled, rank-strided stream of dataset indices."""
import numpy as np
import torch
import torch.distributed as dist
from torch.utils.data.sampler import Sampler


class RandomSampler(Sampler):
    def __init__(self, dataset, batch_size, num_iter, restore_iter=0, weights=None, replacement=True, seed=0, shuffle=True):
        self.dataset, self.batch_size, self.num_iter, self.restore_iter = dataset, batch_size, num_iter, restore_iter
        self.seed, self.shuffle = seed, shuffle
        ready = dist.is_available() and dist.is_initialized()
        self.rank = dist.get_rank() if ready else 0
        self.world = dist.get_world_size() if ready else 1
        self.num_samples = num_iter * batch_size

    def __len__(self):
        return max(0, (self.num_iter - self.restore_iter) * self.batch_size)

    def __iter__(self):
        n = len(self.dataset)
        epoch, produced, skip = 0, 0, self.restore_iter * self.batch_size
        while produced < self.num_samples:
            if self.shuffle:
                g = torch.Generator().manual_seed(self.seed + epoch)
                order = torch.randperm(n, generator=g).numpy()
            else:
                order = np.arange(n)
            for idx in order[self.rank::self.world]:
                if produced >= self.num_samples:
                    return
                produced += 1
                if produced > skip:
                    yield int(idx)
            epoch += 1
