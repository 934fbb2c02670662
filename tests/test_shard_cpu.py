"""CPU, world_size 2 over gloo: the slice-sharding / final-gather logic bench.py uses for N > 1 GPUs."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


import conftest  # noqa: F401  (sys.path)
from ipdm_pytorch_b200.sharding import gather_slices, shard_range


def _worker(rank, world, port, n_slices, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    a, b = shard_range(n_slices, rank, world)
    local = torch.arange(a, b, dtype=torch.float32)[:, None, None, None].expand(b - a, 1, 4, 4).contiguous() * 2 + 1   # "denoise" slice i -> 2i+1
    ms = torch.tensor([10.0 + rank])
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)                       # timing: max over ranks
    full = gather_slices(local, n_slices)                           # the only collective: final gather
    if rank == 0:
        out.put((float(ms), full[:, 0, 0, 0].tolist()))
    dist.destroy_process_group()


def test_sharding_covers_every_slice_once():
    for n, w in ((64, 8), (7, 2), (5, 4), (512, 8), (1, 2)):
        spans = [shard_range(n, r, w) for r in range(w)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))
        assert max(e - s for s, e in spans) - min(e - s for s, e in spans) <= 1


def test_two_rank_gather_and_max_time():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, 7, q)) for r in range(2)]
    [p.start() for p in procs]
    [p.join(120) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    ms, vals = q.get()
    assert ms == 11.0 and vals == [2.0 * i + 1 for i in range(7)]
