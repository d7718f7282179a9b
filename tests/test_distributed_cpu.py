"""world_size-2 gloo test (CPU) of the data-parallel host logic: gradient ranges reported in reverse layer order are
coalesced into buckets, all-reduced asynchronously, and every rank ends with the identical summed gradient buffer.
The NCCL path on GPUs uses exactly this code with CUDA tensors (fastvideotagging_b200/trainer.py)."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


class _FakeFlat:
    def __init__(self, n):
        self.g = torch.zeros(n)


class _FakeNet:
    def __init__(self, n):
        self._flat = _FakeFlat(n)
        self._trainer = None

    def _attach_trainer(self, t):
        self._trainer = t


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from fastvideotagging_b200.trainer import Trainer
        n = 10000
        net = _FakeNet(n)
        tr = Trainer(net, "sgd", {"learning_rate": 0.1, "momentum": 0.9, "wd": 0.0}, bucket_bytes=4 * 1500)
        assert tr._distributed
        rng = np.random.default_rng(100 + rank)
        net._flat.g.copy_(torch.from_numpy(rng.normal(size=n).astype(np.float32)))
        local = net._flat.g.clone()
        # ranges become ready from the end of the buffer towards the start, in uneven pieces
        edges = [10000, 9400, 9100, 7000, 6990, 4000, 1200, 0]
        launched = []
        for hi, lo in zip(edges[:-1], edges[1:]):
            tr.on_grads_ready(lo, hi)
            launched.append(len(tr._handles))
        tr.allreduce_grads()
        assert tr._pending is None and tr._handles == []
        # buckets were launched during "backward", not only at the end
        assert launched[-2] >= 2, launched
        gathered = [torch.zeros(n) for _ in range(world)]
        dist.all_gather(gathered, local)
        expect = sum(gathered)
        assert torch.allclose(net._flat.g, expect, atol=1e-6)
        out[rank] = float(net._flat.g.sum())
    finally:
        dist.destroy_process_group()


def test_bucketed_allreduce_world_size_2():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    assert len(out) == 2 and abs(out[0] - out[1]) < 1e-3
