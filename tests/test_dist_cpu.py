"""world_size-2 gloo test (CPU) of the data-parallel plumbing: sharded-batch gradients averaged with the
flat all-reduce equal the full-batch gradients."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp
import torch.nn as nn

from rubiksnet_b200.dp import FlatGradAllReduce, shard_batch


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _model():
    torch.manual_seed(0)
    return nn.Sequential(nn.Conv2d(3, 8, 1, bias=False), nn.ReLU(), nn.Conv2d(8, 4, 1), nn.AdaptiveAvgPool2d(1),
                         nn.Flatten(), nn.Linear(4, 5))


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(123)
    x, y = torch.randn(6, 3, 5, 5), torch.randint(0, 5, (6,))
    net = _model()
    red = FlatGradAllReduce(net)
    lo, hi = shard_batch(6, rank, world)
    red.zero_grad()
    # equal shard sizes => mean of shard-mean losses == full-batch mean loss
    nn.functional.cross_entropy(net(x[lo:hi]), y[lo:hi]).backward()
    assert not red.check_aliasing()
    red.all_reduce()
    assert red.check_aliasing()
    if rank == 0:
        torch.save(red.flat.clone(), out)
    dist.barrier()
    dist.destroy_process_group()


def test_flat_allreduce_matches_full_batch(tmp_path):
    out = str(tmp_path / "g.pt")
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    got = torch.load(out)
    torch.manual_seed(123)
    x, y = torch.randn(6, 3, 5, 5), torch.randint(0, 5, (6,))
    net = _model()
    nn.functional.cross_entropy(net(x), y).backward()
    want = torch.cat([p.grad.flatten() for p in net.parameters()])
    assert torch.allclose(got, want, atol=1e-6)


def test_shard_batch():
    assert [shard_batch(10, r, 4) for r in range(4)] == [(0, 3), (3, 6), (6, 8), (8, 10)]
    assert shard_batch(32, 0, 1) == (0, 32)
    assert [shard_batch(2, r, 4) for r in range(4)] == [(0, 1), (1, 2), (2, 2), (2, 2)]
