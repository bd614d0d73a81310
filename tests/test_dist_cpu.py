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


# ---------------------------------------------------------------- sharded evaluation loop (scripts/test_models.py:140-200)

class _ClipNet(nn.Module):
    """Stand-in for RubiksNet: [B, T, 3, H, W] clips -> [B, classes] logits."""

    def __init__(self):
        super().__init__()
        torch.manual_seed(1)
        self.fc = nn.Linear(3, 7)

    def forward(self, clips):
        return self.fc(clips.mean(dim=(1, 3, 4)))


def _eval_batches():
    g = torch.Generator().manual_seed(5)
    return [(torch.randn(4, 2 * 8 * 3, 6, 6, generator=g), torch.randint(0, 7, (4,), generator=g)) for _ in range(5)]


def _eval_worker(rank, world, port, out):
    from rubiksnet_b200.evaluate import evaluate
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    res = evaluate(_ClipNet(), _eval_batches(), num_crops=2, frames=8)
    if rank == 0:
        torch.save(res, out)
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_evaluation_matches_single_process(tmp_path):
    from rubiksnet_b200.evaluate import evaluate, topk_hits
    out = str(tmp_path / "e.pt")
    mp.spawn(_eval_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    got = torch.load(out)
    want = evaluate(_ClipNet(), _eval_batches(), num_crops=2, frames=8)
    assert got["videos"] == want["videos"] == 20
    assert got["prec1"] == want["prec1"] and got["prec5"] == want["prec5"]
    # the counters follow the reference's accuracy(): label among the k largest averaged logits
    logits = torch.tensor([[0.1, 0.9, 0.0], [0.8, 0.1, 0.1], [0.2, 0.3, 0.5]])
    assert topk_hits(logits, torch.tensor([1, 2, 2]), (1, 2)) == [2, 2]
