"""Batch data parallelism for RubiksNet training: one process per GPU, the batch is sharded across ranks
and the ONLY exchange is one gradient all-reduce per step (the reference itself has no
distributed training code, only nn.DataParallel for evaluation, scripts/test_models.py:153).

All 8.5 M gradients (34 MB fp32 for RubiksNet-Large) live in ONE flat buffer whose slices are the
parameters' .grad views, so the exchange is a single NCCL all-reduce over NVLink/NVSwitch with no
bucketing copies, and zeroing the gradients is a single memset.  BatchNorm statistics stay per replica
(the reference has no SyncBN) and the shift gradients are normalised inside the op BEFORE the
all-reduce, which is what a replica-sum would see.
"""
import torch
import torch.distributed as dist

__all__ = ["FlatGradAllReduce", "shard_batch"]


class FlatGradAllReduce:
    def __init__(self, module, process_group=None):
        self.params = [p for p in module.parameters() if p.requires_grad]
        self.group = process_group
        total = sum(p.numel() for p in self.params)
        ref = self.params[0]
        self.flat = torch.zeros(total, dtype=torch.float32, device=ref.device)
        off = 0
        for p in self.params:
            assert p.dtype == torch.float32, "master parameters are kept in float32"
            p.grad = self.flat[off:off + p.numel()].view_as(p)
            off += p.numel()

    def zero_grad(self):
        """One memset; keeps .grad aliased to the flat buffer (do NOT call module.zero_grad(set_to_none=True))."""
        self.flat.zero_()

    @property
    def world_size(self):
        return dist.get_world_size(self.group) if dist.is_available() and dist.is_initialized() else 1

    def all_reduce(self):
        """Averages the gradients over ranks in place (sum over ranks / world size)."""
        ws = self.world_size
        if ws == 1:
            return
        backend = dist.get_backend(self.group)
        if backend == "nccl":
            dist.all_reduce(self.flat, op=dist.ReduceOp.AVG, group=self.group)
        else:  # gloo has no AVG
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=self.group)
            self.flat.div_(ws)

    def check_aliasing(self):
        """True if every parameter's .grad still points into the flat buffer."""
        base = self.flat.untyped_storage().data_ptr()
        return all(p.grad is not None and p.grad.untyped_storage().data_ptr() == base for p in self.params)


def shard_batch(global_batch, rank, world_size):
    """[start, stop) of this rank's clips; the remainder goes to the first ranks."""
    per, rem = divmod(global_batch, world_size)
    start = rank * per + min(rank, rem)
    return start, start + per + (1 if rank < rem else 0)
