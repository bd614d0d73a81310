"""Batch data parallelism for RubiksNet training: one process per GPU, the batch is sharded across ranks
and the ONLY exchange is one gradient all-reduce per step (the reference itself has no
distributed training code, only nn.DataParallel for evaluation, scripts/test_models.py:153).

Gradients are produced by autograd as usual (``.grad = None`` before every backward pass, so AccumulateGrad adopts each
gradient tensor instead of launching one add kernel per parameter -- 366 launches for RubiksNet-Large); for the
exchange the 8.5 M gradients (34 MB fp32) are packed into ONE flat buffer by a single multi-tensor copy, reduced with a
single NCCL all-reduce over NVLink/NVSwitch, and handed back as views of that buffer (no unpack copy).  BatchNorm
statistics stay per replica (the reference has no SyncBN) and the shift gradients are normalised inside the op BEFORE
the all-reduce, which is what a replica-sum would see.
"""
import torch
import torch.distributed as dist

__all__ = ["FlatGradAllReduce", "shard_batch"]


class FlatGradAllReduce:
    def __init__(self, module, process_group=None):
        self.params = [p for p in module.parameters() if p.requires_grad]
        self.group = process_group
        for p in self.params:
            assert p.dtype == torch.float32, "master parameters are kept in float32"
        self.flat = None  # the packed gradients of the last all_reduce()

    def zero_grad(self):
        """Drops the gradients: the next backward pass assigns instead of accumulating."""
        for p in self.params:
            p.grad = None

    @property
    def world_size(self):
        return dist.get_world_size(self.group) if dist.is_available() and dist.is_initialized() else 1

    def all_reduce(self):
        """Averages the gradients over ranks (sum over ranks / world size).  One pack kernel + one collective; afterwards
        every .grad is a view into the packed buffer.  No-op for a single rank."""
        ws = self.world_size
        if ws == 1:
            return
        grads = [p.grad if p.grad is not None else torch.zeros_like(p) for p in self.params]
        flat = torch._utils._flatten_dense_tensors(grads)
        if dist.get_backend(self.group) == "nccl":
            dist.all_reduce(flat, op=dist.ReduceOp.AVG, group=self.group)
        else:  # gloo has no AVG
            dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.group)
            flat.div_(ws)
        for p, v in zip(self.params, torch._utils._unflatten_dense_tensors(flat, grads)):
            p.grad = v
        self.flat = flat

    def check_aliasing(self):
        """True if every parameter's .grad points into the packed buffer of the last all_reduce()."""
        if self.flat is None:
            return False
        base = self.flat.untyped_storage().data_ptr()
        return all(p.grad is not None and p.grad.untyped_storage().data_ptr() == base for p in self.params)


def shard_batch(global_batch, rank, world_size):
    """[start, stop) of this rank's clips; the remainder goes to the first ranks."""
    per, rem = divmod(global_batch, world_size)
    start = rank * per + min(rank, rem)
    return start, start + per + (1 if rank < rem else 0)
