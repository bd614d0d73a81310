"""Evaluation loop of scripts/test_models.py (:140-200 of the reference) for one process per GPU, and the GPU side of its
input pipeline.

The reference evaluates under nn.DataParallel with a CPU loader that permutes, scales and normalises every frame stack
(rubiksnet/transforms.py:66-79,329-363).  Here
  * `frames_to_clip` does Stack -> ToTorchFormatTensor -> GroupNormalize in one librubiks_b200 kernel on the uint8 stack
    that crossed PCIe (rb_frames_to_clip), and
  * `evaluate` shards the batches over the ranks of a process group (batch i goes to rank i % world), averages the logits of
    the crops / clips of a video exactly like the reference (`rst.reshape(batch, num_crop, -1).mean(1)`), and all-reduces the
    top-1 / top-5 hit counters once at the end -- the only exchange, like the gradient all-reduce in training.
"""
import ctypes
import time

import torch
import torch.distributed as dist

from . import _lib
from .rubiksnet_cuda import _on_device

__all__ = ["frames_to_clip", "topk_hits", "evaluate"]

_IMAGENET_MEAN = (0.485, 0.456, 0.406)  # RubiksNet.input_mean / input_std (rubiksnet/models.py:112-113)
_IMAGENET_STD = (0.229, 0.224, 0.225)


def frames_to_clip(frames_u8, mean=_IMAGENET_MEAN, std=_IMAGENET_STD, out_dtype=torch.float32, div255=True):
    """uint8 [N, H, W, 3*T] frame stacks (CUDA) -> normalised [N, 3*T, H, W] (= [N*T, 3, H, W] for the model)."""
    assert frames_u8.is_cuda and frames_u8.dtype == torch.uint8 and frames_u8.dim() == 4 and frames_u8.is_contiguous()
    n, h, w, c = frames_u8.shape
    assert c % 3 == 0, "channels must be 3 * frames"
    out = torch.empty(n, c, h, w, dtype=out_dtype, device=frames_u8.device)
    m = (ctypes.c_float * 3)(*[float(v) for v in mean])
    s = (ctypes.c_float * 3)(*[float(v) for v in std])
    with _on_device(frames_u8.device), _lib.timed("frames_to_clip", _lib.nbytes(frames_u8, out)):
        _lib.check(_lib.lib().rb_frames_to_clip(_lib.ptr(frames_u8), _lib.ptr(out), _lib.dtype_code(out), n, h, w, c, m, s,
                                                int(bool(div255)), _lib.stream_handle(frames_u8.device)))
    return out


def topk_hits(logits, target, topk=(1, 5)):
    """Number of samples whose label is among the k largest logits, for every k (scripts/test_models.py:30-41 returns the
    same quantity as a percentage of the batch)."""
    maxk = min(max(topk), logits.shape[1])
    pred = logits.topk(maxk, 1, True, True).indices
    correct = pred.eq(target.view(-1, 1))
    return [int(correct[:, :min(k, maxk)].any(dim=1).sum().item()) for k in topk]


def evaluate(net, batches, num_crops=1, frames=8, process_group=None, device=None, autocast_dtype=None,
             mean=_IMAGENET_MEAN, std=_IMAGENET_STD):
    """Runs `net` (eval mode, no_grad) over `batches` = iterable of (data, label):
         data  float [B, num_crops*frames*3, H, W] (what the reference's loader yields), or
               uint8 [B*num_crops, H, W, frames*3] frame stacks (normalised on the GPU by frames_to_clip);
         label int64 [B].
    Every rank of `process_group` takes the batches with index % world == rank.  Returns a dict with the global
    prec@1 / prec@5 (percent), the number of videos and the wall-clock seconds per video of this rank."""
    world = dist.get_world_size(process_group) if dist.is_available() and dist.is_initialized() else 1
    rank = dist.get_rank(process_group) if world > 1 else 0
    if device is None:
        device = next(net.parameters()).device
    was_training = net.training
    net.eval()
    hits1 = hits5 = count = 0
    t0 = time.time()
    try:
        with torch.no_grad():
            for i, (data, label) in enumerate(batches):
                if i % world != rank:
                    continue
                b = label.numel()
                data = data.to(device, non_blocking=True)
                if data.dtype == torch.uint8:
                    data = frames_to_clip(data.contiguous(), mean, std)
                data = data.contiguous().view(b * num_crops, frames, 3, data.shape[-2], data.shape[-1])
                if autocast_dtype is not None and device.type == "cuda":
                    with torch.autocast("cuda", dtype=autocast_dtype):
                        rst = net(data)
                else:
                    rst = net(data)
                rst = rst.float().reshape(b, num_crops, -1).mean(1)
                h1, h5 = topk_hits(rst, label.to(rst.device), (1, 5))
                hits1, hits5, count = hits1 + h1, hits5 + h5, count + b
    finally:
        net.train(was_training)
    local_videos = count
    if world > 1:
        t = torch.tensor([hits1, hits5, count], dtype=torch.float64, device=device if dist.get_backend(process_group) == "nccl" else "cpu")
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=process_group)
        hits1, hits5, count = (int(v) for v in t.tolist())
    return {"prec1": 100.0 * hits1 / max(count, 1), "prec5": 100.0 * hits5 / max(count, 1), "videos": count,
            "sec_per_video": (time.time() - t0) / max(local_videos, 1)}
