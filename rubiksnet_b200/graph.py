"""Whole-step CUDA graph for RubiksNet training.

One RubiksNet-Large step is ~1 400 kernel launches (765 of them librubiks_b200's, most on 14x14 maps that run for
10-70 us); driven from Python the launch stream, not the GPU, sets the pace.  `GraphedStep` records one
forward + backward (+ gradient all-reduce) + optimizer update into a CUDA graph over static input buffers and replays
it: same kernels, same order, same arithmetic, no per-launch host work.

    step = GraphedStep(step_fn, example_clips, example_labels)   # step_fn(clips, labels) -> loss tensor
    loss = step(clips, labels)                                    # copies into the static buffers, replays

Everything librubiks_b200 launches goes to torch's current stream and its scratch buffers come from torch's allocator,
so the capture needs nothing special; the library's launch counter ticks at capture time and `launches_per_replay`
records how many of its kernels one replay runs.
"""
import torch

from . import _lib

__all__ = ["GraphedStep"]


class GraphedStep:
    def __init__(self, step_fn, clips, labels, warmup=3):
        assert clips.is_cuda and labels.is_cuda
        self.clips = clips.clone()
        self.labels = labels.clone()
        self.step_fn = step_fn
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):  # allocator / cuDNN autotune / optimizer state settle outside the capture
            for _ in range(warmup):
                step_fn(self.clips, self.labels)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        before = _lib.launch_count()
        with torch.cuda.graph(self.graph):
            self.loss = step_fn(self.clips, self.labels)
        self.launches_per_replay = _lib.launch_count() - before

    def __call__(self, clips=None, labels=None):
        """Replays the step on `clips` / `labels` (device or pinned-host tensors; None = reuse the static buffers).
        Returns the static loss tensor (overwritten by the next replay)."""
        if clips is not None and clips.data_ptr() != self.clips.data_ptr():
            self.clips.copy_(clips, non_blocking=True)
        if labels is not None and labels.data_ptr() != self.labels.data_ptr():
            self.labels.copy_(labels, non_blocking=True)
        self.graph.replay()
        return self.loss
