"""GPU input pipeline (rb_frames_to_clip = Stack -> ToTorchFormatTensor -> GroupNormalize of rubiksnet/transforms.py:66-79,
329-363) against the same arithmetic in PyTorch, and the evaluation loop on uint8 stacks vs pre-normalised float input."""
import os
import sys

import pytest
import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from helpers import REPO  # noqa: E402,F401

import rubiksnet_b200 as rb  # noqa: E402
from rubiksnet_b200.evaluate import evaluate, frames_to_clip  # noqa: E402

pytestmark = pytest.mark.gpu
MEAN, STD = (0.485, 0.456, 0.406), (0.229, 0.224, 0.225)


def _reference(frames_u8):
    x = frames_u8.permute(0, 3, 1, 2).float().div(255)          # ToTorchFormatTensor
    c = x.shape[1]
    mean = torch.tensor(MEAN * (c // 3), device=x.device).view(1, c, 1, 1)  # GroupNormalize: mean repeated per frame
    std = torch.tensor(STD * (c // 3), device=x.device).view(1, c, 1, 1)
    return (x - mean) / std


@pytest.mark.parametrize("shape", [(2, 224, 224, 24), (3, 17, 23, 6), (1, 5, 5, 3), (2, 64, 48, 48)])
def test_frames_to_clip_matches_reference_transforms(shape):
    torch.manual_seed(0)
    frames = torch.randint(0, 256, shape, dtype=torch.uint8, device="cuda")
    want = _reference(frames)
    got = frames_to_clip(frames, MEAN, STD)
    assert got.dtype == torch.float32 and got.shape == want.shape
    assert (got - want).abs().max().item() <= 2e-6 * max(1.0, want.abs().max().item())
    got16 = frames_to_clip(frames, MEAN, STD, out_dtype=torch.bfloat16)
    assert (got16.float() - want).abs().max().item() <= 1e-2 * max(1.0, want.abs().max().item())


def test_evaluate_on_uint8_stacks_equals_float_input():
    torch.manual_seed(1)
    net = rb.RubiksNet(tier="tiny", num_classes=9, num_frames=8).cuda()
    g = torch.Generator().manual_seed(2)
    stacks = [torch.randint(0, 256, (2 * 1, 224, 224, 24), dtype=torch.uint8, generator=g) for _ in range(2)]
    labels = [torch.randint(0, 9, (2,), generator=g) for _ in range(2)]
    as_float = [(_reference(s.cuda()).cpu(), l) for s, l in zip(stacks, labels)]
    a = evaluate(net, list(zip(stacks, labels)), num_crops=1, frames=8)
    b = evaluate(net, as_float, num_crops=1, frames=8)
    assert a["videos"] == b["videos"] == 4
    assert a["prec1"] == b["prec1"] and a["prec5"] == b["prec5"]
