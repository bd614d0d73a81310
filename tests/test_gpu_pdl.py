"""Programmatic dependent launch (csrc/common.cuh: prologue -> griddepcontrol.wait -> trigger -> body) must not change a
single bit: every kernel of the library overlaps only its own prologue (and RB_W_RESIDENT weight staging) with the
previous kernel, so whole training steps with the switch on and off -- eager and replayed from a CUDA graph -- have to
agree exactly (the kernels are deterministic: fixed-order reductions, no atomics)."""
import os
import sys

import pytest
import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from helpers import REPO  # noqa: E402,F401

import rubiksnet_b200 as rb  # noqa: E402
from rubiksnet_b200 import _lib, ops  # noqa: E402
from rubiksnet_b200.graph import GraphedStep  # noqa: E402

pytestmark = pytest.mark.gpu


def _train(tier, variant, pdl, steps, graph=False, batch=2):
    _lib.set_dependent_launch(pdl)
    det, bench = torch.backends.cudnn.deterministic, torch.backends.cudnn.benchmark
    torch.backends.cudnn.deterministic, torch.backends.cudnn.benchmark = True, False  # conv1 is cuDNN's
    try:
        torch.manual_seed(0)
        net = rb.RubiksNet(tier=tier, variant=variant, num_classes=10, num_frames=8).cuda().train()
        opt = torch.optim.SGD(net.parameters(), lr=0.05, momentum=0.9)
        clips = torch.randn(batch, 8, 3, 224, 224, device="cuda")
        labels = torch.randint(0, 10, (batch,), device="cuda")

        def step(c, l):
            opt.zero_grad(set_to_none=False)
            with torch.autocast("cuda", dtype=torch.bfloat16):
                logits = net(c)
            loss = torch.nn.functional.cross_entropy(logits.float(), l)
            loss.backward()
            opt.step()
            return loss.detach()

        losses = []
        if graph:
            # GraphedStep's 3 warm-up steps are part of the trajectory: compare against an eager run of steps + 3
            g = GraphedStep(step, clips, labels, warmup=3)
            for _ in range(steps):
                losses.append(float(g()))
        else:
            for _ in range(steps):
                losses.append(float(step(clips, labels)))
        torch.cuda.synchronize()
        return losses, [p.detach().clone() for p in net.parameters()]
    finally:
        _lib.set_dependent_launch(True)
        torch.backends.cudnn.deterministic, torch.backends.cudnn.benchmark = det, bench


@pytest.mark.parametrize("tier,variant", [("tiny", "rubiks3d"), ("tiny", "rubiks3d-aq"), ("small", "rubiks3d")])
def test_dependent_launch_is_bit_exact_eager(tier, variant):
    l_off, p_off = _train(tier, variant, False, 3)
    l_on, p_on = _train(tier, variant, True, 3)
    assert l_on == l_off
    for a, b in zip(p_on, p_off):
        assert torch.equal(a, b)


def test_dependent_launch_is_bit_exact_in_a_graph():
    # GraphedStep runs 3 eager warm-up steps, captures one (which executes nothing) and then replays: 3 + 3 steps
    l_on, p_on = _train("tiny", "rubiks3d", True, 3, graph=True)
    l_ref, p_ref = _train("tiny", "rubiks3d", False, 6)
    assert l_on == l_ref[3:]
    for a, b in zip(p_on, p_ref):
        assert torch.equal(a, b)


def test_dependent_launch_chain_of_small_kernels():
    """Back-to-back dependent launches on small tensors (launch-latency regime): 200 conv -> BN chains, on vs off."""
    torch.manual_seed(1)
    x0 = torch.randn(8, 72, 14, 14, device="cuda").bfloat16()
    w = torch.randn(72, 72, device="cuda") / 72 ** 0.5

    def run(pdl):
        _lib.set_dependent_launch(pdl)
        try:
            x = x0.clone()
            for i in range(200):
                w_nk, w_kn = ops.pw_weight_pack(w * (1.0 + 0.001 * i))
                y = ops.pw_conv(x, w_nk, residual=x, resident=True)
                x = ops.pw_conv(y, w_kn, resident=True)
                x = (x.float() / x.float().abs().max().clamp_min(1.0)).bfloat16()
            torch.cuda.synchronize()
            return x
        finally:
            _lib.set_dependent_launch(True)

    assert torch.equal(run(True), run(False))
