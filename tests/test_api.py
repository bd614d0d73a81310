"""CPU tests of the host-side mirror of the reference's Python API (no kernels run)."""
import os

import numpy as np
import pytest
import torch

import rubiksnet_b200 as rb
from rubiksnet_b200 import shiftlib
from rubiksnet_b200.shiftlib.rubiks2d import primitive as r2
from rubiksnet_b200.shiftlib.rubiks3d import primitive as r3


def test_public_names():
    for n in ("RubiksShift2D", "RubiksShift3D", "RubiksShiftBase", "rubiks_shift_3d", "rubiks_shift_3d_forward",
              "rubiks_shift_3d_backward", "rubiks2d", "rubiks2d_forward", "rubiks2d_backward"):
        assert hasattr(shiftlib, n), n
    assert r3.RubiksShift3DFunc.__name__ == "RubiksShift3DFunc" and r3.RubiksShift3DFunction is r3.RubiksShift3DFunc
    assert r2.VFS2DFunc.__name__ == "VFS2DFunc"
    for n in ("rubiks2d_forward", "rubiks2d_backward", "rubiks_shift_3d_forward_float",
              "rubiks_shift_3d_forward_double", "rubiks_shift_3d_backward_float", "rubiks_shift_3d_backward_double"):
        assert callable(getattr(rb.rubiksnet_cuda, n)), n


def test_output_shapes():
    x = torch.zeros(2, 8, 4, 56, 56)
    assert r3.compute_output_shape(x, (1, 2, 2), (0, 0, 0)) == (2, 8, 4, 28, 28)
    assert r3.compute_output_shape(x, 1, 0) == (2, 8, 4, 56, 56)
    assert r3.compute_output_shape(x, (2, 2, 1), (1, 1, 0)) == (2, 5, 4, 29, 56)
    assert r2.compute_output_shape(torch.zeros(2, 4, 7, 7), 2, 0) == (2, 4, 4, 4)


def test_cpu_tensors_rejected_like_reference():
    x = torch.zeros(1, 2, 3, 4, 4)
    s = torch.zeros(3, 3)
    with pytest.raises(AssertionError, match="only works on CUDA"):
        r3.rubiks_shift_3d_forward(x, s, 1, 0)
    with pytest.raises(AssertionError, match="only works on CUDA"):
        r2.rubiks2d_forward(torch.zeros(1, 3, 4, 4), torch.zeros(2, 3))
    with pytest.raises(AssertionError, match="x must be"):
        r3.rubiks_shift_3d(torch.zeros(1, 3, 4, 4), s)
    with pytest.raises(AssertionError, match="C dim must match"):
        r3.rubiks_shift_3d(x, torch.zeros(3, 5))


def test_modules_and_init_modes():
    torch.manual_seed(0)
    m = shiftlib.RubiksShift3D(16, stride=(1, 2, 2))
    assert m.shift.shape == (3, 16) and float(m.shift.abs().max()) <= 1.0
    assert "shift_channels=16" in repr(m)
    m2 = shiftlib.RubiksShift2D(18, init_shift="group3")
    assert sorted(set(m2.shift[0].tolist())) == [-1.0, 0.0, 1.0]
    m3 = shiftlib.create_3d_from_2d(shiftlib.RubiksShift2D(16, stride=2), init_mode="tsm")
    assert m3.stride == (1, 2, 2) and m3.padding == (0, 0, 0)
    assert m3.shift[0, :2].tolist() == [1.0, 1.0] and m3.shift[0, 2:4].tolist() == [-1.0, -1.0]
    assert float(m3.shift[0, 4:].abs().sum()) == 0.0
    m4 = shiftlib.create_3d_from_2d(shiftlib.RubiksShift2D(16), init_mode="uni0.5")
    assert float(m4.shift[0].abs().max()) <= 0.5
    assert torch.isnan(shiftlib.create_3d_from_2d(shiftlib.RubiksShift2D(8), init_mode="none").shift).all()
    s = shiftlib.init_shift1d_nfold(torch.zeros(1, 16))
    assert s[0, :2].tolist() == [1, 1] and s[0, 2:4].tolist() == [-1, -1]


@pytest.mark.parametrize("tier,params", [("tiny", None), ("large", 8538294)])
def test_rubiksnet_structure(tier, params):
    net = rb.RubiksNet(tier=tier, num_classes=174, num_frames=8)
    keys = list(net.state_dict().keys())
    assert "backbone.conv1.weight" in keys and "new_fc.weight" in keys
    assert "backbone.layer1.0.as3.rubiks3d.shift" in keys and "backbone.layer1.0.shortcut.weight" in keys
    assert "backbone.bn_last.running_var" in keys
    if params:
        assert sum(p.numel() for p in net.parameters()) == params
        assert len(keys) == 675
    assert net.crop_size == 224 and net.scale_size == 256 and net.feature_dim == 8 * net.backbone.width
    net.replace_new_fc(10)
    assert net.new_fc.out_features == 10
    pol = net.backbone.get_optim_policy(shift_lr_mult=0.01)
    assert [g["name"] for g in pol] == ["weight", "bias", "bn", "shift"]
    assert len(pol[3]["params"]) == (51 if tier == "large" else 17) and pol[3]["lr_mult"] == 0.01


def test_rubiksnet_aq_structure():
    net = rb.RubiksNet(tier="tiny", num_classes=11, num_frames=8, variant="rubiks3d-aq")
    keys = set(net.state_dict().keys())
    assert {"backbone.layer0.0.conv2.0.T", "backbone.layer0.0.conv2.0.weight", "backbone.layer0.0.conv2.1.weight",
            "backbone.layer0.0.as3.shift"} <= keys
    assert net.backbone.layer2[0].conv2[0].weight.shape == (54 * 2 // 1 if False else net.backbone.layer2[0].conv2[1].in_channels, 3)


def test_small_tier_has_se():
    net = rb.RubiksNet(tier="small", num_classes=5)
    assert net.backbone.layer1[0].se is not None and net.backbone.layer1[0].se.fc[0].out_features == 6


@pytest.mark.skipif(not os.path.isdir("/root/reference/pretrained"), reason="reference checkpoints not present")
@pytest.mark.parametrize("ckpt", ["ssv2_tiny", "ssv2_large_aq_budget0.125", "kinetics_small"])
def test_strict_checkpoint_loading(ckpt):
    net = rb.RubiksNet.load_pretrained("/root/reference/pretrained/%s.pth.tar" % ckpt)
    assert net.num_frames == 8


def test_attention_taps_match_reference_formula():
    torch.manual_seed(0)
    m = rb.AttentionShift(8, channels=5)
    w = m.weight.detach().double().numpy()
    z = w / (w.std(axis=1, ddof=1, keepdims=True) + 1e-6) / 2.0
    e = np.exp(z - z.max(1, keepdims=True))
    np.testing.assert_allclose(m.taps().detach().numpy(), e / e.sum(1, keepdims=True), rtol=1e-5)
    lazy = rb.AttentionShift(8)
    assert lazy.weight is None and set(lazy.state_dict()) == {"T"}
