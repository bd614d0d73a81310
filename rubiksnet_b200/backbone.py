"""Pre-activation ResNet of RubiksShiftBlocks -- mirror of rubiksnet/backbone.py.

Block (backbone.py:109-135):  o = relu(bn1(x)); sc = x | shortcut(o); o = conv2(o); o = relu(bn2(o));
o = shift(o); [o = se(o)]; o = conv3(o); o += sc.   Module / parameter names follow the reference so
that its checkpoints load with strict=True.
"""
import math

import torch
import torch.nn as nn

from . import fused
from .shiftlib import RubiksShift2D, RubiksShiftBase

__all__ = ["RubiksNetBackbone", "RubiksShiftBlock", "SELayer"]


def _conv1x1(cin, cout, stride=1):
    return nn.Conv2d(cin, cout, kernel_size=1, stride=stride, bias=False)


def _bn(planes):
    bn = nn.BatchNorm2d(planes)
    nn.init.constant_(bn.weight, 1.0)
    nn.init.constant_(bn.bias, 0.0)
    return bn


# CUDA tensors in fp32 / fp16 / bf16 take the fused path (librubiks_b200 BN+ReLU kernels, GEMM 1x1 convs); float64 (the
# reference runs end to end in double, e.g. for gradcheck) and FUSED_BLOCK = False run the plain nn.Module graph
# around the shift kernels (tests use the switch to check that both give the same numbers)
FUSED_BLOCK = True
_FUSED_DTYPES = (torch.float32, torch.float16, torch.bfloat16)


def _use_fused(x):
    return FUSED_BLOCK and x.is_cuda and x.dtype in _FUSED_DTYPES
# identity-shortcut rubiks3d blocks with bf16 activations: one autograd Function, shift fused into conv3 (fused.py)
FUSED_WHOLE_BLOCK = True


class SELayer(nn.Module):
    """Squeeze-and-excitation gate (backbone.py:56-71); only tier "small" uses it."""

    def __init__(self, channel, reduction):
        super().__init__()
        self.avg_pool = nn.AdaptiveAvgPool2d(1)
        self.fc = nn.Sequential(
            nn.Linear(channel, channel // reduction, bias=False), nn.ReLU(inplace=True),
            nn.Linear(channel // reduction, channel, bias=False), nn.Sigmoid())

    def forward(self, x):
        b, c = x.shape[:2]
        gate = self.fc(self.avg_pool(x).view(b, c)).view(b, c, 1, 1)
        return x * gate


class RubiksShiftBlock(nn.Module):
    def __init__(self, in_planes, out_planes, *, stride=1, parent):
        super().__init__()
        mid = int(out_planes * parent.expansion)
        self.relu = nn.ReLU(inplace=True)
        self.bn1 = _bn(in_planes)
        self.conv2 = _conv1x1(in_planes, mid)
        self.bn2 = _bn(mid)
        self.as3 = RubiksShift2D(mid, stride=stride, normalize_grad=parent.normalize_grad,
                                 quantize=parent.quantize, init_shift=parent.init_shift)
        if parent.use_se:
            reduction = 12 if isinstance(parent.use_se, bool) else parent.use_se
            assert reduction > 2, ("SE reduction must > 2", reduction)
            self.se = SELayer(mid, reduction=reduction)
        else:
            self.se = None
        self.conv3 = _conv1x1(mid, out_planes)
        if stride != 1 or in_planes != out_planes:
            self.shortcut = _conv1x1(in_planes, out_planes, stride=stride)
        else:
            self.shortcut = nn.Identity()

    def forward(self, x):
        if _use_fused(x):
            return self._forward_fused(x)
        out = self.relu(self.bn1(x))
        shortcut = x if isinstance(self.shortcut, nn.Identity) else self.shortcut(out)
        out = self.relu(self.bn2(self.conv2(out)))
        out = self.as3(out)
        if self.se is not None:
            out = self.se(out)
        out = self.conv3(out)
        out += shortcut
        return out

    def _forward_fused(self, x):
        """Same arithmetic with librubiks_b200's BN+ReLU passes and GEMM 1x1 convolutions; the residual add is the
        epilogue of the conv3 GEMM.  Identity-shortcut 3D blocks in bf16 run as one fused autograd Function."""
        if FUSED_WHOLE_BLOCK and fused.eval_block_supported(self, x):
            return fused.eval_block(self, x)
        if FUSED_WHOLE_BLOCK and fused.rubiks_block_supported(self, x):
            return fused.rubiks_block(self, x)
        if FUSED_WHOLE_BLOCK and fused.rubiks_down_block_supported(self, x):
            return fused.rubiks_down_block(self, x)
        if FUSED_WHOLE_BLOCK and fused.rubiks_aq_block_supported(self, x):
            return fused.rubiks_aq_block(self, x)
        out = fused.bn_act(x, self.bn1, relu=True)
        if isinstance(self.shortcut, nn.Identity):
            shortcut = x
        else:
            shortcut = fused.conv1x1(out, self.shortcut.weight, stride=self.shortcut.stride[0])
        if isinstance(self.conv2, nn.Sequential):  # rubiks3d-aq: AttentionShift in front of conv2
            out = fused.conv1x1(self.conv2[0](out), self.conv2[1].weight)
        else:
            out = fused.conv1x1(out, self.conv2.weight)
        out = fused.bn_act(out, self.bn2, relu=True)
        out = self.as3(out)
        if self.se is not None:
            out = fused.se_gate(out, self.se)
        return fused.conv1x1(out, self.conv3.weight, residual=shortcut)


class RubiksNetBackbone(nn.Module):
    """conv1 3x3/2 -> layer0 (1 block) -> layer1..4 (repeats, each stage strided by 2 inside its first
    shift) -> bn_last/relu -> avgpool -> fc   (backbone.py:138-200)."""

    def __init__(self, width, repeats, expansion=1, num_classes=1000, use_se=False, quantize=False,
                 normalize_grad=True, init_shift="uniform"):
        super().__init__()
        self.init_shift, self.width, self.expansion = init_shift, width, expansion
        self.use_se, self.quantize, self.normalize_grad = use_se, quantize, normalize_grad
        self.inplanes = width
        self.conv1 = nn.Conv2d(3, width, kernel_size=3, stride=2, padding=1, bias=False)
        self.layer0 = self._make_layer(width, 1, stride=1)
        for i, rep in enumerate(repeats):
            setattr(self, "layer%d" % (i + 1), self._make_layer(width << i, rep, stride=2))
        self.relu = nn.ReLU(inplace=True)
        self.bn_last = _bn(8 * width)
        self.avgpool = nn.AvgPool2d(7, stride=1)
        self.fc = nn.Linear(8 * width, num_classes)
        # parameter init (backbone.py:14-42,172-175)
        for m in self.modules():
            if isinstance(m, nn.Conv2d) and not getattr(m, "skip_global_init", False):
                fan = m.kernel_size[0] * m.kernel_size[1] * m.out_channels
                nn.init.normal_(m.weight, 0, math.sqrt(2.0 / fan))
        self.fc.weight.data.normal_(0, 0.01)

    def _make_layer(self, planes, repeat, stride):
        blocks = []
        for i in range(repeat):
            blocks.append(RubiksShiftBlock(self.inplanes, planes, stride=stride if i == 0 else 1, parent=self))
            self.inplanes = planes
        return nn.Sequential(*blocks)

    def forward(self, x):
        x = fused.stem_conv(self.conv1, x) if FUSED_BLOCK else self.conv1(x)
        packing = _use_fused(x) and x.dtype == torch.bfloat16
        counting = _use_fused(x) and self.training
        if packing:
            fused.begin_step_pack(self)  # every conv-weight image of the network in one launch
        if counting:
            fused.begin_step_counters(self)  # every BatchNorm batch counter in one launch
        try:
            for i in range(5):
                x = getattr(self, "layer%d" % i)(x)
            if _use_fused(x):
                x = self.avgpool(fused.bn_act(x, self.bn_last, relu=True))
            else:
                x = self.avgpool(self.relu(self.bn_last(x)))
        finally:
            if packing:
                fused.end_step_pack(self)
            if counting:
                fused.end_step_counters()
        return self.fc(x.view(x.size(0), -1))

    def get_optim_policy(self, shift_lr_mult=0.01):
        """Parameter groups weight / bias / bn / shift with lr_mult & decay_mult (backbone.py:202-235)."""
        groups = {"weight": [], "bias": [], "bn": [], "shift": []}
        for m in self.modules():
            if isinstance(m, (nn.Conv2d, nn.Conv3d, nn.Linear)):
                ps = list(m.parameters())
                groups["weight"].append(ps[0])
                if len(ps) == 2:
                    groups["bias"].append(ps[1])
            elif isinstance(m, (nn.BatchNorm1d, nn.BatchNorm2d, nn.BatchNorm3d)):
                groups["bn"].extend(m.parameters())
            elif isinstance(m, (RubiksShift2D, RubiksShiftBase)):
                groups["shift"].extend(m.parameters())
            elif len(m._modules) == 0 and len(list(m.parameters())) > 0:
                raise ValueError("New atomic module type: {}. Need to give it a learning policy".format(type(m)))
        return [
            {"params": groups["weight"], "lr_mult": 1, "decay_mult": 1, "name": "weight"},
            {"params": groups["bias"], "lr_mult": 1, "decay_mult": 0, "name": "bias"},
            {"params": groups["bn"], "lr_mult": 1, "decay_mult": 0, "name": "bn"},
            {"params": groups["shift"], "lr_mult": shift_lr_mult, "decay_mult": 0, "name": "shift"},
        ]
