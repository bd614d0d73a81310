"""RubiksNet(tier, num_classes, num_frames=8, variant) -- mirror of rubiksnet/models.py."""
import os

import torch
import torch.nn as nn

from .attention_shift import AttentionShift
from .backbone import RubiksNetBackbone
from .shiftlib import RubiksShift2D, RubiksShift3D
from .utils import make_tuple

__all__ = ["RubiksNet"]

# tier -> (width, repeats, use_se)   (models.py:28-43)
_TIERS = {
    "tiny": (54, [3, 4, 6, 3], False),
    "small": (72, [3, 4, 6, 3], True),
    "medium": (72, [3, 4, 23, 3], False),
    "large": (72, [3, 8, 36, 3], False),
}
_VARIANTS = ("rubiks3d", "rubiks3d-aq")


class _Rubiks3DWrap(nn.Module):
    """Views the backbone's [N*T,C,H,W] activations as [N,T,C,H,W] around a fresh RubiksShift3D that
    takes its (H,W) stride / padding from the block's 2D shift (models.py:128-145)."""

    def __init__(self, rubiks2d, n_segment=8):
        super().__init__()
        assert isinstance(rubiks2d, RubiksShift2D)
        self.rubiks3d = RubiksShift3D(rubiks2d.num_channels, stride=(1, *make_tuple(rubiks2d.stride, 2)),
                                      padding=(0, *make_tuple(rubiks2d.padding, 2)))
        self.n_segment = n_segment

    def forward(self, x):
        nt, c, h, w = x.size()
        out = self.rubiks3d(x.view(nt // self.n_segment, self.n_segment, c, h, w))
        n, t, c, h, w = out.size()
        return out.view(n * t, c, h, w)


class RubiksNet(nn.Module):
    def __init__(self, tier, num_classes, num_frames=8, variant="rubiks3d"):
        super().__init__()
        assert tier in _TIERS
        assert variant in _VARIANTS
        self.num_frames, self.tier, self.variant = num_frames, tier, variant
        width, repeats, use_se = _TIERS[tier]
        self.backbone = RubiksNetBackbone(width=width, repeats=repeats, num_classes=num_classes, use_se=use_se)
        self._prepare_backbone()
        self.feature_dim = getattr(self.backbone, self.backbone.last_layer_name).in_features
        setattr(self.backbone, self.backbone.last_layer_name, nn.Identity())
        self.new_fc = nn.Linear(self.feature_dim, num_classes)

    @classmethod
    def load_pretrained(cls, ckpt_path):
        """Checkpoint = {model, tier, variant, num_classes, num_frames} (models.py:52-62); strict load."""
        ckpt = torch.load(os.path.expanduser(ckpt_path), map_location="cpu")
        net = cls(tier=ckpt["tier"], num_classes=ckpt["num_classes"], num_frames=ckpt["num_frames"],
                  variant=ckpt["variant"])
        net.load_state_dict(ckpt["model"])
        return net

    def replace_new_fc(self, num_classes):
        self.new_fc = nn.Linear(self.feature_dim, num_classes)

    def _prepare_backbone(self):
        net = self.backbone
        for stage in ("layer0", "layer1", "layer2", "layer3", "layer4"):
            for block in getattr(net, stage):
                if self.variant == "rubiks3d":
                    block.as3 = _Rubiks3DWrap(block.as3, n_segment=self.num_frames)
                else:
                    # AttentionShift before conv2; weight created eagerly (the reference needs a dummy CUDA
                    # forward for that, models.py:100-104), the block keeps its 2D spatial shift
                    block.conv2 = nn.Sequential(
                        AttentionShift(self.num_frames, channels=block.conv2.in_channels), block.conv2)
        net.last_layer_name = "fc"
        self.input_size = 224
        self.input_mean = [0.485, 0.456, 0.406]
        self.input_std = [0.229, 0.224, 0.225]
        net.avgpool = nn.AdaptiveAvgPool2d(1)

    def forward(self, input):
        feats = self.backbone(input.view((-1, 3) + input.size()[-2:]))
        logits = self.new_fc(feats)
        logits = logits.view((-1, self.num_frames) + logits.size()[1:])
        return logits.mean(dim=1)

    @property
    def crop_size(self):
        return self.input_size

    @property
    def scale_size(self):
        return self.input_size * 256 // 224
