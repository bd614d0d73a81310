"""Helper of tests/test_gpu_dropin.py (run as a subprocess): imports the UNMODIFIED reference `rubiksnet` package from
baseline/_ref with either its own CUDA extension (--native reference) or this repo's `rubiksnet_cuda` stand-in
(--native b200: dropin/ goes in front of sys.path, so `import rubiksnet_cuda` inside the reference resolves to
librubiks_b200), runs the recipe of the reference's scripts/test_installation.py:6-10 plus one backward pass, and
saves prediction, loss and a few gradients."""
import argparse
import contextlib
import io
import os
import sys

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--native", choices=["reference", "b200"], required=True)
    ap.add_argument("--tier", default="large")
    ap.add_argument("--out", required=True)
    a = ap.parse_args()
    import torch
    if a.native == "b200":
        sys.path.insert(0, REPO)
        sys.path.insert(0, os.path.join(REPO, "dropin"))  # in front of baseline/_ref: shadows the reference's .so
        sys.path.insert(2, os.path.join(REPO, "baseline", "_ref"))
    else:
        sys.path.insert(0, os.path.join(REPO, "baseline", "_ref"))
    import rubiksnet_cuda
    impl = getattr(rubiksnet_cuda, "NATIVE_IMPL", "reference extension")
    assert (impl == "librubiks_b200") == (a.native == "b200"), (impl, rubiksnet_cuda.__file__)
    with contextlib.redirect_stdout(io.StringIO()):
        from rubiksnet.models import RubiksNet  # the reference's Python package, untouched
    import rubiksnet
    assert os.path.realpath(rubiksnet.__file__).startswith(os.path.realpath(os.path.join(REPO, "baseline", "_ref")))
    torch.backends.cuda.matmul.allow_tf32 = torch.backends.cudnn.allow_tf32 = False
    torch.manual_seed(0)
    num_frames = 8
    # scripts/test_installation.py:6-10
    net = RubiksNet(tier=a.tier, num_classes=42, num_frames=num_frames)
    net.cuda()
    video = torch.randn((2, num_frames, 3, 224, 224), device="cuda")
    prediction = net(video)
    # ... plus one backward pass through every shift layer
    loss = torch.nn.functional.cross_entropy(prediction, torch.tensor([1, 40], device="cuda"))
    loss.backward()
    named = dict(net.named_parameters())
    keys = ["new_fc.weight", "backbone.conv1.weight", "backbone.layer1.0.conv3.weight", "backbone.layer3.5.conv2.weight",
            "backbone.layer0.0.as3.rubiks3d.shift", "backbone.layer2.0.as3.rubiks3d.shift",
            "backbone.layer3.7.as3.rubiks3d.shift", "backbone.layer4.2.as3.rubiks3d.shift"]
    out = {"prediction": prediction.detach().cpu().numpy(), "loss": np.float64(loss.item()), "impl": np.array(impl)}
    for k in keys:
        if k in named and named[k].grad is not None:
            out["grad:" + k] = named[k].grad.detach().cpu().numpy()
    np.savez(a.out, **out)
    print("dropin_check ok:", impl, "loss=%.6f" % loss.item())


if __name__ == "__main__":
    main()
