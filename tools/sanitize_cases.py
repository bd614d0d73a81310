"""Small invocations of every hand-written kernel family for compute-sanitizer (tools/sanitize.sh):
k_pw_conv (plain / +residual / bn+relu producer / dgrad, on 14x14, 7x7 and 28x28 maps), k_pw_wgrad (+ reduce),
the fused shift+conv launches, strip / tiled / generic 3D shift forward + backward, BatchNorm passes, 2D shift and
AttentionShift.  Sizes are tiny: the sanitizer slows kernels down 10-100x."""
import os
import sys

import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from rubiksnet_b200 import _lib, ops  # noqa: E402
from rubiksnet_b200.attention_shift import attention_shift_mix  # noqa: E402
from rubiksnet_b200.shiftlib.rubiks2d.primitive import rubiks2d  # noqa: E402
from rubiksnet_b200.shiftlib.rubiks3d.primitive import rubiks_shift_3d  # noqa: E402

BF = torch.bfloat16


def pw_cases():
    for c, h, ni in ((288, 14, 8), (144, 28, 8), (54, 7, 8)):
        x = torch.randn(ni, c, h, h, device="cuda").to(BF)
        g = torch.randn(ni, c, h, h, device="cuda").to(BF)
        w = torch.randn(c, c, device="cuda") / c ** 0.5
        w_nk, w_kn = ops.pw_weight_pack(w)
        sb = torch.stack([torch.rand(c, device="cuda") + 0.5, torch.randn(c, device="cuda")], dim=1).contiguous()
        shift = torch.rand(3, c, device="cuda") * 2 - 1
        ops.pw_conv(x, w_nk)
        ops.pw_conv(x, w_nk, residual=g)
        ops.pw_conv(x, w_nk, in_scale_bias=sb)
        ops.pw_conv(g, w_kn)
        ops.pw_conv(x, w, transposed=True)
        ops.pw_conv(x, w_nk, residual=g, stats=True)
        ops.pw_conv_wgrad(g, x)
        ops.pw_conv_wgrad(g, x, in_scale_bias=sb)
        ops.shift3d_pw_conv(x, shift, w_nk, g, 8)
        ops.shift3d_pw_conv_wgrad(g, x, shift, 8)
        torch.cuda.synchronize()
        print("pw ok", c, h, flush=True)


def shift_cases():
    for impl in (_lib.RB_IMPL_STRIP, _lib.RB_IMPL_TILED, _lib.RB_IMPL_GENERIC):
        for dtype in (torch.float32, BF):
            for c, h, stride in ((24, 14, 1), (16, 28, 2), (40, 7, 1)):
                if impl == _lib.RB_IMPL_STRIP and stride != 1:
                    continue
                _lib.set_impl(impl)
                x = torch.randn(2, 8, c, h, h, device="cuda").to(dtype).requires_grad_()
                s = (torch.rand(3, c, device="cuda") * 3 - 1.5).requires_grad_()
                s.data[:, 0] = 0.0
                s.data[:, 1] = 1.0
                out = rubiks_shift_3d(x, s, stride=(1, stride, stride))
                out.backward(torch.randn_like(out))
                torch.cuda.synchronize()
        print("shift ok impl", impl, flush=True)
    _lib.set_impl(_lib.RB_IMPL_AUTO)


def misc_cases():
    for dtype in (torch.float32, BF):
        x = torch.randn(16, 24, 14, 14, device="cuda").to(dtype).requires_grad_()
        s = (torch.rand(2, 24, device="cuda") * 2 - 1).requires_grad_()
        rubiks2d(x, s, stride=1).float().sum().backward()
        taps = torch.softmax(torch.randn(24, 3, device="cuda"), 1).requires_grad_()
        attention_shift_mix(x.detach().requires_grad_(), taps, 8).float().sum().backward()
        bn_g, bn_b = torch.rand(24, device="cuda") + 0.5, torch.randn(24, device="cuda")
        rm, rv = torch.zeros(24, device="cuda"), torch.ones(24, device="cuda")
        xd = x.detach()
        y, mi, sb = ops.bn_forward(xd, bn_g, bn_b, rm, rv, True, 0.1, 1e-5, relu=True, apply=True)
        ops.bn_backward(xd, torch.randn_like(xd), xd, bn_g, mi, sb, True, relu=True)
        from rubiksnet_b200.attention_shift import attention_mix_backward, attention_mix_forward
        attention_mix_forward(xd, taps.detach(), 8, in_scale_bias=sb)  # bn1 -> relu folded into the temporal mix
        attention_mix_backward(xd, taps.detach(), torch.randn_like(xd), 8, in_scale_bias=sb)
        torch.cuda.synchronize()
    # tcgen05 kind::tf32 kernel: vector and scalar paths, producer / epilogue folds, split contraction; SE kernels; pack fence
    for ni, k, n, h in ((4, 54, 54, 14), (2, 216, 432, 7), (2, 432, 72, 7), (3, 20, 12, 5)):
        x = torch.randn(ni, k, h, h, device="cuda")
        w = torch.randn(n, k, device="cuda") / k ** 0.5
        res = torch.randn(ni, n, h, h, device="cuda")
        isb = torch.stack([torch.rand(k, device="cuda") + 0.5, torch.randn(k, device="cuda")], dim=1).contiguous()
        osb = torch.stack([torch.rand(n, device="cuda") + 0.5, torch.randn(n, device="cuda")], dim=1).contiguous()
        ops.pw_conv_f32(x, w)
        ops.pw_conv_f32(x, w, residual=res, in_scale_bias=isb, out_scale_bias=osb, relu=True, resident=True)
        torch.cuda.synchronize()
    # conv1 path (im2col in bf16 / fp32), GPU input pipeline, SE kernels, 2D shift on the tiled kernel (stride 2)
    from rubiksnet_b200 import backbone, fused
    from rubiksnet_b200.evaluate import frames_to_clip
    conv = torch.nn.Conv2d(3, 24, 3, stride=2, padding=1, bias=False).cuda()
    xi = torch.randn(2, 3, 32, 32, device="cuda")
    with torch.autocast("cuda", dtype=BF):
        y = fused.stem_conv(conv, xi)
    y.float().sum().backward()
    with torch.no_grad():
        fused.stem_conv(conv, xi)
    frames_to_clip(torch.randint(0, 256, (2, 9, 11, 6), dtype=torch.uint8, device="cuda"))
    se = backbone.SELayer(24, reduction=12).cuda()
    xs = torch.randn(4, 24, 7, 7, device="cuda").to(BF).requires_grad_()
    fused.se_gate(xs, se).float().sum().backward()
    for dtype in (torch.float32, BF):
        x2 = torch.randn(8, 6, 28, 28, device="cuda").to(dtype).requires_grad_()
        s2 = (torch.rand(2, 6, device="cuda") * 2 - 1).requires_grad_()
        s2.data[:, 0] = 1.0
        rubiks2d(x2, s2, stride=2).float().sum().backward()
    torch.cuda.synchronize()
    print("misc ok", flush=True)


def bn_cases():
    """BatchNorm passes on every dispatch path: the halves kernel of the 16-bit types (planes of a multiple of 4 elements,
    vectors that straddle two planes, a 4-element tail), the per-element kernel (odd planes, fp32), the channel-resident
    forward, with / without the shortcut gradient."""
    for dtype in (BF, torch.float16, torch.float32):
        for ni, c, h in ((16, 24, 14), (3, 5, 6), (2, 7, 7), (4, 8, 28), (40, 6, 14)):
            x = torch.randn(ni, c, h, h, device="cuda").to(dtype)
            g, b = torch.rand(c, device="cuda") + 0.5, torch.randn(c, device="cuda")
            rm, rv = torch.zeros(c, device="cuda"), torch.ones(c, device="cuda")
            for relu in (True, False):
                y, mi, sb = ops.bn_forward(x, g, b, rm, rv, True, 0.1, 1e-5, relu=relu, apply=True)
                ops.bn_forward(x, g, b, rm, rv, True, 0.1, 1e-5, relu=relu, apply=False)
                ops.bn_apply(x, sb, relu=relu)
                dy = torch.randn_like(x)
                ops.bn_backward(x, dy, None, g, mi, sb, True, relu=relu)
                ops.bn_backward(x, dy, torch.randn_like(x), g, mi, sb, True, relu=relu)
            ops.bn_forward(x, g, b, rm, rv, False, 0.1, 1e-5, relu=True, apply=True)  # eval coefficients
            torch.cuda.synchronize()
    print("bn ok", flush=True)


if __name__ == "__main__":
    which = sys.argv[1:] or ["pw", "shift", "misc", "bn"]
    if "bn" in which:
        bn_cases()
    if "pw" in which:
        pw_cases()
    if "shift" in which:
        shift_cases()
    if "misc" in which:
        misc_cases()
