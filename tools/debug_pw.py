"""GPU bring-up check of the tcgen05 pointwise-conv kernel against torch.matmul (fp32 on the same bf16 inputs)."""
import os
import sys

import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from rubiksnet_b200 import _lib  # noqa: E402


def run(NI, K, N, HW, res=False, bn=False):
    torch.manual_seed(0)
    x = torch.randn(NI, K, HW, device="cuda").bfloat16()
    w = (torch.randn(N, K, device="cuda") / K ** 0.5).bfloat16()
    r = torch.randn(NI, N, HW, device="cuda").bfloat16() if res else None
    sc = torch.rand(K, device="cuda") + 0.5 if bn else None
    bi = torch.randn(K, device="cuda") if bn else None
    out = torch.full((NI, N, HW), float("nan"), device="cuda", dtype=torch.bfloat16)
    L = _lib.lib()
    rc = L.rb_pw_conv_forward(_lib.ptr(x), _lib.ptr(w), _lib.ptr(r), _lib.ptr(out), _lib.RB_BF16, NI, K, N, HW,
                              _lib.ptr(sc), _lib.ptr(bi), _lib.stream_handle(x.device))
    if rc:
        return "rc=%d %s" % (rc, L.rb_last_error().decode())
    torch.cuda.synchronize()
    a = x.float()
    if bn:
        a = torch.relu(a * sc[None, :, None] + bi[None, :, None]).bfloat16().float()
    ref = torch.matmul(w.float(), a)
    if res:
        ref = ref + r.float()
    err = (out.float() - ref).abs().max().item()
    nan = torch.isnan(out.float()).sum().item()
    return "max_err=%.4g nan=%d ref_absmax=%.3g" % (err, nan, ref.abs().max().item())


if __name__ == "__main__":
    cases = [(2, 16, 16, 128), (2, 64, 16, 128), (2, 72, 72, 256), (3, 144, 144, 784), (4, 288, 288, 196), (2, 576, 576, 49),
             (2, 54, 108, 196)]
    for swap in ("0", "1", "2", "3"):
        os.environ["RB_PW_SWAP"] = swap
        for c in cases[:3] if swap != "0" else cases:
            print("swap", swap, c, run(*c), flush=True)
    os.environ["RB_PW_SWAP"] = "0"
    print("res", run(4, 288, 288, 196, res=True))
    print("bn ", run(4, 72, 72, 3136, bn=True))


def run_wgrad(NI, K, N, HW, bn=False):
    torch.manual_seed(1)
    x = torch.randn(NI, K, HW, device="cuda").bfloat16()
    g = torch.randn(NI, N, HW, device="cuda").bfloat16()
    sc = torch.rand(K, device="cuda") + 0.5 if bn else None
    bi = torch.randn(K, device="cuda") if bn else None
    dw = torch.full((N, K), float("nan"), device="cuda")
    L = _lib.lib()
    nbytes = L.rb_pw_conv_wgrad_workspace_bytes(NI, K, N, HW)
    ws = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
    rc = L.rb_pw_conv_wgrad(_lib.ptr(g), _lib.ptr(x), _lib.ptr(dw), _lib.RB_BF16, NI, K, N, HW, _lib.ptr(sc), _lib.ptr(bi),
                            _lib.ptr(ws), nbytes, _lib.stream_handle(x.device))
    if rc:
        return "rc=%d %s" % (rc, L.rb_last_error().decode())
    torch.cuda.synchronize()
    a = x.float()
    if bn:
        a = torch.relu(a * sc[None, :, None] + bi[None, :, None]).bfloat16().float()
    ref = torch.einsum("inp,ikp->nk", g.float(), a)
    err = (dw - ref).abs().max().item()
    return "max_err=%.4g nan=%d ref_absmax=%.3g" % (err, torch.isnan(dw).sum().item(), ref.abs().max().item())


def run_shift(N, T, C, H, W, Cout, wgrad=False):
    from rubiksnet_b200.shiftlib.rubiks3d.primitive import rubiks_shift_3d_forward
    torch.manual_seed(2)
    x = torch.randn(N, T, C, H, W, device="cuda").bfloat16()
    shift = torch.rand(3, C, device="cuda") * 3 - 1.5
    shift[:, 0] = 0.0
    shift[0, 1] = -1.0
    w = (torch.randn(Cout, C, device="cuda") / C ** 0.5).bfloat16()
    L = _lib.lib()
    s_ref = rubiks_shift_3d_forward(x, shift, (1, 1, 1), 0).view(N * T, C, H * W)
    if not wgrad:
        r = torch.randn(N * T, Cout, H * W, device="cuda").bfloat16()
        out = torch.full((N * T, Cout, H * W), float("nan"), device="cuda", dtype=torch.bfloat16)
        rc = L.rb_shift3d_pw_conv_forward(_lib.ptr(x), _lib.ptr(shift), _lib.ptr(w), _lib.ptr(r), _lib.ptr(out), _lib.RB_BF16,
                                          _lib.RB_F32, N, T, C, H, W, Cout, _lib.stream_handle(x.device))
        if rc:
            return "rc=%d %s" % (rc, L.rb_last_error().decode())
        torch.cuda.synchronize()
        ref = torch.matmul(w.float(), s_ref.float()) + r.float()
        return "max_err=%.4g nan=%d ref_absmax=%.3g" % ((out.float() - ref).abs().max().item(),
                                                        torch.isnan(out.float()).sum().item(), ref.abs().max().item())
    g = torch.randn(N * T, Cout, H * W, device="cuda").bfloat16()
    dw = torch.full((Cout, C), float("nan"), device="cuda")
    nbytes = L.rb_pw_conv_wgrad_workspace_bytes(N * T, C, Cout, H * W)
    ws = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
    rc = L.rb_shift3d_pw_conv_wgrad(_lib.ptr(g), _lib.ptr(x), _lib.ptr(shift), _lib.ptr(dw), _lib.RB_BF16, _lib.RB_F32, N, T, C,
                                    H, W, Cout, _lib.ptr(ws), nbytes, _lib.stream_handle(x.device))
    if rc:
        return "rc=%d %s" % (rc, L.rb_last_error().decode())
    torch.cuda.synchronize()
    ref = torch.einsum("inp,ikp->nk", g.float(), s_ref.float())
    return "max_err=%.4g nan=%d ref_absmax=%.3g" % ((dw - ref).abs().max().item(), torch.isnan(dw).sum().item(),
                                                    ref.abs().max().item())


if __name__ == "__main__":
    for c in [(2, 16, 16, 128), (2, 72, 72, 256), (3, 144, 144, 784), (8, 288, 288, 196), (4, 576, 576, 49), (2, 54, 108, 196),
              (2, 72, 144, 3136)]:
        print("wgrad", c, run_wgrad(*c), flush=True)
    print("wgrad bn", run_wgrad(4, 72, 72, 3136, bn=True), flush=True)
    print("wgrad bn", run_wgrad(8, 288, 288, 196, bn=True), flush=True)
    for c in [(1, 8, 16, 14, 14, 16), (2, 8, 72, 56, 56, 72), (2, 8, 288, 14, 14, 288), (1, 8, 576, 7, 7, 576), (1, 4, 144, 28, 28, 144)]:
        print("shift+conv fwd", c, run_shift(*c), flush=True)
        print("shift+conv wgrad", c, run_shift(*c, wgrad=True), flush=True)
