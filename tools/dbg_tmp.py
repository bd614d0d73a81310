import torch, sys
sys.path.insert(0, '/root/repo'); sys.path.insert(0,'/root/repo/tests')
from rubiksnet_b200 import ops
from rubiksnet_b200.shiftlib.rubiks3d.primitive import rubiks_shift_3d_forward
BF=torch.bfloat16
n,t,c,h,w,cout=2,8,576,7,7,576
torch.manual_seed(3)
x = torch.randn(n * t, c, h, w, device="cuda").to(BF)
torch.rand(3, c, device="cuda")
shift = torch.rand(3, c, device="cuda") * 6 - 3
g = torch.randn(n * t, cout, h, w, device="cuda").to(BF)
shifted = rubiks_shift_3d_forward(x.view(n, t, c, h, w), shift, (1, 1, 1), 0).view(n * t, c, h, w)
dw = ops.shift3d_pw_conv_wgrad(g, x, shift, t)
ref = torch.einsum("inp,ikp->nk", g.float().flatten(2), shifted.float().flatten(2))
d=(dw-ref).abs()
print("max err", d.max().item(), "ref max", ref.abs().max().item(), "bad cols", (d.max(0).values>0.05).sum().item(), "bad rows", (d.max(1).values>0.05).sum().item())
print((d.max(0).values>0.05).nonzero().flatten()[:40])
dw2 = ops.pw_conv_wgrad(g, shifted)
print("plain wgrad on shifted:", (dw2-ref).abs().max().item())
