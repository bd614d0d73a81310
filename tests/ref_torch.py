"""Independent differentiable PyTorch formulation of the shifts (CPU, float64) used to cross-check the
C oracle: floor() carries no gradient, the remainder does, so autograd yields the adjoint (input
gradient) and -- for non-integer shifts -- the exact shift gradient.  Test infrastructure only."""
import torch


def _axis(n_out, n_in, stride, pad, s):
    """index [C, n_out, 2], weight [C, 1, 2] (differentiable), mask [C, n_out, 2] for one axis."""
    f = torch.floor(s.detach().float()).to(torch.long) if s.dtype == torch.float32 else torch.floor(s.detach()).to(torch.long)
    r = s - f.to(s.dtype)
    base = torch.arange(n_out) * stride - pad
    idx = base[None, :, None] + f[:, None, None] + torch.arange(2)[None, None, :]
    mask = (idx >= 0) & (idx < n_in)
    w = torch.stack([1 - r, r], dim=-1)[:, None, :]
    return idx.clamp(0, max(n_in - 1, 0)), w, mask


def shift3d(x, shift, stride, padding):
    N, T, C, H, W = x.shape
    To, Ho, Wo = [(n + 2 * p - 1) // s + 1 for n, s, p in zip((T, H, W), stride, padding)]
    it, wt, mt = _axis(To, T, stride[0], padding[0], shift[0])
    ih, wh, mh = _axis(Ho, H, stride[1], padding[1], shift[1])
    iw, ww, mw = _axis(Wo, W, stride[2], padding[2], shift[2])
    xc = x.permute(2, 0, 1, 3, 4)  # C,N,T,H,W
    out = 0
    cidx = torch.arange(C)[:, None, None, None]
    for a in range(2):
        for b in range(2):
            for d in range(2):
                g = xc[cidx, :, it[:, :, a][:, :, None, None], ih[:, :, b][:, None, :, None],
                       iw[:, :, d][:, None, None, :]]  # C,To,Ho,Wo,N
                m = (mt[:, :, a][:, :, None, None] & mh[:, :, b][:, None, :, None] & mw[:, :, d][:, None, None, :])
                wgt = wt[:, :, a][:, :, None, None] * wh[:, :, b][:, :, None, None].transpose(1, 2) * \
                    ww[:, :, d][:, :, None, None].permute(0, 2, 3, 1)
                out = out + g * (m.to(x.dtype) * wgt)[..., None]
    return out.permute(4, 1, 0, 2, 3)  # N,To,C,Ho,Wo


def shift2d(x, shift, stride, padding):
    N, C, H, W = x.shape
    s3 = torch.cat([torch.zeros(1, C, dtype=shift.dtype), shift], dim=0)
    out = shift3d(x[:, None], s3, (1, *stride), (0, *padding))
    return out[:, 0]
