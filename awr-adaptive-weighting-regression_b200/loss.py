"""Drop-in for the reference's model/loss.py::My_SmoothL1Loss (loss.py:3-25): Huber(delta=0.01), mean."""
import torch

from . import _lib as L


class _Huber(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, y):
        L.require_cuda(x, y)
        xf = x.detach().float().contiguous()
        yf = y.detach().float().contiguous()
        ws = torch.zeros(2 * L.HUBER_MAX_BLOCKS + 4, dtype=torch.float32, device=x.device)
        out = torch.empty((), dtype=torch.float32, device=x.device)
        L.check(L.lib().awr_huber_fwd(L.ptr(xf), L.ptr(yf), xf.numel(), L.ptr(ws), L.ptr(out), L.stream()), "awr_huber_fwd")
        ctx.save_for_backward(xf, yf)
        ctx.x_dtype = x.dtype
        ctx.need_y = y.requires_grad
        return out

    @staticmethod
    def backward(ctx, g):
        xf, yf = ctx.saved_tensors
        gg = g.detach().float().contiguous()
        dx = torch.empty_like(xf)
        L.check(L.lib().awr_huber_bwd(L.ptr(xf), L.ptr(yf), xf.numel(), L.ptr(gg), L.ptr(dx), L.stream()), "awr_huber_bwd")
        dx = dx.to(ctx.x_dtype)
        return dx, (-dx if ctx.need_y else None)


class My_SmoothL1Loss(torch.nn.Module):
    def forward(self, x, y):
        assert x.shape == y.shape           # loss.py:10
        return _Huber.apply(x, y)
