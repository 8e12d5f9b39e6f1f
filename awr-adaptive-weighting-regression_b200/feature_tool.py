"""Drop-in for the reference's util/feature_tool.py::FeatureModule (feature_tool.py:10-65),
backed by the fused sm_100a kernels in csrc/head.cu through the C-ABI."""
import torch

from . import _lib as L


def _prep(t, dtype=torch.float32):
    return t.detach().to(dtype).contiguous()


def _head_ws(B, J, device):
    return torch.zeros(4 * B * J + 4, dtype=torch.float32, device=device)


class _Offset2Joint(torch.autograd.Function):
    @staticmethod
    def forward(ctx, offset, img, kernel_size):
        L.require_cuda(offset, img)
        B, C4, F, F2 = offset.shape
        if C4 % 4 or F != F2:
            raise ValueError("offset must be (B,4J,F,F)")
        J = C4 // 4
        H = img.shape[-1]
        if img.shape[1] != 1 or img.shape[-2] != H or H % F:
            raise ValueError("img must be (B,1,H,H) with H a multiple of F")
        pred = offset.detach().contiguous()
        if pred.dtype not in (torch.float32, torch.bfloat16):
            pred = pred.float()
        im = _prep(img)
        uvd = torch.empty(B, J, 3, dtype=torch.float32, device=offset.device)
        ws = _head_ws(B, J, offset.device)
        L.check(L.lib().awr_head_fwd(L.ptr(pred), L.dtype_code(pred), L.ptr(im), None, L.ptr(uvd), None, L.ptr(ws),
                                     B, J, F, H, float(kernel_size), L.stream()), "awr_head_fwd")
        ctx.save_for_backward(pred, im, uvd, ws)
        ctx.ks = float(kernel_size)
        ctx.in_dtype = offset.dtype
        return uvd

    @staticmethod
    def backward(ctx, g_uvd):
        pred, im, uvd, ws = ctx.saved_tensors
        B, C4, F, _ = pred.shape
        J = C4 // 4
        g = _prep(g_uvd)
        dpred = torch.empty(pred.shape, dtype=torch.float32, device=pred.device)
        L.check(L.lib().awr_head_bwd(L.ptr(pred), L.dtype_code(pred), L.ptr(im), None, L.ptr(uvd), L.ptr(ws), L.ptr(g), None,
                                     L.ptr(dpred), B, J, F, im.shape[-1], ctx.ks, 0.0, 0.0, L.stream()), "awr_head_bwd")
        return dpred.to(ctx.in_dtype), None, None


class FeatureModule:
    """Same two methods, argument meaning and return layout as the reference class."""

    def joint2offset(self, jt_uvd, img, kernel_size, feature_size):
        L.require_cuda(jt_uvd, img)
        B, J, _ = jt_uvd.shape
        H = img.shape[-1]
        jt = _prep(jt_uvd)
        im = _prep(img)
        out = torch.empty(B, 4 * J, feature_size, feature_size, dtype=torch.float32, device=jt.device)
        L.check(L.lib().awr_joint2offset(L.ptr(jt), L.ptr(im), L.ptr(out), B, J, int(feature_size), H, float(kernel_size),
                                         L.stream()), "awr_joint2offset")
        return out

    def offset2joint_softmax(self, offset, img, kernel_size):
        return _Offset2Joint.apply(offset, img, kernel_size)


def head_loss_forward(pred, img, uvd_gt, kernel_size, ws=None, uvd_out=None, loss_out=None):
    """Fused forward: UVD + both unweighted SmoothL1 means in ONE kernel. Returns (uvd, loss[2], ws)."""
    B, C4, F, _ = pred.shape
    J = C4 // 4
    dev = pred.device
    ws = _head_ws(B, J, dev) if ws is None else ws
    uvd_out = torch.empty(B, J, 3, dtype=torch.float32, device=dev) if uvd_out is None else uvd_out
    loss_out = torch.empty(2, dtype=torch.float32, device=dev) if loss_out is None else loss_out
    L.check(L.lib().awr_head_fwd(L.ptr(pred), L.dtype_code(pred), L.ptr(img), L.ptr(uvd_gt), L.ptr(uvd_out), L.ptr(loss_out),
                                 L.ptr(ws), B, J, F, img.shape[-1], float(kernel_size), L.stream()), "awr_head_fwd")
    return uvd_out, loss_out, ws


def head_loss_backward(pred, img, uvd_gt, uvd, ws, kernel_size, coord_weight, dense_weight, dpred=None, g_uvd=None,
                       loss_grad=None):
    """Fused backward: d(cw*L_joint + dw*L_dense)/d pred in ONE kernel (fp32 NCHW)."""
    B, C4, F, _ = pred.shape
    J = C4 // 4
    dpred = torch.empty(pred.shape, dtype=torch.float32, device=pred.device) if dpred is None else dpred
    L.check(L.lib().awr_head_bwd(L.ptr(pred), L.dtype_code(pred), L.ptr(img), L.ptr(uvd_gt), L.ptr(uvd), L.ptr(ws), L.ptr(g_uvd),
                                 L.ptr(loss_grad), L.ptr(dpred), B, J, F, img.shape[-1], float(kernel_size), float(coord_weight),
                                 float(dense_weight), L.stream()), "awr_head_bwd")
    return dpred
