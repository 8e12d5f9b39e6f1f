"""GPU numerics of the tcgen05 implicit-GEMM convolution kernels (through the C-ABI) vs plain PyTorch fp32 on the same
bf16-rounded operands.  Tolerance: fp32 accumulation of bf16 products, output rounded to bf16 -> 1e-2 relative to the
tensor's max (bf16 has 8 mantissa bits: 2^-8 = 3.9e-3 per rounding)."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _nhwc(t):          # (N,C,H,W) fp32 -> NHWC bf16 cuda
    return t.permute(0, 2, 3, 1).contiguous().to(torch.bfloat16).cuda()


def _from_nhwc(t):     # NHWC bf16 cuda -> (N,C,H,W) fp32 cpu
    return t.float().cpu().permute(0, 3, 1, 2).contiguous()


def _rb(t):
    return t.to(torch.bfloat16).float()


def _conv_tc(x, w_phys, bias, out, N, Hi, Wi, Ck, Ho, Wo, Cn, k, stride, pad, transposed, w_sk, w_sn, w_tap, out_mode=0, n_valid=0, acc=0, stats=None):
    from awr_b200 import _lib as L
    L.check(L.lib().awr_conv_tc(x.data_ptr(), w_phys.data_ptr(), None if bias is None else bias.data_ptr(), out.data_ptr(),
                                None if stats is None else stats.data_ptr(), N, Hi, Wi, Ck,
                                Ho, Wo, Cn, k, k, stride, pad, transposed, w_sk, w_sn, w_tap, out_mode, n_valid, acc, L.stream()), "awr_conv_tc")
    torch.cuda.synchronize()


def _wgrad_tc(pw, ga, dW, N, Hc, Wc, Cp, Hf, Wf, Cg, k, stride, pad, s_p, s_g, w_tap):
    """dW: fp32 (zero-filled).  In the bit-reproducible build the kernel accumulates into awr_acc_t slots that awr_grad_acc_finalize folds
    into the fp32 array -- the same two calls engine.Plan makes."""
    from awr_b200 import _lib as L
    tgt = L.acc_zeros(dW.numel(), dW.device) if L.deterministic() else dW
    L.check(L.lib().awr_conv_wgrad_tc(pw.data_ptr(), ga.data_ptr(), tgt.data_ptr(), N, Hc, Wc, Cp, Hf, Wf, Cg, k, k, stride, pad, s_p, s_g, w_tap,
                                      L.stream()), "awr_conv_wgrad_tc")
    if L.deterministic():
        L.check(L.lib().awr_grad_acc_finalize(tgt.data_ptr(), dW.data_ptr(), dW.numel(), L.stream()), "awr_grad_acc_finalize")
        torch.cuda.synchronize()
        assert int(tgt.abs().sum()) == 0            # slots are re-armed for the next step


CONV_CASES = [  # N, Cin, Cout, H, k, stride, pad
    (2, 64, 64, 16, 3, 1, 1),
    (3, 128, 256, 8, 1, 1, 0),
    (2, 64, 128, 32, 3, 2, 1),
    (2, 64, 128, 32, 1, 2, 0),
    (2, 128, 128, 64, 3, 1, 1),
    (5, 256, 512, 4, 3, 1, 1),
    (1, 64, 64, 128, 3, 1, 1),
]


@pytest.mark.parametrize("case", CONV_CASES, ids=lambda c: "N%d_%dto%d_H%d_k%ds%dp%d" % c)
def test_conv2d_fprop_dgrad(case):
    N, Ci, Co, H, k, s, pad = case
    g = torch.Generator().manual_seed(sum(case))
    x = _rb(torch.randn(N, Ci, H, H, generator=g))
    w = _rb(torch.randn(Co, Ci, k, k, generator=g) / (Ci * k * k) ** 0.5)
    b = torch.randn(Co, generator=g)
    Ho = (H + 2 * pad - k) // s + 1
    xr = x.clone().requires_grad_(True)
    ref = F.conv2d(xr, w, b, stride=s, padding=pad)
    gy = _rb(torch.randn(ref.shape, generator=g))
    ref.backward(gy)
    w_phys = w.permute(2, 3, 0, 1).contiguous().to(torch.bfloat16).cuda()       # [kh][kw][Co][Ci]
    y = torch.full((N, Ho, Ho, Co), float("nan"), dtype=torch.bfloat16, device="cuda")
    from awr_b200 import _lib as L
    stats = L.acc_zeros(2 * Co, "cuda")
    _conv_tc(_nhwc(x), w_phys, b.cuda(), y, N, H, H, Ci, Ho, Ho, Co, k, s, pad, 0, 1, Ci, Co * Ci, stats=stats)
    stats = L.acc_to_float(stats)
    got = _from_nhwc(y)
    assert (got - ref.detach()).abs().max().item() < 1e-2 * ref.abs().max().item()
    # fused BatchNorm statistics == per-channel sum / sum of squares of the stored (bf16) output
    s1, s2 = got.double().sum(dim=(0, 2, 3)), (got.double() ** 2).sum(dim=(0, 2, 3))
    assert torch.allclose(stats[:Co].cpu().double(), s1, rtol=1e-4, atol=1e-2) and torch.allclose(stats[Co:].cpu().double(), s2, rtol=1e-4, atol=1e-2)
    # dgrad: dx = conv_transpose(dy, w)  -> contraction over Cout, MN-major B from the same weight buffer
    dx = torch.full((N, H, H, Ci), float("nan"), dtype=torch.bfloat16, device="cuda")
    _conv_tc(_nhwc(gy), w_phys, None, dx, N, Ho, Ho, Co, H, H, Ci, k, s, pad, 1, Ci, 1, Co * Ci)
    gdx = _from_nhwc(dx)
    assert (gdx - xr.grad).abs().max().item() < 1e-2 * xr.grad.abs().max().item()
    # wgrad: contraction over pixels, both operands MN-major
    ref_dw = torch.autograd.grad(F.conv2d(x.clone().requires_grad_(False), wr := w.clone().requires_grad_(True), None, stride=s, padding=pad), wr, gy)[0]
    dW = torch.zeros(k, k, Co, Ci, dtype=torch.float32, device="cuda")
    _wgrad_tc(_nhwc(gy), _nhwc(x), dW, N, Ho, Ho, Co, H, H, Ci, k, s, pad, Ci, 1, Co * Ci)
    got_dw = dW.cpu().permute(2, 3, 0, 1)
    assert (got_dw - ref_dw).abs().max().item() < 2e-3 * ref_dw.abs().max().item() + 1e-5     # fp32 accumulate of exact bf16 products
    # accumulate mode adds into the existing tensor
    _conv_tc(_nhwc(gy), w_phys, None, dx, N, Ho, Ho, Co, H, H, Ci, k, s, pad, 1, Ci, 1, Co * Ci, acc=1)
    assert (_from_nhwc(dx) - 2 * xr.grad).abs().max().item() < 2e-2 * xr.grad.abs().max().item()


DECONV_CASES = [(2, 128, 64, 8), (2, 512, 256, 8), (3, 256, 256, 16), (1, 64, 64, 64)]   # N, Cin, Cout, Hin  (k4 s2 p1)


@pytest.mark.parametrize("case", DECONV_CASES, ids=lambda c: "N%d_%dto%d_H%d" % c)
def test_conv_transpose2d_fprop_dgrad(case):
    N, Ci, Co, H = case
    g = torch.Generator().manual_seed(sum(case))
    x = _rb(torch.randn(N, Ci, H, H, generator=g))
    w = _rb(torch.randn(Ci, Co, 4, 4, generator=g) / (Ci * 4) ** 0.5)           # ConvTranspose2d IOHW
    xr = x.clone().requires_grad_(True)
    ref = F.conv_transpose2d(xr, w, None, stride=2, padding=1)
    gy = _rb(torch.randn(ref.shape, generator=g))
    ref.backward(gy)
    Ho = 2 * H
    w_phys = w.permute(2, 3, 1, 0).contiguous().to(torch.bfloat16).cuda()       # [kh][kw][Co][Ci]
    y = torch.full((N, Ho, Ho, Co), float("nan"), dtype=torch.bfloat16, device="cuda")
    from awr_b200 import _lib as L
    stats = L.acc_zeros(2 * Co, "cuda")
    _conv_tc(_nhwc(x), w_phys, None, y, N, H, H, Ci, Ho, Ho, Co, 4, 2, 1, 1, 1, Ci, Co * Ci, stats=stats)
    stats = L.acc_to_float(stats)
    got = _from_nhwc(y)
    assert (got - ref.detach()).abs().max().item() < 1e-2 * ref.abs().max().item()
    s1, s2 = got.double().sum(dim=(0, 2, 3)), (got.double() ** 2).sum(dim=(0, 2, 3))
    assert torch.allclose(stats[:Co].cpu().double(), s1, rtol=1e-4, atol=1e-2) and torch.allclose(stats[Co:].cpu().double(), s2, rtol=1e-4, atol=1e-2)
    dx = torch.full((N, H, H, Ci), float("nan"), dtype=torch.bfloat16, device="cuda")
    _conv_tc(_nhwc(gy), w_phys, None, dx, N, Ho, Ho, Co, H, H, Ci, 4, 2, 1, 0, Ci, 1, Co * Ci)
    assert (_from_nhwc(dx) - xr.grad).abs().max().item() < 1e-2 * xr.grad.abs().max().item()
    wr = w.clone().requires_grad_(True)
    ref_dw = torch.autograd.grad(F.conv_transpose2d(x, wr, None, stride=2, padding=1), wr, gy)[0]      # (Ci,Co,4,4)
    dW = torch.zeros(4, 4, Co, Ci, dtype=torch.float32, device="cuda")
    _wgrad_tc(_nhwc(x), _nhwc(gy), dW, N, H, H, Ci, Ho, Ho, Co, 4, 2, 1, 1, Ci, Co * Ci)
    got_dw = dW.cpu().permute(3, 2, 0, 1)
    assert (got_dw - ref_dw).abs().max().item() < 2e-3 * ref_dw.abs().max().item() + 1e-5


def test_head_conv_nchw_fp32_output():
    """1x1 conv with 64 (56 valid) output channels written as the fp32 NCHW prediction volume."""
    N, Ci, H, nv = 2, 256, 64, 56
    g = torch.Generator().manual_seed(7)
    x = _rb(torch.randn(N, Ci, H, H, generator=g))
    w = torch.zeros(64, Ci, 1, 1)
    w[:nv] = _rb(torch.randn(nv, Ci, 1, 1, generator=g) / Ci ** 0.5)
    b = torch.zeros(64); b[:nv] = torch.randn(nv, generator=g)
    ref = F.conv2d(x, w[:nv], b[:nv])
    out = torch.full((N, nv, H, H), float("nan"), dtype=torch.float32, device="cuda")
    w_phys = w.permute(2, 3, 0, 1).contiguous().to(torch.bfloat16).cuda()
    _conv_tc(_nhwc(x), w_phys, b.cuda(), out, N, H, H, Ci, H, H, 64, 1, 1, 0, 0, 1, Ci, 64 * Ci, out_mode=1, n_valid=nv)
    assert (out.cpu() - ref).abs().max().item() < 2e-5 * ref.abs().max().item() + 1e-5      # fp32 out: only accumulation-order error
