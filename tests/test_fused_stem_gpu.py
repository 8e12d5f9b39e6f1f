"""GPU numerics of the fused stem kernels (through the C-ABI, fp32 storage) vs plain PyTorch fp32:
stem conv + fused BN statistics, BN+ReLU+MaxPool forward, MaxPool+BN backward (both the 3/2/1 specialisation and the generic path)."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _nhwc(t):
    return t.permute(0, 2, 3, 1).contiguous()


def test_stem_conv_with_fused_statistics():
    from awr_b200 import _lib as L
    N, H, Co, k = 3, 32, 64, 5
    g = torch.Generator().manual_seed(1)
    x = torch.randn(N, 1, H, H, generator=g).cuda()
    w = (torch.randn(Co, 1, k, k, generator=g) * 0.2).cuda()
    b = torch.randn(Co, generator=g).cuda()
    ref = F.conv2d(x, w, b, padding=2)
    y = torch.empty(N, H, H, Co, device="cuda")
    stats = L.acc_zeros(2 * Co, "cuda")
    w_phys = w.permute(2, 3, 0, 1).contiguous().view(k * k, Co)
    L.check(L.lib().awr_stem_conv(x.data_ptr(), w_phys.data_ptr(), b.data_ptr(), y.data_ptr(), stats.data_ptr(), L.F32, N, H, H, Co, k, L.stream()), "stem")
    torch.cuda.synchronize()
    assert torch.allclose(y.permute(0, 3, 1, 2), ref, atol=1e-5, rtol=1e-5)                    # fp32, tolerance 1e-5
    stats = L.acc_to_float(stats).float()
    assert torch.allclose(stats[:Co], ref.sum(dim=(0, 2, 3)), rtol=1e-4, atol=1e-3)
    assert torch.allclose(stats[Co:], (ref * ref).sum(dim=(0, 2, 3)), rtol=1e-4, atol=1e-3)


@pytest.mark.parametrize("k,s,p", [(3, 2, 1), (2, 2, 0)])
def test_bn_relu_maxpool_fwd_and_bwd(k, s, p):
    from awr_b200 import _lib as L
    N, C, H = 2, 64, 16
    g = torch.Generator().manual_seed(2)
    y = torch.randn(N, C, H, H, generator=g).cuda()
    gamma, beta = (torch.rand(C, generator=g) + 0.5).cuda(), (torch.randn(C, generator=g) * 0.3).cuda()
    yr = y.clone().requires_grad_(True)
    gr, br = gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
    rm, rv = torch.zeros(C, device="cuda"), torch.ones(C, device="cuda")
    ref = F.max_pool2d(F.relu(F.batch_norm(yr, rm.clone(), rv.clone(), gr, br, training=True, momentum=0.1, eps=1e-5)), k, s, p)
    dout = torch.randn(ref.shape, generator=g).cuda()
    ref.backward(dout)
    lib = L.lib()
    yn = _nhwc(y)
    sums = torch.cat([yn.sum(dim=(0, 1, 2)), (yn * yn).sum(dim=(0, 1, 2))]).contiguous()
    Ho = ref.shape[-1]
    out = torch.empty(N, Ho, Ho, C, device="cuda")
    idx = torch.empty(N, Ho, Ho, C, dtype=torch.uint8, device="cuda")
    mi = torch.empty(2 * C, device="cuda")
    nbt = torch.zeros((), dtype=torch.long, device="cuda")
    L.check(lib.awr_bn_relu_maxpool_fwd(yn.data_ptr(), L.acc_from_float(sums).data_ptr(), gamma.data_ptr(), beta.data_ptr(), rm.data_ptr(), rv.data_ptr(), nbt.data_ptr(),
                                        mi.data_ptr(), out.data_ptr(), idx.data_ptr(), L.F32, N, H, H, C, k, s, p, 0.1, 1e-5, 1, L.stream()), "fwd")
    torch.cuda.synchronize()
    assert torch.allclose(out.permute(0, 3, 1, 2), ref.detach(), atol=1e-5, rtol=1e-5)
    assert nbt.item() == 1 and torch.allclose(rm, 0.1 * y.mean(dim=(0, 2, 3)), atol=1e-6)
    dsums = L.acc_zeros(2 * C, "cuda")
    dy = torch.empty_like(yn)
    dg, db = torch.zeros(C, device="cuda"), torch.zeros(C, device="cuda")
    dpool = _nhwc(dout)
    for ps in (0, 1):
        L.check(lib.awr_maxpool_bn_bwd(dpool.data_ptr(), idx.data_ptr(), yn.data_ptr(), mi.data_ptr(), gamma.data_ptr(), beta.data_ptr(), dsums.data_ptr(),
                                       dy.data_ptr(), dg.data_ptr(), db.data_ptr(), L.F32, N, H, H, C, k, s, p, ps, 0, L.stream()), "bwd")
    torch.cuda.synchronize()
    scale = yr.grad.abs().max().item()
    assert (dy.permute(0, 3, 1, 2) - yr.grad).abs().max().item() < 1e-4 * scale + 1e-6
    assert torch.allclose(dg, gr.grad, rtol=1e-4, atol=1e-4) and torch.allclose(db, br.grad, rtol=1e-4, atol=1e-4)


def test_bn_relu_maxpool3_fwd_bf16_value_and_argmax():
    """bf16 storage path of the 3x3/2/1 kernel (packed value|tap keys): pooled values match torch on the bf16-rounded activation,
    every arg-max byte points at a tap holding the pooled value, and ties (all-zero windows) resolve to the first valid tap as ATen does."""
    from awr_b200 import _lib as L
    N, C, H, k, s, p = 2, 64, 32, 3, 2, 1
    g = torch.Generator().manual_seed(5)
    y = (torch.randn(N, C, H, H, generator=g) - 0.4).cuda().bfloat16()
    gamma, beta = (torch.rand(C, generator=g) + 0.5).cuda(), (torch.randn(C, generator=g) * 0.3).cuda()
    yf = y.float()
    yn = _nhwc(y)
    ynf = yn.float()
    sums = torch.cat([ynf.sum(dim=(0, 1, 2)), (ynf * ynf).sum(dim=(0, 1, 2))]).contiguous()
    rm, rv = torch.zeros(C, device="cuda"), torch.ones(C, device="cuda")
    nbt = torch.zeros((), dtype=torch.long, device="cuda")
    mi = torch.empty(2 * C, device="cuda")
    Ho = (H + 2 * p - k) // s + 1
    out = torch.empty(N, Ho, Ho, C, device="cuda", dtype=torch.bfloat16)
    idx = torch.empty(N, Ho, Ho, C, dtype=torch.uint8, device="cuda")
    L.check(L.lib().awr_bn_relu_maxpool_fwd(yn.data_ptr(), L.acc_from_float(sums).data_ptr(), gamma.data_ptr(), beta.data_ptr(), rm.data_ptr(), rv.data_ptr(),
                                            nbt.data_ptr(), mi.data_ptr(), out.data_ptr(), idx.data_ptr(), L.BF16, N, H, H, C, k, s, p, 0.1, 1e-5, 1,
                                            L.stream()), "fwd")
    torch.cuda.synchronize()
    cnt = N * H * H
    mean = sums[:C] / cnt
    var = (sums[C:] / cnt - mean * mean).clamp_min(0)
    sc = gamma * torch.rsqrt(var + 1e-5)
    sh = beta - mean * sc
    act = F.relu(yf * sc.view(1, C, 1, 1) + sh.view(1, C, 1, 1)).bfloat16().float()          # what the unfused path would have stored
    ref = F.max_pool2d(act, k, s, p)
    o = out.float().permute(0, 3, 1, 2)
    assert (o - ref).abs().max().item() <= 1e-2 * ref.abs().max().item()                      # one bf16 ulp of slack for sc/sh rounding
    # arg-max consistency: the tap each byte names holds the pooled value
    tap = idx.long().permute(0, 3, 1, 2)
    ho = torch.arange(Ho, device="cuda").view(1, 1, Ho, 1)
    wo = torch.arange(Ho, device="cuda").view(1, 1, 1, Ho)
    hi, wi = ho * s - p + tap // k, wo * s - p + tap % k
    assert (hi >= 0).all() and (hi < H).all() and (wi >= 0).all() and (wi < H).all()
    picked = act[torch.arange(N, device="cuda").view(N, 1, 1, 1), torch.arange(C, device="cuda").view(1, C, 1, 1), hi, wi]
    assert (picked - o).abs().max().item() <= 1e-2 * ref.abs().max().item()
    # ties: an all-zero window resolves to its first valid tap (row-major scan)
    zero = (o == 0)
    first_r = (p - ho * s).clamp_min(0)
    first_c = (p - wo * s).clamp_min(0)
    first = (first_r * k + first_c).expand_as(tap)
    assert zero.any() and (tap[zero] == first[zero]).all()


@pytest.mark.parametrize("N,H,W,with_bias", [(3, 8, 128, True), (2, 16, 256, False), (5, 128, 128, False)])
def test_stem_conv_tensor_core_bf16(N, H, W, with_bias):
    """csrc/stem_tc.cu: the 1-channel 5x5 stem on tcgen05 with a hand-built im2col operand (hi/lo bf16 split of the fp32 depth and of the
    weights).  The bf16 output equals torch's fp32 convolution to output rounding (one bf16 ulp = 2^-8 relative), far tighter than a plain
    bf16 x bf16 product would be; the fused statistics are the sums of the stored values."""
    from awr_b200 import _lib as L
    Co, k = 64, 5
    g = torch.Generator().manual_seed(N * 1000 + W)
    x = torch.randn(N, 1, H, W, generator=g).cuda()
    x[:, :, :, :3] = 1.0                                                      # background columns at the left border
    w = (torch.randn(Co, 1, k, k, generator=g) * 0.2).cuda()
    b = torch.randn(Co, generator=g).cuda() if with_bias else None
    ref = F.conv2d(x, w, b, padding=2)
    y = torch.full((N, H, W, Co), float("nan"), device="cuda", dtype=torch.bfloat16)
    stats = L.acc_zeros(2 * Co, "cuda")
    w_phys = w.permute(2, 3, 0, 1).contiguous().view(k * k, Co)
    L.check(L.lib().awr_stem_conv(x.data_ptr(), w_phys.data_ptr(), None if b is None else b.data_ptr(), y.data_ptr(), stats.data_ptr(), L.BF16,
                                  N, H, W, Co, k, L.stream()), "stem")
    torch.cuda.synchronize()
    got = y.float().permute(0, 3, 1, 2)
    assert not torch.isnan(got).any()
    err = (got - ref).abs()
    assert (err <= ref.abs() * 2.0 ** -8 + 2e-5).all(), (err.max().item(), (err / (ref.abs() + 1e-3)).max().item())
    st = L.acc_to_float(stats).double()
    assert torch.allclose(st[:Co], got.double().sum(dim=(0, 2, 3)), rtol=1e-4, atol=1e-2)
    assert torch.allclose(st[Co:], (got.double() ** 2).sum(dim=(0, 2, 3)), rtol=1e-4, atol=1e-2)


@pytest.mark.parametrize("N,H,W,with_bias", [(3, 8, 128, True), (2, 16, 256, False), (5, 128, 128, True)])
def test_stem_wgrad_tensor_core_bf16(N, H, W, with_bias):
    """csrc/stem_tc.cu: stem weight gradient as a pixel-contraction GEMM on tcgen05 (MN-major hand-built im2col operand x TMA-loaded dy; the
    bias gradient rides along as a constant-one tap).  dy is bf16, the depth keeps fp32 accuracy through the hi/lo split, accumulation is
    fp32: the result matches torch's fp32 weight gradient on the same bf16 dy to accumulation-order error."""
    from awr_b200 import _lib as L
    Co, k = 64, 5
    g = torch.Generator().manual_seed(N * 100 + W)
    x = torch.randn(N, 1, H, W, generator=g).cuda()
    dy = torch.randn(N, Co, H, W, generator=g).bfloat16().cuda()
    w = torch.zeros(Co, 1, k, k, device="cuda", dtype=torch.float64, requires_grad=True)      # float64 reference (no TF32, no ordering noise)
    b = torch.zeros(Co, device="cuda", dtype=torch.float64, requires_grad=True)
    F.conv2d(x.double(), w, b, padding=2).backward(dy.double())
    dyn = dy.permute(0, 2, 3, 1).contiguous()
    det = L.deterministic()
    dW = L.acc_zeros(k * k * Co, "cuda") if det else torch.zeros(k * k, Co, device="cuda")
    db = (L.acc_zeros(Co, "cuda") if det else torch.zeros(Co, device="cuda")) if with_bias else None
    L.check(L.lib().awr_stem_wgrad(x.data_ptr(), dyn.data_ptr(), dW.data_ptr(), None if db is None else db.data_ptr(), L.BF16, N, H, W, Co, k,
                                   L.stream()), "stem_wgrad")
    torch.cuda.synchronize()
    got = (L.acc_to_float(dW) if det else dW).float().view(k, k, Co).permute(2, 0, 1)
    ref = w.grad[:, 0]
    err = (got.double() - ref).abs().max().item()
    print(f"stem wgrad N={N} {H}x{W}: max |dW - float64 reference| = {err:.3e} of {ref.abs().max().item():.1f}")
    assert err < 2e-5 * ref.abs().max().item() + 1e-3, (err, ref.abs().max().item())
    if with_bias:
        gb = (L.acc_to_float(db) if det else db).double()
        assert (gb - b.grad).abs().max().item() < 2e-5 * b.grad.abs().max().item() + 1e-3
