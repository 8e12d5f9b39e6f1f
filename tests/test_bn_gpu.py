"""GPU numerics of the BatchNorm backward kernels through the C-ABI: the single-launch kernel (bulk-TMA staging + grid barrier,
awr_bn_bwd_fused) against the two-pass kernels (awr_bn_bwd_reduce + awr_bn_bwd_apply) in every mask / residual / accumulate variant, and
both against torch autograd of F.batch_norm (fp32)."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _run(lib, L, fused, dt, dout, act, y, mi, gamma, beta, remask, dy_add, want_dres, dres_add, M, C):
    tdt = torch.float32 if dt == L.F32 else torch.bfloat16
    dsums = L.acc_zeros(2 * C, "cuda")            # order-independent accumulators (awr_acc_t)
    dy = dy_add.clone() if dy_add is not None else torch.empty(M, C, device="cuda", dtype=tdt)
    dres = (dres_add.clone() if dres_add is not None else torch.empty(M, C, device="cuda", dtype=tdt)) if want_dres else None
    dg, db = torch.zeros(C, device="cuda"), torch.zeros(C, device="cuda")
    p = lambda t: None if t is None else t.data_ptr()
    mg, mb = (gamma, beta) if remask else (None, None)
    a = None if remask else act
    if fused:
        bar = torch.zeros(1, dtype=torch.int32, device="cuda")
        rc = lib.awr_bn_bwd_fused(p(dout), p(a), p(y), p(mi), p(gamma), p(mb), p(dsums), p(bar), p(dy), p(dy) if dy_add is not None else None,
                                  p(dres), p(dres) if dres_add is not None else None, p(dg), p(db), dt, M, C, 1, L.stream())
        if rc == -2:
            pytest.skip("operands do not fit in one CTA per SM: two-pass kernels only")
        L.check(rc, "fused")
    else:
        L.check(lib.awr_bn_bwd_reduce(p(dout), p(a), p(y), p(mi), p(mg), p(mb), dt, M, C, p(dsums), L.stream()), "reduce")
        L.check(lib.awr_bn_bwd_apply(p(dout), p(a), p(y), p(mi), p(dsums), p(gamma), p(dy), p(dy) if dy_add is not None else None, p(dres),
                                     p(dres) if dres_add is not None else None, p(dg), p(db), p(mb), dt, M, C, 1, L.stream()), "apply")
    torch.cuda.synchronize()
    return dy, dres, L.acc_to_float(dsums).float(), dg, db


@pytest.mark.parametrize("M,C", [(2048, 512), (8192, 256), (32768, 128), (600, 64)])
@pytest.mark.parametrize("variant", ["remask", "act_res", "plain_acc"])
@pytest.mark.parametrize("dtname", ["f32", "bf16"])
def test_fused_matches_two_pass(M, C, variant, dtname):
    from awr_b200 import _lib as L
    lib = L.lib()
    dt = L.F32 if dtname == "f32" else L.BF16
    tdt = torch.float32 if dtname == "f32" else torch.bfloat16
    g = torch.Generator().manual_seed(M + C)
    y = torch.randn(M, C, generator=g).cuda().to(tdt)
    dout = torch.randn(M, C, generator=g).cuda().to(tdt)
    gamma, beta = (torch.rand(C, generator=g) + 0.5).cuda(), (torch.randn(C, generator=g) * 0.3).cuda()
    yf = y.float()
    mean, var = yf.mean(0), yf.var(0, unbiased=False)
    mi = torch.cat([mean, torch.rsqrt(var + 1e-5)]).contiguous()
    res = torch.randn(M, C, generator=g).cuda()
    act = F.relu((yf - mean) * mi[C:] * gamma + beta + (res if variant == "act_res" else 0)).to(tdt)
    remask = variant == "remask"
    dy_add = torch.randn(M, C, generator=g).cuda().to(tdt) if variant == "plain_acc" else None
    dres_add = torch.randn(M, C, generator=g).cuda().to(tdt) if variant == "act_res" else None
    a = None if variant == "plain_acc" else act
    outs = [_run(lib, L, f, dt, dout, a, y, mi, gamma, beta, remask, dy_add, variant == "act_res", dres_add, M, C) for f in (False, True)]
    tol = 2e-5 if dtname == "f32" else 1.6e-2
    for t2, tf, name in zip(outs[0], outs[1], ("dy", "dres", "dsums", "dgamma", "dbeta")):
        if t2 is None:
            assert tf is None
            continue
        scale = t2.float().abs().max().item() + 1e-6
        assert (t2.float() - tf.float()).abs().max().item() <= tol * scale, name


def test_two_pass_and_fused_vs_autograd_fp32():
    from awr_b200 import _lib as L
    lib = L.lib()
    N, C, H = 4, 128, 16
    M = N * H * H
    g = torch.Generator().manual_seed(9)
    x = torch.randn(N, C, H, H, generator=g).cuda().requires_grad_(True)
    gamma = (torch.rand(C, generator=g) + 0.5).cuda().requires_grad_(True)
    beta = (torch.randn(C, generator=g) * 0.3).cuda().requires_grad_(True)
    out = F.relu(F.batch_norm(x, None, None, gamma, beta, training=True, eps=1e-5))
    go = torch.randn(out.shape, generator=g).cuda()
    out.backward(go)
    nhwc = lambda t: t.detach().permute(0, 2, 3, 1).reshape(M, C).contiguous()
    y, dout = nhwc(x), nhwc(go)
    mean, var = y.mean(0), y.var(0, unbiased=False)
    mi = torch.cat([mean, torch.rsqrt(var + 1e-5)]).contiguous()
    for fused in (False, True):
        dy, _, _, dg, db = _run(lib, L, fused, L.F32, dout, None, y, mi, gamma.detach(), beta.detach(), True, None, False, None, M, C)
        ref = nhwc(x.grad)
        assert (dy - ref).abs().max().item() <= 1e-4 * ref.abs().max().item() + 1e-6
        assert torch.allclose(dg, gamma.grad, rtol=1e-4, atol=1e-4) and torch.allclose(db, beta.grad, rtol=1e-4, atol=1e-4)
