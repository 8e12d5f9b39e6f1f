"""Pin the CPU oracle (oracle/awr_oracle.py) against golden vectors recorded from the
unmodified reference modules (tests/golden/make_golden.py)."""
import os

import pytest
import torch

from oracle import awr_oracle as O

GOLD = os.path.join(os.path.dirname(__file__), "golden")
HEAD = torch.load(os.path.join(GOLD, "head_cases.pt"))
BACK = torch.load(os.path.join(GOLD, "backbone_cases.pt"))


def sub(t, step=8):
    return t[..., ::step, ::step]


def checks(t):
    t = t.double()
    return torch.tensor([t.sum(), t.abs().sum(), (t * t).sum()], dtype=torch.float64)


def assert_chk(t, chk, rtol=1e-5):
    got = checks(t)
    scale = chk[1].abs().clamp_min(1e-12)
    assert ((got - chk).abs() / torch.stack((scale, scale, chk[2].abs().clamp_min(1e-12)))).max() < rtol, (got, chk)


@pytest.mark.parametrize("c", HEAD, ids=lambda c: f"B{c['B']}J{c['J']}F{c['F']}ks{c['ks']}")
def test_head_and_loss_vs_reference(c):
    img, jt, pred, g_uvd = O.head_case_inputs(c["B"], c["J"], c["F"], c["H"], c["ks"], c["seed"])
    assert_chk(pred, c["pred_chk"])
    gt = O.joint2offset(jt, img, c["ks"], c["F"])
    assert torch.allclose(sub(gt), c["gt_sub"], atol=1e-6)
    assert_chk(gt, c["gt_chk"])
    uvd = O.offset2joint_softmax(pred, img, c["ks"])
    assert torch.allclose(uvd, c["uvd"], atol=2e-6)
    assert torch.allclose(O.smooth_l1(uvd, jt), c["l_coord"], rtol=1e-5, atol=1e-9)
    assert torch.allclose(O.smooth_l1(pred, gt), c["l_dense"], rtol=1e-5, atol=1e-9)
    # analytic backward of head + both losses == reference autograd
    g_from_coord = c["cw"] * O.smooth_l1_grad(uvd, jt)
    dpred = O.offset2joint_softmax_bwd(pred, img, c["ks"], g_from_coord) + c["dw"] * O.smooth_l1_grad(pred, gt)
    assert torch.allclose(sub(dpred), c["dpred_sub"], atol=1e-9, rtol=1e-4)
    assert_chk(dpred, c["dpred_chk"], rtol=1e-4)
    dhead = O.offset2joint_softmax_bwd(pred, img, c["ks"], g_uvd)
    assert torch.allclose(sub(dhead), c["dpred_head_sub"], atol=1e-6, rtol=1e-4)
    assert_chk(dhead, c["dpred_head_chk"], rtol=1e-4)


def _weights(c):
    kind, n = c["net"].split("_")
    if kind == "resnet":
        return O.randomize_bn(O.resnet_deconv_init(int(n), c["J"], c["ds"], c["seed"], head_std=c["head_std"]), c["seed"] + 1)
    return O.randomize_bn(O.hourglass_init(int(n), c["J"], c["seed"], head_gain=c["head_std"]), c["seed"] + 1)


@pytest.mark.parametrize("c", BACK, ids=lambda c: f"{c['net']}_ds{c['ds']}_B{c['B']}")
def test_backbone_eval_vs_reference(c):
    sd = _weights(c)
    img, jt = O.synthetic_batch(c["B"], c["H"], c["J"], c["seed"] + 2)
    kind, n = c["net"].split("_")
    with torch.no_grad():
        outs = [O.resnet_deconv_forward(sd, img, int(n), c["ds"])] if kind == "resnet" else O.hourglass_forward(sd, img, int(n))
    assert len(outs) == len(c["eval_out_sub"])
    for o, s, k, u in zip(outs, c["eval_out_sub"], c["eval_out_chk"], c["eval_uvd"]):
        assert torch.allclose(sub(o), s, atol=1e-5, rtol=1e-4)
        assert_chk(o, k, rtol=1e-4)
        assert torch.allclose(O.offset2joint_softmax(o, img, c["ks"]), u, atol=1e-4)


@pytest.mark.parametrize("c", [c for c in BACK if "l_dense" in c], ids=lambda c: f"{c['net']}_B{c['B']}")
def test_train_step_vs_reference(c):
    sd = _weights(c)
    img, jt = O.synthetic_batch(c["B"], c["H"], c["J"], c["seed"] + 2)
    loss, lc, ld, uvd, pred, grads, new_stats = O.loss_and_grads(sd, img, jt, c["net"], c["ds"], c["ks"], 1.0, 1.0)
    assert torch.allclose(sub(pred), c["train_out_sub"], atol=1e-5, rtol=1e-4)
    assert torch.allclose(uvd, c["train_uvd"], atol=1e-4)
    assert torch.allclose(lc, c["l_coord"], rtol=1e-4) and torch.allclose(ld, c["l_dense"], rtol=1e-4)
    for k, chk in c["grad_chk"].items():
        if chk is None:
            assert grads[k] is None          # hourglass skip_layer params never get a gradient
        else:
            assert_chk(grads[k], chk, rtol=2e-3)
    for k, g in c["grad_small"].items():
        assert torch.allclose(grads[k], g, rtol=2e-3, atol=1e-7 + 1e-3 * g.abs().max().item()), k
    for k, v in c["running"].items():
        assert torch.allclose(new_stats[k], v, rtol=1e-4, atol=1e-6), k


def test_adam_matches_torch():
    torch.manual_seed(0)
    p = torch.randn(1000); g = torch.randn(1000)
    p_ref = torch.nn.Parameter(p.clone()); opt = torch.optim.Adam([p_ref], lr=1e-3)
    m = torch.zeros(1000); v = torch.zeros(1000); p2 = p.clone()
    for step in range(1, 4):
        p_ref.grad = g * step; opt.step()
        O.adam_step(p2, g * step, m, v, step)
    assert torch.allclose(p2, p_ref.detach(), atol=1e-7)
