"""Generate tests/golden/*.pt by running the UNMODIFIED reference modules (CPU, fp32).

Run in the build container only (needs /root/reference):
    python tests/golden/make_golden.py

The reference is imported from /root/reference (never copied).  Inputs and weights
come from seeded generators in oracle/awr_oracle.py, so the fixtures only need to
hold the reference's OUTPUTS (plus small inputs); tests rebuild the same inputs
from the seeds on any box.
"""
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")

from oracle import awr_oracle as O                                   # noqa: E402
from model.resnet_deconv import get_deconv_net                       # noqa: E402  (reference)
from model.hourglass import PoseNet                                  # noqa: E402  (reference)
from model.loss import My_SmoothL1Loss                               # noqa: E402  (reference)
from util.feature_tool import FeatureModule                          # noqa: E402  (reference)

torch.set_num_threads(8)
FM = FeatureModule()
CRIT = My_SmoothL1Loss()


def sub(t, step=8):
    """Strided subsample kept in the fixture for dense tensors (full tensors are checked by checksum)."""
    return t[..., ::step, ::step].contiguous().clone()


def checks(t):
    t = t.double()
    return torch.tensor([t.sum(), t.abs().sum(), (t * t).sum()], dtype=torch.float64)


def head_cases():
    out = []
    for (B, J, Fs, H, ks, seed) in O.HEAD_CASES:
        img, jt, pred0, g_uvd = O.head_case_inputs(B, J, Fs, H, ks, seed)
        gt = FM.joint2offset(jt, img, ks, Fs)
        pred = pred0.clone().requires_grad_(True)
        uvd = FM.offset2joint_softmax(pred, img, ks)
        l_coord = CRIT(uvd, jt)
        l_dense = CRIT(pred, gt)
        (0.7 * l_coord + 1.3 * l_dense).backward()
        pred2 = pred0.clone().requires_grad_(True)
        (FM.offset2joint_softmax(pred2, img, ks) * g_uvd).sum().backward()
        out.append(dict(B=B, J=J, F=Fs, H=H, ks=ks, seed=seed,
                        gt_sub=sub(gt), gt_chk=checks(gt), pred_chk=checks(pred0), uvd=uvd.detach().clone(),
                        l_coord=l_coord.detach().clone(), l_dense=l_dense.detach().clone(),
                        cw=0.7, dw=1.3, dpred_sub=sub(pred.grad), dpred_chk=checks(pred.grad),
                        dpred_head_sub=sub(pred2.grad), dpred_head_chk=checks(pred2.grad)))
    return out


def backbone_case(net, J, ds, B, H, seed, head_std, ks, train_step):
    kind, n = net.split("_")
    if kind == "resnet":
        sd = O.randomize_bn(O.resnet_deconv_init(int(n), J, ds, seed, head_std=head_std), seed + 1)
        ref = get_deconv_net(int(n), J, ds)
    else:
        sd = O.randomize_bn(O.hourglass_init(int(n), J, seed, head_gain=head_std), seed + 1)
        ref = PoseNet(net, J)
    missing = ref.load_state_dict(sd, strict=True)
    img, jt = O.synthetic_batch(B, H, J, seed + 2)
    case = dict(net=net, J=J, ds=ds, B=B, H=H, seed=seed, head_std=head_std, ks=ks)
    # eval-mode forward (running statistics)
    ref.eval()
    with torch.no_grad():
        o = ref(img)
        outs = o if isinstance(o, list) else [o]
        case["eval_out_sub"] = [sub(t) for t in outs]
        case["eval_out_chk"] = [checks(t) for t in outs]
        case["eval_uvd"] = [FM.offset2joint_softmax(t, img, ks) for t in outs]
    if train_step:
        ref.train()
        Fs = H // ds
        gt = FM.joint2offset(jt, img, ks, Fs)
        o = ref(img)
        pred = o[-1] if isinstance(o, list) else o
        uvd = FM.offset2joint_softmax(pred, img, ks)
        l_coord, l_dense = CRIT(uvd, jt), CRIT(pred, gt)
        loss = 1.0 * l_coord + 1.0 * l_dense
        ref.zero_grad()
        loss.backward()
        case.update(train_out_sub=sub(pred.detach()), train_out_chk=checks(pred.detach()), train_uvd=uvd.detach().clone(),
                    l_coord=l_coord.detach().clone(), l_dense=l_dense.detach().clone())
        grads = {k: p.grad for k, p in ref.named_parameters()}
        case["grad_chk"] = {k: (checks(g) if g is not None else None) for k, g in grads.items()}
        small = {}
        for k, g_ in grads.items():        # keep a few whole small gradients for element-wise comparison
            if g_ is not None and g_.numel() <= 4096:
                small[k] = g_.clone()
        case["grad_small"] = small
        rs = {k: v.clone() for k, v in ref.state_dict().items() if k.endswith(("running_mean", "running_var"))}
        # keep only the first few BN layers' running stats
        case["running"] = {k: rs[k] for k in list(rs)[:8]}
    return case


def schemas():
    out = {}
    for name, ctor in [("resnet_18_ds2", lambda: get_deconv_net(18, 14, 2)), ("resnet_50_ds2", lambda: get_deconv_net(50, 14, 2)),
                       ("resnet_18_ds4", lambda: get_deconv_net(18, 14, 4)), ("resnet_18_ds1_J21", lambda: get_deconv_net(18, 21, 1)),
                       ("hourglass_1", lambda: PoseNet("hourglass_1", 14)), ("hourglass_2", lambda: PoseNet("hourglass_2", 14))]:
        m = ctor()
        out[name] = dict(state=[(k, tuple(v.shape), str(v.dtype)) for k, v in m.state_dict().items()],
                         params=[k for k, _ in m.named_parameters()])
    return out


def main():
    torch.manual_seed(0)
    torch.save(schemas(), os.path.join(HERE, "schemas.pt"))
    torch.save(head_cases(), os.path.join(HERE, "head_cases.pt"))
    cases = [
        # C1: ResNet18 B=1 fp32 forward; peaked head so the softmax is not vacuous
        backbone_case("resnet_18", 14, 2, 1, 128, 100, 0.02, 1.0, False),
        # small full train step (forward, losses, every parameter gradient checksum)
        backbone_case("resnet_18", 14, 2, 2, 128, 110, 0.02, 1.0, True),
        backbone_case("resnet_18", 14, 4, 2, 128, 120, 0.02, 1.0, False),
        backbone_case("resnet_50", 14, 2, 1, 128, 130, 0.004, 1.0, False),
        backbone_case("hourglass_1", 14, 2, 2, 128, 140, 1.0, 0.4, True),
        backbone_case("hourglass_2", 14, 2, 1, 128, 150, 1.0, 0.4, False),
    ]
    torch.save(cases, os.path.join(HERE, "backbone_cases.pt"))
    for c in cases:
        print(c["net"], c["B"], [float(x.abs().max()) for x in c["eval_uvd"]],
              "eval heat range", float(c["eval_out_sub"][-1][:, 3 * c["J"]:].min()), float(c["eval_out_sub"][-1][:, 3 * c["J"]:].max()))
    sz = sum(os.path.getsize(os.path.join(HERE, f)) for f in os.listdir(HERE) if f.endswith(".pt"))
    print("golden bytes", sz)


if __name__ == "__main__":
    main()
