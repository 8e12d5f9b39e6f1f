import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import awr_oracle as O
import awr_b200
from test_backbone_gpu import BACK, _build
which = sys.argv[1] if len(sys.argv) > 1 else "hourglass_1"
prec = sys.argv[2] if len(sys.argv) > 2 else "fp32"
c = [c for c in BACK if c["net"] == which and "l_dense" in c][0]
m, sd = _build(c, prec)
m.train()
img, jt = O.synthetic_batch(c["B"], c["H"], c["J"], c["seed"] + 2)
# oracle on CPU (full grads)
loss, lc, ld, uvd_o, pred_o, grads_o, ns = O.loss_and_grads(sd, img, jt, c["net"], c["ds"], c["ks"], 1.0, 1.0)
img, jt = img.cuda(), jt.cuda()
FM = awr_b200.FeatureModule(); crit = awr_b200.My_SmoothL1Loss().cuda()
gt = FM.joint2offset(jt, img, c["ks"], c["H"] // c["ds"])
o = m(img); pred = o[-1] if isinstance(o, list) else o
uvd = FM.offset2joint_softmax(pred, img, c["ks"])
l = crit(uvd, jt) + crit(pred, gt)
m.zero_grad(); l.backward()
print("pred err", (pred.detach().cpu() - pred_o).abs().max().item(), "scale", pred_o.abs().max().item())
print("loss", l.item(), loss.item())
rows = []
for k, p in m.named_parameters():
    go = grads_o[k]
    if go is None:
        continue
    g = p.grad.cpu()
    e = (g - go).abs().max().item(); s = go.abs().max().item()
    rows.append((e / (s + 1e-12), k, e, s))
rows = [r for r in rows if r[3] > 1e-8]; rows.sort(reverse=True)
for r in rows[:25]:
    print("%.3e  %-40s err %.3e scale %.3e" % r)
