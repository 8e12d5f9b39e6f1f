"""Records what the UNMODIFIED reference util/eval_tool.py::EvalUtil produces on seeded synthetic inputs -> tests/golden/eval_cases.pt.
Run in the build container (needs /root/reference; matplotlib is absent there, so an empty stub module is pre-seeded: only plot_pck uses it)."""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, "/root/reference")
for name in ("matplotlib", "matplotlib.pyplot"):
    sys.modules.setdefault(name, types.ModuleType(name))
if not hasattr(np, "trapz"):
    np.trapz = np.trapezoid
from util.eval_tool import EvalUtil          # noqa: E402  (the reference class)
from oracle import awr_oracle as O           # noqa: E402  (input generator only)


def case(N, J, seed, use_vis):
    uvd, gt, center, M, cube, vis = O.eval_case_inputs(N, J, seed)
    ev = EvalUtil(128, O.NYU_PARAS, O.NYU_FLIP, J)
    dist = np.zeros((N, J), np.float32)
    for n in range(N):
        ev.feed(uvd[n].copy(), gt[n].copy(), center[n], M[n], cube[n], jt_vis=(vis[n] if use_vis else 0))
    # euclidean distances are only kept per joint inside the class: recover them sample by sample
    for j in range(J):
        rows = [n for n in range(N) if (not use_vis) or vis[n, j]]
        for r, n in enumerate(rows):
            dist[n, j] = ev.data[j][r]
    mean, median, auc, curve, thr = ev.get_measures()
    return dict(N=N, J=J, seed=seed, use_vis=use_vis, jt_uvd_img=np.stack(ev.jt_uvd_pred).astype(np.float32), diff=np.stack(ev.diff).astype(np.float32),
                dist=dist, mean=float(mean), median=float(median), auc=float(auc), curve=np.asarray(curve, np.float64), thresholds=np.asarray(thr))


if __name__ == "__main__":
    # jt_vis arrays cannot be recorded: the reference's own `if jt_vis == 0` (eval_tool.py:54) raises on an array, so only the default
    # "all joints visible" path exists there; the visibility mask of awr_b200.EvalUtil is checked against the oracle alone
    cases = [case(64, 14, 7, False), case(37, 21, 8, False), case(300, 14, 9, False)]
    torch.save(cases, os.path.join(HERE, "eval_cases.pt"))
    for c in cases:
        print(c["N"], c["J"], c["use_vis"], "MPE %.4f median %.4f AUC %.4f" % (c["mean"], c["median"], c["auc"]), "max dist %.2f" % c["dist"].max())
