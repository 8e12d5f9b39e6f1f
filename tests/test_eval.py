"""EvalUtil (util/eval_tool.py): the numpy oracle against vectors recorded from the unmodified reference (CPU), and the device kernels +
drop-in class against both (GPU)."""
import os

import numpy as np
import pytest
import torch

from oracle import awr_oracle as O

GOLD = torch.load(os.path.join(os.path.dirname(__file__), "golden", "eval_cases.pt"), weights_only=False)


def _inputs(c):
    return O.eval_case_inputs(c["N"], c["J"], c["seed"])


@pytest.mark.parametrize("c", GOLD, ids=lambda c: f"N{c['N']}J{c['J']}")
def test_oracle_matches_reference_vectors(c):
    uvd, gt, center, M, cube, _ = _inputs(c)
    outs = [O.eval_feed_np(uvd[n], gt[n], center[n], M[n], cube[n], 128) for n in range(c["N"])]
    assert np.array_equal(np.stack([o[0] for o in outs]), c["jt_uvd_img"])          # same numpy ops, same dtypes: bit-exact
    dist = np.stack([o[1] for o in outs])
    assert np.array_equal(dist, c["dist"]) and np.array_equal(np.stack([o[2] for o in outs]), c["diff"])
    mean, median, auc, curve, thr = O.eval_measures_np(dist)
    assert float(mean) == c["mean"] and float(median) == c["median"] and float(auc) == c["auc"]
    assert np.array_equal(curve, c["curve"]) and np.array_equal(thr, c["thresholds"])


@pytest.mark.gpu
@pytest.mark.parametrize("c", GOLD, ids=lambda c: f"N{c['N']}J{c['J']}")
def test_device_eval_matches_reference_vectors(c):
    import awr_b200
    uvd, gt, center, M, cube, _ = _inputs(c)
    ev = awr_b200.EvalUtil(128, O.NYU_PARAS, O.NYU_FLIP, c["J"])
    t = lambda a: torch.from_numpy(a).cuda()
    half = c["N"] // 2                                                                  # two device batches + per-sample numpy feeds
    ev.feed_batch(t(uvd[:half]), t(gt[:half]), t(center[:half]), t(M[:half]), t(cube[:half]))
    for n in range(half, c["N"]):
        keep = uvd[n].copy()
        ev.feed(uvd[n], gt[n], center[n], M[n], cube[n])
        assert np.array_equal(keep, uvd[n])                                             # unlike the reference, the argument is left alone
    dist = ev.errors().cpu().numpy()
    # tolerance: float32 intrinsics / adjugate inverse instead of LAPACK -> a few 1e-4 px, i.e. < 2e-3 mm on ~1 m depths
    assert np.abs(dist - c["dist"]).max() < 2e-3
    got = np.stack(ev.jt_uvd_pred)
    assert np.abs(got - c["jt_uvd_img"]).max() < 2e-3 and np.abs(np.stack(ev.diff) - c["diff"]).max() < 2e-3
    mean, median, auc, curve, thr = ev.get_measures()
    assert abs(float(mean) - c["mean"]) < 1e-3 and abs(float(median) - c["median"]) < 2e-3
    # a PCK count can only differ where an error sits within 2e-3 mm of a threshold
    assert abs(float(auc) - c["auc"]) < 2e-3 and np.abs(curve - c["curve"]).max() <= 2.0 / c["N"] and np.array_equal(thr, c["thresholds"])


@pytest.mark.gpu
def test_device_eval_visibility_mask_vs_oracle():
    import awr_b200
    N, J = 50, 14
    uvd, gt, center, M, cube, vis = O.eval_case_inputs(N, J, 21)
    vis[:, 3] = False                                                                    # a joint that is never visible is skipped entirely
    ev = awr_b200.EvalUtil(128, O.NYU_PARAS, O.NYU_FLIP, J)
    t = lambda a: torch.from_numpy(a).cuda()
    ev.feed_batch(t(uvd), t(gt), t(center), t(M), t(cube), t(vis))
    dist = np.stack([O.eval_feed_np(uvd[n], gt[n], center[n], M[n], cube[n], 128)[1] for n in range(N)])
    ref = O.eval_measures_np(dist, vis)
    got = ev.get_measures()
    d = ev.errors().cpu().numpy()
    assert (d[~vis] == -1).all() and np.abs(d[vis] - dist[vis]).max() < 2e-3
    assert abs(float(got[0]) - float(ref[0])) < 1e-3 and abs(float(got[1]) - float(ref[1])) < 2e-3 and abs(float(got[2]) - float(ref[2])) < 2e-3
    assert np.abs(got[3] - ref[3]).max() <= 2.0 / 40
