"""Drop-in for the reference's util/eval_tool.py::EvalUtil (eval_tool.py:5-135) with the arithmetic on the device.

The reference feeds ONE sample at a time from numpy copies (train.py:141-148: five `.cpu().numpy()` per frame, every step), keeps
Python lists per joint and reduces them in numpy.  Here `feed_batch` takes the step's device tensors as they are (no host sync),
`awr_eval_feed` (csrc/eval.cu) turns them into per-joint errors in one launch, and `get_measures` reduces everything collected so
far with `awr_eval_measures`; only J sums and J x 100 counters cross to the host.  `feed` keeps the reference's per-sample
numpy signature (it stages the sample and flushes in batches), `jt_uvd_pred` / `diff` expose the same lists the callers save
(train.py:219-221, test.py:103-108).  Unlike the reference, `feed` does not modify its `jt_uvd_pred` argument in place.
"""
import numpy as np
import torch

_trapz = getattr(np, "trapezoid", None) or np.trapz      # numpy >= 2 renames trapz

from . import _lib as L


class EvalUtil:
    def __init__(self, img_size, paras, flip, num_kp, device=None):
        self.img_size, self.paras, self.flip, self.num_kp = img_size, tuple(float(p) for p in paras), float(flip), int(num_kp)
        self.device = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        self._dist, self._uvd, self._diff = [], [], []          # device chunks
        self._staged = []                                       # per-sample host tuples waiting for a flush
        self.lib = L.lib()

    # ---- feeding -------------------------------------------------------------------------------------------------------
    def feed_batch(self, jt_uvd_pred, jt_xyz_gt, center_xyz, M, cube, jt_vis=None):
        """Device tensors of one step: (B,J,3), (B,J,3), (B,3), (B,3,3), (B,3) [, (B,J) visibility].  No host synchronisation."""
        dev = self.device
        prep = lambda t: torch.as_tensor(t).detach().to(dev, torch.float32).contiguous()
        uvd, gt, ctr, Mt, cb = prep(jt_uvd_pred), prep(jt_xyz_gt), prep(center_xyz), prep(M), prep(cube)
        B, J, _ = uvd.shape
        if J != self.num_kp or gt.shape != uvd.shape or ctr.shape != (B, 3) or Mt.shape != (B, 3, 3) or cb.shape != (B, 3):
            raise ValueError("feed_batch: expected (B,J,3), (B,J,3), (B,3), (B,3,3), (B,3)")
        vis = None
        if jt_vis is not None:
            vis = torch.as_tensor(jt_vis).detach().to(dev).reshape(B, J).ne(0).to(torch.uint8).contiguous()
        uvd_img = torch.empty(B, J, 3, dtype=torch.float32, device=dev)
        dist = torch.empty(B, J, dtype=torch.float32, device=dev)
        diff = torch.empty(B, 3, dtype=torch.float32, device=dev)
        fx, fy, fu, fv = self.paras
        L.check(self.lib.awr_eval_feed(L.ptr(uvd), L.ptr(gt), L.ptr(ctr), L.ptr(Mt), L.ptr(cb), L.ptr(vis), B, J, float(self.img_size), fx, fy, fu,
                                       fv, self.flip, L.ptr(uvd_img), L.ptr(dist), L.ptr(diff), L.stream()), "awr_eval_feed")
        self._dist.append(dist); self._uvd.append(uvd_img); self._diff.append(diff)

    def feed(self, jt_uvd_pred, jt_xyz_gt, center_xyz, M, cube, jt_vis=0, skip_check=False):
        """Reference signature (eval_tool.py:20): one sample as numpy arrays (or tensors)."""
        a = lambda t: np.squeeze(t.detach().cpu().numpy() if isinstance(t, torch.Tensor) else np.asarray(t)).astype(np.float32)
        uvd, gt = a(jt_uvd_pred), a(jt_xyz_gt)
        if not skip_check:
            assert uvd.ndim == 2 and gt.ndim == 2
        vis = None if (np.isscalar(jt_vis) or np.ndim(jt_vis) == 0) and not np.any(jt_vis) else np.squeeze(np.asarray(jt_vis)).astype(bool)
        if vis is not None and vis.ndim == 0:
            vis = np.full(uvd.shape[0], bool(vis))
        self._staged.append((uvd, gt, a(center_xyz), a(M), a(cube), vis))
        if len(self._staged) >= 256:
            self._flush()

    def _flush(self):
        if not self._staged:
            return
        st, self._staged = self._staged, []
        stack = lambda k: torch.from_numpy(np.stack([s[k] for s in st]))
        vis = None
        if any(s[5] is not None for s in st):
            vis = torch.from_numpy(np.stack([s[5] if s[5] is not None else np.ones(self.num_kp, bool) for s in st]))
        self.feed_batch(stack(0), stack(1), stack(2), stack(3), stack(4), vis)

    # ---- results -------------------------------------------------------------------------------------------------------
    @property
    def jt_uvd_pred(self):
        """Per-sample (J,3) float32 arrays in image coordinates, as the reference's list (saved by train.py:219-221 / test.py:103-108)."""
        self._flush()
        return list(torch.cat(self._uvd).cpu().numpy()) if self._uvd else []

    @property
    def diff(self):
        self._flush()
        return list(torch.cat(self._diff).cpu().numpy()) if self._diff else []

    def errors(self):
        """(N,J) device tensor of Euclidean errors in mm (-1 = joint not visible)."""
        self._flush()
        return torch.cat(self._dist) if self._dist else torch.empty(0, self.num_kp, device=self.device)

    def get_measures(self):
        """(epe_mean_all, epe_median_all, auc_all, pck_curve_all, thresholds) as eval_tool.py:80-122."""
        dist = self.errors().contiguous()
        N, J = dist.shape
        thresholds = np.linspace(0, 50, 100)
        norm_factor = _trapz(np.ones_like(thresholds), thresholds)
        if N == 0:
            raise ValueError("get_measures: nothing was fed")
        dev = dist.device
        s = torch.empty(J, dtype=torch.float64, device=dev)
        cnt = torch.empty(J, dtype=torch.int32, device=dev)
        pck = torch.empty(J, len(thresholds), dtype=torch.int32, device=dev)
        L.check(self.lib.awr_eval_measures(L.ptr(dist), N, J, len(thresholds), float(thresholds[-1]), L.ptr(s), L.ptr(cnt), L.ptr(pck), L.stream()),
                "awr_eval_measures")
        # median per joint (np.median: mean of the two middle values): device sort with invisible entries pushed to +inf
        srt = torch.where(dist < 0, torch.full_like(dist, float("inf")), dist).sort(dim=0).values
        cnt_h, s_h, pck_h = cnt.cpu().numpy().astype(np.int64), s.cpu().numpy(), pck.cpu().numpy().astype(np.float64)
        lo = torch.as_tensor(np.maximum((cnt_h - 1) // 2, 0), device=dev).view(1, J)
        hi = torch.as_tensor(np.maximum(cnt_h // 2, 0), device=dev).view(1, J)
        med = ((srt.gather(0, lo) + srt.gather(0, hi)) * 0.5).view(J).cpu().numpy()
        means, medians, aucs, curves = [], [], [], []
        for j in range(J):
            if cnt_h[j] == 0:
                continue                                   # no valid measurement for this keypoint (eval_tool.py:98-100)
            means.append(np.float32(s_h[j] / cnt_h[j])); medians.append(np.float32(med[j]))
            curve = pck_h[j] / float(cnt_h[j])
            curves.append(curve)
            aucs.append(_trapz(curve, thresholds) / norm_factor)
        return (np.mean(np.array(means)), np.mean(np.array(medians)), np.mean(np.array(aucs)), np.mean(np.array(curves), 0), thresholds)

    def plot_pck(self, path, pck_curve_all, thresholds):
        import matplotlib.pyplot as plt                    # optional, only for the figure (eval_tool.py:124-135)
        fig = plt.figure()
        ax = fig.add_subplot(111)
        ax.plot(thresholds, pck_curve_all * 100, '-*', label='model')
        ax.set_xlabel('threshold in mm'); ax.set_ylabel('% of correct keypoints')
        plt.ylim([0.0, 100.0]); plt.grid(); plt.legend(loc='lower right')
        plt.savefig(path); plt.close()
