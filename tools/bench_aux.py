"""Device timing of the auxiliary kernels around the hot path: depth preprocessing and evaluation.
CUDA events, 20 repetitions after warm-up, L2 flushed between repetitions.  Prints one line per kernel with its algorithmic bytes."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import awr_b200
from awr_b200 import preprocess as PP
from oracle import awr_oracle as O

flush = torch.empty(256 * 1024 * 1024 // 4, device="cuda")


def timeit(fn, reps=20):
    for _ in range(3): fn()
    ts = []
    for _ in range(reps):
        torch.cuda._sleep(2_000_000); flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    ts = sorted(ts)[2:-2]
    return sum(ts) / len(ts)

N, D = 32, 128
frames, centers, cubes = O.preprocess_case_inputs(N, 3)
cz = centers[:, 2].astype(np.float64)
P, _ = PP.crop_params(centers, cz, cubes, D, O.NYU_PARAS)
box_bytes = float((P[:, 2] * P[:, 3]).sum()) * 4                                   # the crop boxes actually gathered (upper bound: every box pixel once)
fr = torch.from_numpy(frames).cuda()
d16 = frames.astype(np.uint16)
bgr = torch.from_numpy(np.stack([(d16 & 255).astype(np.uint8), (d16 >> 8).astype(np.uint8), np.zeros_like(d16, dtype=np.uint8)], -1)).cuda()
params = torch.from_numpy(P).cuda(); out = torch.empty(N, 1, D, D, device="cuda")
from awr_b200 import _lib as L
lib = L.lib()
for name, src, fmt in (("crop_normalize f32 frames", fr, 0), ("crop_normalize packed BGR", bgr, 1)):
    us = timeit(lambda: L.check(lib.awr_crop_normalize(src.data_ptr(), fmt, N, 480, 640, params.data_ptr(), D, out.data_ptr(), L.stream()), "crop"))
    alg = min(box_bytes, N * D * D * 4.0) + 3 * N * D * D * 4                        # <= one source pixel per output pixel + write, re-read, write of the crop
    print(f"{name}: N={N} 480x640 -> {D}x{D}: {us:.2f} us ({N / us * 1e6:.0f} frames/s), algorithmic {alg / 1e6:.2f} MB -> {alg / us / 1e3:.1f} GB/s")
# evaluation
J = 14
uvd, gt, center, M, cube, _ = O.eval_case_inputs(8252, J, 1)
t = lambda a: torch.from_numpy(a).cuda()
tu, tg, tc, tm, tcb = t(uvd[:32]), t(gt[:32]), t(center[:32]), t(M[:32]), t(cube[:32])
ev = awr_b200.EvalUtil(128, O.NYU_PARAS, O.NYU_FLIP, J)
uo, do, df = torch.empty(32, J, 3, device="cuda"), torch.empty(32, J, device="cuda"), torch.empty(32, 3, device="cuda")
us = timeit(lambda: L.check(lib.awr_eval_feed(tu.data_ptr(), tg.data_ptr(), tc.data_ptr(), tm.data_ptr(), tcb.data_ptr(), None, 32, J, 128.0, *[float(p) for p in O.NYU_PARAS],
                                               float(O.NYU_FLIP), uo.data_ptr(), do.data_ptr(), df.data_ptr(), L.stream()), "feed"))
print(f"eval_feed: B=32 J={J}: {us:.2f} us per step (the reference: 32 x 5 .cpu() copies + numpy per frame)")
ev.feed_batch(t(uvd), t(gt), t(center), t(M), t(cube))
dist = ev.errors().contiguous()
s, c, pk = torch.empty(J, dtype=torch.float64, device="cuda"), torch.empty(J, dtype=torch.int32, device="cuda"), torch.empty(J, 100, dtype=torch.int32, device="cuda")
us = timeit(lambda: L.check(lib.awr_eval_measures(dist.data_ptr(), dist.shape[0], J, 100, 50.0, s.data_ptr(), c.data_ptr(), pk.data_ptr(), L.stream()), "meas"))
print(f"eval_measures: N=8252 (NYU test set) J={J}, 100 thresholds: {us:.2f} us")
import time
t0 = time.perf_counter(); r = ev.get_measures(); torch.cuda.synchronize(); t1 = time.perf_counter()
print(f"EvalUtil.get_measures() wall: {(t1 - t0) * 1e3:.2f} ms  (MPE {r[0]:.3f} mm, AUC {r[2]:.4f})")
