#!/usr/bin/env python
"""Headline benchmark of the AWR hot path on B200: training depth-frames/sec, device-timed.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--net resnet_18] [--batch 32]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...

A step = one full train.py:107-131 iteration (GT volume, backbone fwd, AWR head, joint+dense SmoothL1, backward,
[NCCL grad all-reduce], Adam) over one synthetic batch of `--batch` 128x128 depth crops per GPU, 14 joints.
Prints ONE JSON line (rank 0).  `value`: inputs resident in HBM.  `e2e`: same step through FusedTrainer.train_step_lagged
with pinned HOST buffers (H2D of every batch + D2H of every step's losses inside the timed region; losses are handed back one call late).
`--impl reference` times the CPU port of the reference's own path (oracle/) on the host cores.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "training depth-frames/sec (device-timed) ResNet18-AWR 128x128x14J"
UNIT = "frames/s"
J = 14
H = 128


CONV_TRAFFIC_PER_LAUNCH = 16158251      # bytes; ncu, headline config (see roofline.traffic_source)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--net", default="resnet_18")
    ap.add_argument("--batch", type=int, default=32, help="frames per GPU (weak scaling)")
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--img-size", type=int, default=128, help="depth crop side (128 = headline config; 256 = high-res stress config)")
    ap.add_argument("--cpu-steps", type=int, default=3, help="timed steps of the cpu_baseline leg (0 disables)")
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--layers", default="", help="write a per-layer conv timing table to gpurun_out/<name>")
    ap.add_argument("--no-gpu-eager-baseline", dest="gpu_eager_baseline", action="store_false",
                    help="skip the N=1 leg that times the oracle's torch ops on this GPU (eager cuDNN/cuBLAS, fp32 and bf16 autocast): a reported bar, never the product")
    ap.add_argument("--keep-grads", action="store_true", help="A/B: separate gradient fill per step instead of zeroing in the optimizer kernel")
    ap.add_argument("--no-parity", action="store_true", help="skip the C1 parity leg (UVD max-abs-diff vs the reference golden vector)")
    return ap.parse_args()


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=d["hbm_gbs"], tf_burst=d["bf16_tflops"], tf_sust=d.get("bf16_tflops_sustained", d["bf16_tflops"]), src="measured")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sust=1400.0, src="fallback")


# ---------------------------------------------------------------------------------------------------------
# CPU baseline / reference arm: the oracle's restatement of train.py:107-131 on the host cores
# ---------------------------------------------------------------------------------------------------------
def cpu_train_steps(net, ds, ks, batch, steps, warmup=1):
    import torch
    from oracle import awr_oracle as O        # checker / CPU baseline only
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    kind, n = net.split("_")
    sd = O.resnet_deconv_init(int(n), J, ds, 1) if kind == "resnet" else O.hourglass_init(int(n), J, 1)
    img, jt = O.synthetic_batch(batch, H, J, 0)
    state = {k: (torch.zeros_like(v), torch.zeros_like(v)) for k, v in sd.items()
             if v.is_floating_point() and not k.endswith(("running_mean", "running_var"))}
    times = []
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        _, lc, ld, _, _, grads, new_stats = O.loss_and_grads(sd, img, jt, net, ds, ks, 1.0, 1.0)
        with torch.no_grad():
            for k, g in grads.items():
                if g is not None:
                    O.adam_step(sd[k], g, state[k][0], state[k][1], it + 1)
            for k, v in new_stats.items():
                sd[k] = v
        dt = time.perf_counter() - t0
        if it >= warmup:
            times.append(dt)
    return batch * len(times) / sum(times), cores, times


def parity_c1(dev):
    """BASELINE.json configs[0] / the metric's second half: ResNet18-deconv + AWR head, 1x128x128 crop, 14 joints, batch 1, fp32 forward;
    max |UVD - reference| against the vector the UNMODIFIED reference produced (tests/golden/backbone_cases.pt, case 0).  The oracle is
    used only to regenerate the seeded state dict and input of that case (checker role)."""
    import torch
    import awr_b200
    from oracle import awr_oracle as O
    c = torch.load(os.path.join(ROOT, "tests", "golden", "backbone_cases.pt"))[0]
    sd = O.randomize_bn(O.resnet_deconv_init(18, c["J"], c["ds"], c["seed"], head_std=c["head_std"]), c["seed"] + 1)
    m = awr_b200.get_deconv_net(18, c["J"], c["ds"], precision="fp32")
    m.load_state_dict(sd, strict=True)
    m = m.to(dev).eval()
    img, _ = O.synthetic_batch(c["B"], c["H"], c["J"], c["seed"] + 2)
    with torch.no_grad():
        pred = m(img.to(dev))
        uvd = awr_b200.FeatureModule().offset2joint_softmax(pred, img.to(dev), c["ks"])
    diff = (uvd.cpu() - c["eval_uvd"][0]).abs().max().item()
    mm = mean3d_diff_mm(uvd.cpu(), c["eval_uvd"][0], c["B"], c["J"], c["H"])
    return {"config": "resnet_18-deconv + AWR head, 1x128x128 depth crop, 14 joints, batch 1, fp32 forward (eval-mode BN)",
            "uvd_max_abs_diff": diff, "tolerance": 1e-3, "pass": diff < 1e-3 and mm < 0.05,
            "mean_3d_error_diff_mm": mm, "tolerance_mm": 0.05,
            "against": "unmodified reference on CPU, recorded in tests/golden/backbone_cases.pt[0] by tests/golden/make_golden.py"}


def parity_headline(dev):
    """The bf16 tensor-core path the throughput number is quoted on, at the headline batch (ResNet18, 32 frames): max |UVD - reference| and
    the mean-3-D-error difference against the vector the unmodified reference produced (tests/golden/trajectory.pt["headline"], eval- and
    train-mode BN), for our fp32 and bf16 forward; stock torch.autocast(bfloat16) over the oracle's torch ops on the same case is printed
    beside it as the yardstick for what bf16 storage costs anyone.  Oracle: checker role (state dict, input, autocast yardstick)."""
    import torch
    import awr_b200
    from oracle import awr_oracle as O
    c = torch.load(os.path.join(ROOT, "tests", "golden", "trajectory.pt"))["headline"]
    sd = O.randomize_bn(O.resnet_deconv_init(18, c["J"], c["ds"], c["seed"], head_std=c["head_std"]), c["seed"] + 1)
    img, _ = O.synthetic_batch(c["B"], c["H"], c["J"], c["seed"] + 2)
    img = img.to(dev)
    fm = awr_b200.FeatureModule()
    out = {"config": f"resnet_18-deconv + AWR head, {c['B']}x1x{c['H']}x{c['H']} depth crops, {c['J']} joints (the headline batch)",
           "against": "unmodified reference on CPU, recorded in tests/golden/trajectory.pt['headline'] by tests/golden/make_golden2.py"}

    def record(key, uvd, mode):
        uvd = uvd.float().cpu()
        out[f"uvd_max_abs_diff_{key}_{mode}"] = round((uvd - c[mode + "_uvd"]).abs().max().item(), 7)
        out[f"mean_3d_error_diff_mm_{key}_{mode}"] = round(mean3d_diff_mm(uvd, c[mode + "_uvd"], c["B"], c["J"], c["H"]), 5)
    for prec in ("fp32", "bf16"):
        m = awr_b200.get_deconv_net(18, c["J"], c["ds"], precision=prec)
        m.load_state_dict(sd, strict=True)
        m = m.to(dev)
        for mode in ("eval", "train"):
            m.train(mode == "train")
            with torch.no_grad():
                record(prec, fm.offset2joint_softmax(m(img), img, c["ks"]), mode)
    sdc = {k: v.to(dev) for k, v in sd.items()}
    for mode in ("eval", "train"):
        with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
            pred = O.backbone_forward(sdc, img, "resnet_18", c["ds"], training=(mode == "train"))
        record("torch_autocast_bf16", O.offset2joint_softmax(pred.float(), img, c["ks"]), mode)
    out["uvd_max_abs_diff_bf16"] = out["uvd_max_abs_diff_bf16_eval"]
    out["mean_3d_error_diff_mm_bf16"] = out["mean_3d_error_diff_mm_bf16_eval"]
    out["tolerance_fp32"] = 1e-3
    out["tolerance_bf16"] = "no worse than 1.5x torch.autocast(bf16) on the same case, or 3e-2 (eval) / 5e-2 (train) UVD; tests/test_backbone_gpu.py"
    out["pass"] = bool(out["uvd_max_abs_diff_fp32_eval"] < 1e-3 and out["uvd_max_abs_diff_fp32_train"] < 1e-3
                       and out["uvd_max_abs_diff_bf16_eval"] < max(3e-2, 1.5 * out["uvd_max_abs_diff_torch_autocast_bf16_eval"])
                       and out["uvd_max_abs_diff_bf16_train"] < max(5e-2, 1.5 * out["uvd_max_abs_diff_torch_autocast_bf16_train"]))
    return out


def preprocess_leg(dev, batch, img_size):
    """SURVEY 8 f.2 (the stage that feeds the step): raw 480x640 frames -> augmented, normalised crops through awr_b200.preprocess.train_batch.
    Device time of the kernel (CUDA events, params resident) and host time of the per-frame float64 geometry, separately.  The oracle only
    generates the synthetic raw frames."""
    import numpy as np
    import torch
    from awr_b200 import preprocess as PP, _lib as L
    from oracle import awr_oracle as O
    frames, jt_xyz, center_xyz = O.augment_case_inputs(batch, 31)
    rs = np.random.RandomState(23455)
    augs = [PP.random_aug(rs, 10, 0.1, 180) for _ in range(batch)]
    cube = np.asarray([300, 300, 300])
    fr = torch.from_numpy(frames).to(dev)
    t0 = time.perf_counter()
    for _ in range(5):
        geo = PP.train_batch_geometry(jt_xyz, center_xyz, cube, img_size, O.NYU_PARAS, O.NYU_FLIP, augs)
    host_ms = 1e3 * (time.perf_counter() - t0) / 5
    params = torch.from_numpy(geo[0]).to(dev)
    out = torch.empty(batch, 1, img_size, img_size, device=dev)
    call = lambda: L.check(L.lib().awr_crop_augment_normalize(L.ptr(fr), 0, batch, frames.shape[1], frames.shape[2], L.ptr(params), img_size, L.ptr(out),
                                                              L.stream()), "awr_crop_augment_normalize")
    for _ in range(3):
        call()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(20):
        call()
    e1.record()
    torch.cuda.synchronize()
    us = 1e3 * e0.elapsed_time(e1) / 20
    return {"what": f"{batch} raw 480x640 float32 frames -> crop + random translate/scale/rotate (cv2-exact warps) + normalize, {img_size}x{img_size}",
            "kernel_us_per_batch": round(us, 2), "kernel_frames_per_s": round(batch / (us * 1e-6), 0), "ctas": batch * 8,
            "host_geometry_ms_per_batch": round(host_ms, 2), "host_frames_per_s_one_core": round(batch / (host_ms * 1e-3), 0),
            "ops": {str(k): sum(1 for a_ in augs if a_[0] == k) for k in ("trans", "scale", "rot", None)}}


def mean3d_diff_mm(uvd_ours, uvd_ref, B, J, img_size):
    """north_star's second bound: |mean 3-D error(ours) - mean 3-D error(reference)| in mm on the same synthetic batch, through the
    reference's UVD -> XYZ chain (util/eval_tool.py:34-49) with a synthetic NYU-like crop geometry and ground truth (checker role)."""
    import torch
    from oracle import awr_oracle as O
    _, gt, center, M, cube, _ = O.eval_case_inputs(B, J, 77, img_size)
    t = lambda a: torch.from_numpy(a)
    e_ours = O.mean_3d_error_mm(uvd_ours.float(), t(gt), t(center), t(M), t(cube), img_size)
    e_ref = O.mean_3d_error_mm(uvd_ref.float(), t(gt), t(center), t(M), t(cube), img_size)
    return abs(e_ours - e_ref)


def gpu_eager_steps(net, ds, ks, batch, dev, steps=10, autocast=False):
    """The oracle's functional torch ops (cuDNN conv, eager elementwise, autograd) on the GPU: what `train.py` costs on this box without
    our kernels.  Same step as cpu_train_steps (Adam through torch ops).  Reported for scale only."""
    import torch
    from oracle import awr_oracle as O
    kind, n = net.split("_")
    sd = O.resnet_deconv_init(int(n), J, ds, 1) if kind == "resnet" else O.hourglass_init(int(n), J, 1)
    sd = {k: v.to(dev) for k, v in sd.items()}
    img, jt = O.synthetic_batch(batch, H, J, 0)
    img, jt = img.to(dev), jt.to(dev)
    state = {k: (torch.zeros_like(v), torch.zeros_like(v)) for k, v in sd.items()
             if v.is_floating_point() and not k.endswith(("running_mean", "running_var"))}

    def one(it):
        with torch.autocast("cuda", dtype=torch.bfloat16, enabled=autocast):
            _, lc, ld, _, _, grads, new_stats = O.loss_and_grads(sd, img, jt, net, ds, ks, 1.0, 1.0)
        with torch.no_grad():
            for k, g in grads.items():
                if g is not None:
                    O.adam_step(sd[k], g.float(), state[k][0], state[k][1], it + 1)
            for k, v in new_stats.items():
                sd[k] = v
    for it in range(3):
        one(it)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for it in range(steps):
        one(3 + it)
    e1.record()
    torch.cuda.synchronize()
    return batch * steps / (e0.elapsed_time(e1) * 1e-3)


def run_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    ds, ks = 2, (1.0 if a.net.startswith("resnet") else 0.4)
    steps = max(1, min(a.steps, 4))
    fps, cores, times = cpu_train_steps(a.net, ds, ks, a.batch, steps, warmup=min(a.warmup, 1))
    sample = f"{steps} full train steps (after 1 warm-up) of the same workload at batch {a.batch}, fp32, torch CPU ops, {cores} threads"
    line = {"impl": "reference", "metric": METRIC, "value": round(fps, 3), "unit": UNIT, "n_gpus": a.gpus, "steps": steps,
            "warmup": min(a.warmup, 1), "ms_per_step": round(1e3 * sum(times) / len(times), 2), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(a, 1),
            "cpu_baseline": {"value": round(fps, 3), "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": round(fps, 3), "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def workload_config(a, world):
    return {"workload": f"{a.net}-deconv AWR full train step, {H}x{H}x1 depth crops, {J} joints, batch {a.batch}/GPU",
            "global_batch": a.batch * world, "img_size": H, "joints": J, "downsample": 2, "optimizer": "Adam lr 1e-3",
            "parallelism": f"dp{world}"}


# ---------------------------------------------------------------------------------------------------------
# clocks sampler (B200_PROFILING.md recipe)
# ---------------------------------------------------------------------------------------------------------
class Clocks:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), [x.strip() for x in line.split(",")]))

    def stop(self):
        if self.proc is not None:
            self.proc.terminate()

    def summary(self, t0, t1):
        rows = [r for t, r in self.rows if t0 <= t <= t1 and len(r) >= 9] or [r for _, r in self.rows if len(r) >= 9]
        if not rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        sm = [float(r[1]) for r in rows]
        reasons = set()
        for r in rows:
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": float(rows[0][2]), "reasons": sorted(reasons), "samples": len(rows),
                "power_w_max": max(float(r[3]) for r in rows)}


# ---------------------------------------------------------------------------------------------------------
# per-kernel-class device timing of one eager step (CUDA events on the launch stream, queue pre-filled behind a spin)
# ---------------------------------------------------------------------------------------------------------
def classify_step(tr, steps, layers=None):
    import torch
    from awr_b200 import _lib as L
    pl = tr.plan
    agg = {}
    for _ in range(steps):
        evs = []
        s = L.stream()
        torch.cuda._sleep(int(40e6))          # ~20 ms: lets the host enqueue the whole step so events are back-to-back
        pl.arena_used().zero_(); tr.store.grads.zero_()
        seq = [(f, m) for f, m in zip(pl.fwd, pl.fwd_meta)]
        hd = tr.head
        lib = tr.lib
        head_f = lambda st: L.check(lib.awr_head_fwd(hd.pred.data_ptr(), L.F32, pl.img.data_ptr(), tr.jt.data_ptr(), tr.uvd.data_ptr(),
                                                     tr.loss.data_ptr(), tr.ws.data_ptr(), tr.B, tr.J, tr.F, tr.H, tr.ks, st), "head_fwd")
        head_b = lambda st: L.check(lib.awr_head_bwd(hd.pred.data_ptr(), L.F32, pl.img.data_ptr(), tr.jt.data_ptr(), tr.uvd.data_ptr(),
                                                     tr.ws.data_ptr(), None, None, hd.dpred.data_ptr(), tr.B, tr.J, tr.F, tr.H, tr.ks,
                                                     tr.cw, tr.dw, st), "head_bwd")
        P = tr.F * tr.F
        fwd_bytes = tr.B * (4 * tr.J * P * 4 + P * 4) + tr.B * tr.J * 24
        bwd_bytes = tr.B * (2 * 4 * tr.J * P * 4 + P * 4) + tr.B * tr.J * 24
        seq.append((head_f, ("head_fwd", 0, fwd_bytes)))
        seq.append((head_b, ("head_bwd", 0, bwd_bytes)))
        seq += [(f, m) for f, m in zip(pl.bwd, pl.bwd_meta)]
        for f, m in seq:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); f(s); e1.record()
            evs.append((m, e0, e1))
        tr._opt()
        torch.cuda.synchronize()
        for m, e0, e1 in evs:
            tag, fl, nb = m[0], m[1], m[2]
            a = agg.setdefault(tag, [0.0, 0, 0, 0])
            dt = e0.elapsed_time(e1)
            a[0] += dt; a[1] += fl; a[2] += nb; a[3] += 1
            if layers is not None and len(m) > 3 and m[3]:
                l = layers.setdefault((tag, m[3]), [0.0, fl, 0])
                l[0] += dt; l[2] += 1
    return agg


def conv_class_time(tr, reps=5):
    """Device time of ALL tensor-core conv launches of one step (fprop, dgrad, wgrad; the other kernels skipped), issued back to back
    between two CUDA events on the launch stream.  Buffers are static, so the sequence is valid on its own; the ~1 GB of activations it
    walks exceeds the 126 MB L2.  Returns (ms per step, flops per step, launches per step)."""
    import torch
    from awr_b200 import _lib as L
    pl = tr.plan
    seq = [(f, m) for f, m in list(zip(pl.fwd, pl.fwd_meta)) + list(zip(pl.bwd, pl.bwd_meta)) if m[0].startswith("conv_")]
    s = L.stream()
    for f, _ in seq:
        f(s)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda._sleep(int(20e6))
    e0.record()
    for _ in range(reps):
        for f, _m in seq:
            f(s)
    e1.record()
    torch.cuda.synchronize()
    tr.store.grads.zero_()
    return e0.elapsed_time(e1) / reps, sum(m[1] for _, m in seq), len(seq)


def head_pair_time(tr, reps=8):
    """Cold-cache device time of the fused head+loss forward + backward kernels (L2 flushed by a 512 MB fill before every pair)."""
    import torch
    from awr_b200 import _lib as L
    pl, hd, lib = tr.plan, tr.head, tr.lib
    flush = torch.empty(512 * 1024 * 1024 // 4, dtype=torch.float32, device=tr.device)
    tot = 0.0
    for _ in range(reps):
        torch.cuda._sleep(int(4e6))
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s = L.stream()
        e0.record()
        L.check(lib.awr_head_fwd(hd.pred.data_ptr(), L.F32, pl.img.data_ptr(), tr.jt.data_ptr(), tr.uvd.data_ptr(), tr.loss.data_ptr(),
                                 tr.ws.data_ptr(), tr.B, tr.J, tr.F, tr.H, tr.ks, s), "head_fwd")
        L.check(lib.awr_head_bwd(hd.pred.data_ptr(), L.F32, pl.img.data_ptr(), tr.jt.data_ptr(), tr.uvd.data_ptr(), tr.ws.data_ptr(), None, None,
                                 hd.dpred.data_ptr(), tr.B, tr.J, tr.F, tr.H, tr.ks, tr.cw, tr.dw, s), "head_bwd")
        e1.record()
        torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    del flush
    return tot / reps


def main():
    global H, METRIC
    a = parse()
    H = a.img_size
    if a.net != "resnet_18" or H != 128:
        METRIC = f"training depth-frames/sec (device-timed) {a.net}-AWR {H}x{H}x{J}J"
    if a.impl == "reference":
        return run_reference(a)

    import torch
    import awr_b200
    from awr_b200.trainer import FusedTrainer

    from awr_b200 import dp
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (sm_100a kernels; there is no CPU fallback)")
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    rank, local, world = dp.init_from_env("nccl", dev)
    ds, ks = 2, (1.0 if a.net.startswith("resnet") else 0.4)

    torch.manual_seed(1)
    kind, n = a.net.split("_")
    net = awr_b200.get_deconv_net(int(n), J, ds, precision=a.precision) if kind == "resnet" else awr_b200.PoseNet(a.net, J, precision=a.precision)
    net = net.to(dev)
    tr = FusedTrainer(net, a.batch, H, ks, 1.0, 1.0, lr=1e-3, world_size=world, use_graph=not a.no_graph, keep_grads=a.keep_grads)
    tr.broadcast_parameters(0)

    # synthetic batch, per-rank seed; a few distinct batches rotate through the e2e leg
    g = torch.Generator().manual_seed(1000 + rank)
    yy, xx = torch.meshgrid(torch.linspace(-1, 1, H), torch.linspace(-1, 1, H), indexing="ij")

    def make_batch():
        cx, cy = (torch.rand(a.batch, 1, 1, generator=g) - 0.5) * 0.4, (torch.rand(a.batch, 1, 1, generator=g) - 0.5) * 0.4
        rx, ry = 0.45 + 0.3 * torch.rand(a.batch, 1, 1, generator=g), 0.45 + 0.3 * torch.rand(a.batch, 1, 1, generator=g)
        inside = ((xx - cx) / rx) ** 2 + ((yy - cy) / ry) ** 2 < 1.0
        depth = (0.5 * (xx - cx) + 0.3 * (yy - cy) + 0.05 * torch.randn(a.batch, H, H, generator=g)).clamp(-1.0, 0.98)
        img = torch.where(inside, depth, torch.ones_like(depth)).unsqueeze(1).contiguous().float()
        jt = (torch.rand(a.batch, J, 3, generator=g) - 0.5).float()
        return img.pin_memory(), jt.pin_memory()

    host_batches = [make_batch() for _ in range(4)]
    dev_batches = [(i.to(dev), j.to(dev)) for i, j in host_batches]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident leg (value) --------------------------------------------------------------------
    for w in range(max(a.warmup, 3)):
        tr.load_batch(*dev_batches[w % 4]); tr.run_step()
    clk = Clocks(local)
    if rank == 0:
        clk.start()
        time.sleep(0.3)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for k in range(a.steps):
        tr.load_batch(*dev_batches[k % 4])      # device->device 2 MB copy: the batch of this step, already in HBM
        tr.run_step()
    e1.record()
    barrier()
    t1 = time.perf_counter()
    ms = e0.elapsed_time(e1)
    ms = dp.max_over_ranks(ms, dev)
    value = a.batch * world * a.steps / (ms * 1e-3)

    # ---- end-to-end leg: host pinned batch -> train_step -> host losses ----------------------------------
    for w in range(3):
        tr.train_step(*host_batches[w % 4])
    barrier()
    e0.record()
    for k in range(a.steps):                     # every step: H2D of its batch (pinned host), the step, D2H of its two losses
        r = tr.train_step_lagged(*host_batches[k % 4])
        if r is not None:
            lc, ld = r
    lc, ld = tr.collect()                         # the last step's losses: read inside the timed region too
    e1.record()
    barrier()
    t2 = time.perf_counter()
    ms_e2e = e0.elapsed_time(e1)
    ms_e2e = dp.max_over_ranks(ms_e2e, dev)
    e2e = a.batch * world * a.steps / (ms_e2e * 1e-3)
    h2d = host_batches[0][0].numel() * 4 + host_batches[0][1].numel() * 4
    clocks = None
    if rank == 0:
        clk.stop()
        clocks = clk.summary(t0, t2)

    # ---- the data-parallel exchange alone: all-reduce of a gradient-sized buffer (all ranks) -------------------
    allreduce = None
    if world > 1:
        try:
            scratch = torch.zeros_like(tr.store.grads)
            ar_ms, ar_bw = dp.allreduce_busbw(scratch, reps=10)
            allreduce = {"bytes": scratch.numel() * 4, "ms": round(ar_ms, 4), "busbw_gbs": round(ar_bw, 1), "nvlink_peak_gbs_per_direction": 900,
                         "how": "10 NCCL sum all-reduces of a gradient-sized fp32 buffer alone, CUDA events, max over ranks; in the step the buffer "
                                "goes in buckets as backward finalises it (captured in the step's CUDA graph), all but the last, "
                                "sub-megabyte bucket overlapping the rest of backward",
                         "buckets_bytes": [4 * (b - a_) for a_, b in (tr.plan.bucket_range(i) for i in range(len(tr.plan.bwd_splits) + 1))],
                         "nccl_max_ctas": os.environ.get("NCCL_MAX_CTAS")}
            del scratch
        except Exception as e:              # never lose the bench line to a reporting leg
            allreduce = {"error": f"{type(e).__name__}: {e}"[:200]}

    # ---- roofline of the dominant kernels (rank 0, eager step, events around every launch) ----------------
    roof = roof_head = None
    classes = {}
    if rank == 0:
        pk = peaks()
        layers = {} if a.layers else None
        agg = classify_step(tr, 3, layers)
        if layers:
            os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
            with open(os.path.join(ROOT, "gpurun_out", a.layers), "w") as f:
                f.write("| kernel | layer | launches/step | us/launch | TFLOP/s |\n|---|---|---:|---:|---:|\n")
                for (tag, det), (ms_, fl, n_) in sorted(layers.items(), key=lambda kv: -kv[1][0]):
                    us = 1e3 * ms_ / n_
                    f.write(f"| {tag} | {det} | {n_ // 3} | {us:.1f} | {fl / (us * 1e-6) / 1e12:.1f} |\n")
        tot = sum(v[0] for v in agg.values())
        classes = {k: {"ms_per_step": round(v[0] / 3, 4), "share": round(v[0] / tot, 4), "launches_per_step": v[3] // 3} for k, v in
                   sorted(agg.items(), key=lambda kv: -kv[1][0])}
        # dominant kernel class: the tcgen05 implicit-GEMM convolutions, timed back to back (no per-launch event overhead)
        cms, cfl, cn = conv_class_time(tr)
        ach = cfl / (cms * 1e-3) / 1e12
        roof = {"kernel": "tcgen05 conv/deconv implicit-GEMM kernels (conv_halo_kernel, conv_tc_kernel, wgrad_tc_kernel: fprop+dgrad+wgrad)",
                "bound": "tensor", "achieved": round(ach, 2), "peak": pk["tf_sust"], "unit": "TFLOP/s", "frac": round(ach / pk["tf_sust"], 4),
                "traffic": CONV_TRAFFIC_PER_LAUNCH if (a.net == "resnet_18" and H == 128 and a.batch == 32) else None,
                "traffic_source": "mean dram__bytes_read+write per launch over the 18 conv launches captured with ncu --set full "
                                  "(profiles/r02_final_halo_full.md, r02_final_wgrad_full.md); the 64-channel 64x64 layers read their 16.9 MB "
                                  "input once, writes stay in L2",
                "peak_source": pk["src"] + " (sustained cuBLAS bf16 8192^3)", "launches_per_step": cn,
                "avg_launch_us": round(1e3 * cms / cn, 2), "ms_per_step": round(cms, 4), "share_of_step": round(cms / (ms / a.steps), 4),
                "flops_per_step": cfl, "how": "all conv launches of one step issued back to back between two CUDA events, 5 repetitions"}
        hv = [agg["head_fwd"], agg["head_bwd"]]
        hb = sum(v[2] for v in hv) // 3
        hms = head_pair_time(tr)
        hach = hb / (hms * 1e-3) / 1e9
        roof_head = {"kernel": "fused AWR head+loss (head_fwd_kernel + head_bwd_kernel)", "bound": "hbm", "achieved": round(hach, 1), "peak": pk["hbm"],
                     "unit": "GB/s", "frac": round(hach / pk["hbm"], 4), "traffic": 61189632, "peak_source": pk["src"] + " (copy)",
                     "pair_us": round(1e3 * hms, 2), "bytes_per_step": hb, "how": "cold L2 (512 MB fill before each fwd+bwd pair), 8 repetitions; "
                     "traffic = ncu dram bytes of the pair (profiles/r01_final2_head_full.md)"}

    # ---- CPU baseline (rank 0, N=1 only) -------------------------------------------------------------------
    cpu = None
    if rank == 0 and world == 1 and a.cpu_steps > 0:
        fps, cores, times = cpu_train_steps(a.net, ds, ks, a.batch, a.cpu_steps)
        cpu = {"value": round(fps, 3), "unit": UNIT, "cores": cores, "kind": "port",
               "sample": f"{a.cpu_steps} full train steps (1 warm-up) at batch {a.batch}, fp32, oracle/awr_oracle.py on torch CPU ops"}

    parity = eager = None
    if rank == 0 and world == 1 and not a.no_parity and a.net == "resnet_18" and H == 128:
        try:
            parity = parity_c1(dev)
        except Exception as e:              # a checker leg must never cost the bench line
            parity = {"error": f"{type(e).__name__}: {e}"[:200]}
        try:
            parity["headline_batch"] = parity_headline(dev)
            parity["uvd_max_abs_diff_bf16"] = parity["headline_batch"]["uvd_max_abs_diff_bf16"]
            parity["mean_3d_error_diff_mm_bf16"] = parity["headline_batch"]["mean_3d_error_diff_mm_bf16"]
        except Exception as e:
            parity["headline_batch"] = {"error": f"{type(e).__name__}: {e}"[:200]}
    pre = None
    if rank == 0 and world == 1 and not a.no_parity:
        try:
            pre = preprocess_leg(dev, a.batch, H)
        except Exception as e:              # a reporting leg must never cost the bench line
            pre = {"error": f"{type(e).__name__}: {e}"[:200]}
    if rank == 0 and world == 1 and a.gpu_eager_baseline:
        eager = {"unit": UNIT, "what": "oracle's torch ops on this GPU (eager, cuDNN), same train step, device-resident batch"}
        for key, ac in (("fp32", False), ("bf16_autocast", True)):
            try:
                eager[key] = round(gpu_eager_steps(a.net, ds, ks, a.batch, dev, autocast=ac), 1)
            except Exception as e:          # a baseline leg must never cost the bench line
                eager[key] = f"failed: {type(e).__name__}: {e}"[:200]

    if rank == 0:
        act_bytes = sum(op.y.t.numel() * op.y.t.element_size() for op in tr.plan.ops if hasattr(op, "y"))
        cfg = workload_config(a, world)
        cfg.update({"precision": a.precision, "l2": f"no flush needed: per-step conv outputs alone are {act_bytes / 2**20:.0f} MiB (> 126 MB L2), "
                    "rotating 4 input batches", "cuda_graph": not a.no_graph, "e2e_api": "awr_b200.trainer.FusedTrainer.train_step_lagged(pinned host img, jt) -> losses of the previous step; collect() at the end"})
        line = {"metric": METRIC, "value": round(value, 1), "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": max(a.warmup, 3),
                "ms_per_step": round(ms / a.steps, 4), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": a.precision, "data": "synthetic", "config": cfg, "clocks": clocks,
                "e2e": {"value": round(e2e, 1), "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 8,
                        "ms_per_step": round(ms_e2e / a.steps, 4)},
                "gpu_launches": tr.launches_per_step * a.steps, "launches_per_step": tr.launches_per_step,
                "roofline": roof, "roofline_head": roof_head, "kernel_classes": classes, "cpu_baseline": cpu,
                "loss": {"coord": lc, "dense": ld}, "parity": parity}
        if allreduce is not None:
            line["allreduce"] = allreduce
        if eager is not None:
            line["gpu_eager_baseline"] = eager
        if pre is not None:
            line["preprocess"] = pre
        print(json.dumps(line), flush=True)
    if world > 1:
        # the captured step graph holds NCCL kernels: release it, then leave without the communicator teardown (destroy_process_group can
        # block on a communicator that graph-captured collectives used); every rank exits 0 after the line is out
        sys.stdout.flush()
        dist.barrier()
        tr.release()
        os._exit(0)


if __name__ == "__main__":
    main()
