"""Device timing of the fused head+loss kernels alone, L2 flushed between launches (cold) and back-to-back (warm)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from awr_b200.feature_tool import head_loss_forward, head_loss_backward
from oracle import awr_oracle as O
for (B, J, Fs, H) in [(32, 14, 64, 128), (64, 14, 128, 256)]:
    img, jt = O.synthetic_batch(B, H, J, 3)
    img, jt = img.cuda(), jt.cuda()
    pred = torch.randn(B, 4 * J, Fs, Fs, device="cuda")
    flush = torch.empty(256 * 1024 * 1024 // 4, device="cuda")
    uvd, loss, ws = head_loss_forward(pred, img, jt, 1.0)
    dp = torch.empty_like(pred)
    P = Fs * Fs
    fb = B * (4 * J * P * 4 + P * 4) + B * J * 24
    bb = B * (2 * 4 * J * P * 4 + P * 4) + B * J * 24
    for name, fn, nbytes in [("fwd", lambda: head_loss_forward(pred, img, jt, 1.0, ws=ws, uvd_out=uvd, loss_out=loss), fb),
                             ("bwd", lambda: head_loss_backward(pred, img, jt, uvd, ws, 1.0, 1.0, 1.0, dpred=dp), bb)]:
        for mode in ("cold", "warm"):
            ts = []
            for _ in range(12):
                torch.cuda._sleep(2_000_000)                 # keep the GPU busy while the host enqueues (events then sit back-to-back)
                if mode == "cold": flush.zero_()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                if mode == "warm": fn()
                e0.record(); fn(); e1.record(); torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1) * 1e3)
            ts = sorted(ts)[2:-2]
            us = sum(ts) / len(ts)
            print(f"B{B} F{Fs} {name} {mode}: {us:7.2f} us  {nbytes / us / 1e3:7.1f} GB/s  ({nbytes / 1e6:.1f} MB algorithmic)")
