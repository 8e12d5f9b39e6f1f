"""Next-round probe: fixed cost of one tcgen05 conv launch inside a PDL-chained CUDA graph.

Chains of 64 dependent launches are captured into a graph and replayed; time per launch is reported for
  (a) awr_adam_tick          -- a 1-thread kernel: the launch floor of the chain itself,
  (b) a conv whose whole problem is ONE super-tile (1 x 16 x 16 x 64 -> 64, 3x3): prologue + one pipeline pass + epilogue,
  (c) the same with fused BatchNorm statistics (per-CTA atomics flush),
  (d) conv alternating with a small awr_bn_act (what a ResNet block does),
  (e) the headline 3x3 64->64 @64x64 N=32 layer (19 us net in profiles/r01_final2_conv_analysis.md, 6.9 us of tensor time).
(b)-(a) is the per-launch fixed cost that 69 conv launches per step pay.  Run:  gpurun -- 'python tools/probe_conv_floor.py'"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from awr_b200 import _lib as L

lib = L.lib()
dev = torch.device("cuda")
CH = 64


def chain_time(step, n=CH, reps=20):
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        for _ in range(3):
            step()
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for _ in range(n):
                step()
        g.replay(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            g.replay()
        e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / (reps * n)


def conv_step(N, H, C, stats):
    x = torch.randn(N, H, H, C, device=dev).bfloat16()
    w = (torch.randn(3, 3, C, C, device=dev) * 0.05).bfloat16()
    y = torch.empty(N, H, H, C, device=dev, dtype=torch.bfloat16)
    st = L.acc_zeros(2 * C, dev) if stats else None
    keep = (x, w, y, st)

    def step():
        L.check(lib.awr_conv_tc(x.data_ptr(), w.data_ptr(), None, y.data_ptr(), None if st is None else st.data_ptr(), N, H, H, C, H, H, C, 3, 3, 1, 1,
                                0, 1, C, C * C, 0, 0, 0, L.stream()), "conv")
    step.keep = keep
    return step


tick_buf = torch.zeros(1, device=dev)
t_tick = chain_time(lambda: L.check(lib.awr_adam_tick(tick_buf.data_ptr(), L.stream()), "tick"))
print(f"(a) 1-thread kernel chain              : {t_tick:6.2f} us / launch")
for tag, N, H, stats in (("(b) conv 1 super-tile", 1, 16, False), ("(c) conv 1 super-tile + BN statistics", 1, 16, True),
                         ("(e) conv 64->64 @64x64 N=32 + statistics", 32, 64, True)):
    t = chain_time(conv_step(N, H, 64, stats))
    print(f"{tag:39s}: {t:6.2f} us / launch   (fixed cost over the chain floor: {t - t_tick:5.2f} us)")
# (d) conv + bn_act alternating on a layer3-sized tensor (256 ch @16x16, N=32)
N, H, C = 32, 16, 256
cs = conv_step(N, H, C, True)
x, w, y, st = cs.keep
gamma, beta = torch.ones(C, device=dev), torch.zeros(C, device=dev)
rm, rv, nbt, mi = torch.zeros(C, device=dev), torch.ones(C, device=dev), torch.zeros((), dtype=torch.long, device=dev), torch.empty(2 * C, device=dev)
out = torch.empty_like(y)


def pair():
    cs()
    L.check(lib.awr_bn_act(y.data_ptr(), st.data_ptr(), gamma.data_ptr(), beta.data_ptr(), rm.data_ptr(), rv.data_ptr(), nbt.data_ptr(), mi.data_ptr(),
                           None, None, None, None, None, None, None, None, out.data_ptr(), L.BF16, N * H * H, C, 0.1, 1e-5, 1, 1, L.stream()), "bn")
t_pair = chain_time(pair, n=CH // 2)
t_conv = chain_time(cs)
print(f"(d) conv 256->256 @16x16 N=32 alone    : {t_conv:6.2f} us / launch;  conv + bn_act pair: {t_pair:6.2f} us  (bn_act adds {t_pair - t_conv:5.2f} us)")
