"""Debug (profile build only: `make PROFILE=1`, AWR_B200_LIB=.../libawr_b200_prof.so): event timeline of CTA 0 of conv_halo_kernel and a
knock-out study (no global stores / no statistics / epilogue drains TMEM only / no MMAs) for the layer shapes of ResNet18 at 32 frames."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from awr_b200 import _lib as L
lib = L.lib()
tl = lib.awr_debug_halo_timeline; tl.argtypes = [C.c_void_p, C.c_int]; tl.restype = C.c_int
buf = np.zeros(769, dtype=np.uint64)
TAGS = {1: "prologue done", 2: "pdl_wait done", 3: "final sync", 4: "exit", 10: "prod: halo load issued", 11: "prod: weight tile issued", 23: "mma: wait weights", 24: "mma: weights ready", 20: "mma: tmem stage free", 21: "mma: halo full",
        22: "mma: item issued (tfull commit)", 30: "epi0: tfull", 31: "epi1: tfull", 32: "epi0: tmem drained", 33: "epi1: tmem drained", 34: "epi0: stats done",
        35: "epi1: stats done", 36: "epi0: stores done", 37: "epi1: stores done"}


def make(N, Ci, Co, H, k, s, pad, transposed=0, mn=0, stats=False):
    Ho = (H - 1) * s - 2 * pad + k if transposed else (H + 2 * pad - k) // s + 1
    cin = Co if mn else Ci
    x = torch.randn(N, H, H, cin, device="cuda").bfloat16()
    w = torch.randn(k, k, Co, Ci, device="cuda").bfloat16()
    y = torch.empty(N, Ho, Ho, Ci if mn else Co, device="cuda", dtype=torch.bfloat16)
    st = L.acc_zeros(2 * Co, "cuda") if stats else None

    def step():
        if not mn:
            L.check(lib.awr_conv_tc(x.data_ptr(), w.data_ptr(), None, y.data_ptr(), None if st is None else st.data_ptr(), N, H, H, Ci, Ho, Ho, Co, k, k, s, pad,
                                    transposed, 1, Ci, Co * Ci, 0, 0, 0, L.stream()), "conv")
        else:
            L.check(lib.awr_conv_tc(x.data_ptr(), w.data_ptr(), None, y.data_ptr(), None, N, H, H, Co, Ho, Ho, Ci, k, k, s, pad,
                                    transposed, Ci, 1, Co * Ci, 0, 0, 0, L.stream()), "conv")
    step.keep = (x, w, y, st)
    return step


def chain_us(step, n=32, reps=10):
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        for _ in range(2):
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


def study(name, *a, **kw):
    step = make(*a, **kw)
    print(f"==== {name}")
    # knock-outs need a library built with -DAWR_CONV_KNOCKOUT (they perturb the MMA issue loop); otherwise every row repeats "as shipped"
    for flags, what in ((0, "as shipped"),) if not os.environ.get("AWR_KNOCKOUT") else ((0, "as shipped"), (2, "no statistics"), (1, "no global stores"),
                        (3, "no stores, no stats"), (4, "epilogue drains TMEM only"), (8, "no MMAs"), (12, "no MMAs, epilogue drains only")):
        tl(buf.ctypes.data, flags)
        print(f"   {what:32s}: {chain_us(step):7.2f} us / launch (PDL chain)")
    for flags in ((0, 8, 12) if os.environ.get("AWR_KNOCKOUT") else (0,)):
        tl(buf.ctypes.data, flags)
        step(); step(); torch.cuda.synchronize()
        tl(buf.ctypes.data, flags)          # reset the log
        step(); torch.cuda.synchronize()
        tl(buf.ctypes.data, 0)
        ev = sorted(((int(v) & ((1 << 48) - 1), int(v) >> 48) for v in buf[1:769] if int(v) != 0))
        print(f"   timeline of CTA 0, knock-out flags {flags} (cycles since kernel entry):")
        for t, tag in ev:
            print(f"     {t:8d}  {TAGS.get(tag, tag)}")


import sys
if len(sys.argv) < 2:
    study("layer1 fprop 3x3 64->64 @64 +stats", 32, 64, 64, 64, 3, 1, 1, stats=True)
    study("layer1 dgrad 3x3 64->64 @64", 32, 64, 64, 64, 3, 1, 1, transposed=1, mn=1)
    study("layer2 fprop 3x3 128->128 @32 +stats", 32, 128, 128, 32, 3, 1, 1, stats=True)
study("layer3 fprop 3x3 256->256 @16 +stats", 32, 256, 256, 16, 3, 1, 1, stats=True)
study("deconv3 fprop 256->256 @32->64 +stats", 32, 256, 256, 32, 4, 2, 1, transposed=1, stats=True)
