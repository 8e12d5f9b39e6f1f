"""Debug: role-level cycle breakdown of conv_tc_kernel (needs a library built with -DAWR_CONV_PROFILE: `make PROFILE=1`)."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from awr_b200 import _lib as L
lib = L.lib()
fn = lib.awr_debug_conv_profile; fn.argtypes = [C.c_void_p, C.c_int]; fn.restype = C.c_int
buf = np.zeros(148 * 8, dtype=np.uint64)
# (the halo-tile kernel has its own event timeline: tools/dbg_halo_timeline.py)

def run(name, N, Ci, Co, H, k, s, pad, transposed=0, mn=0, stats=False, reps=3):
    Ho = (H - 1) * s - 2 * pad + k if transposed else (H + 2 * pad - k) // s + 1
    x = torch.randn(N, H, H, Ci, device="cuda").bfloat16()
    w = torch.randn(k, k, Co, Ci, device="cuda").bfloat16()
    if mn:   # dgrad: in has Co channels
        x = torch.randn(N, H, H, Co, device="cuda").bfloat16()
    y = torch.empty(N, Ho, Ho, Ci if mn else Co, device="cuda", dtype=torch.bfloat16)
    st = L.acc_zeros(2 * Co, "cuda") if stats else None
    fn(buf.ctypes.data, 1)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for r in range(reps + 1):
        if r == 1:
            fn(buf.ctypes.data, 1); e0.record()
        if not mn:
            L.check(lib.awr_conv_tc(x.data_ptr(), w.data_ptr(), None, y.data_ptr(), None if st is None else st.data_ptr(), N, H, H, Ci, Ho, Ho, Co, k, k, s, pad,
                                    transposed, 1, Ci, Co * Ci, 0, 0, 0, L.stream()), "conv")
        else:
            L.check(lib.awr_conv_tc(x.data_ptr(), w.data_ptr(), None, y.data_ptr(), None, N, H, H, Co, Ho, Ho, Ci, k, k, s, pad,
                                    transposed, Ci, 1, Co * Ci, 0, 0, 0, L.stream()), "conv")
    e1.record(); torch.cuda.synchronize()
    fn(buf.ctypes.data, 0)
    p = buf.reshape(148, 8).astype(np.float64) / reps
    act = p[:, 5] > 0
    if not act.any():
        print(f"{name:34s} {1e3 * e0.elapsed_time(e1) / reps:7.1f} us  (halo-tile kernel: see tools/dbg_halo_timeline.py)")
        return
    m = p[act].mean(axis=0)
    print(f"{name:34s} {1e3 * e0.elapsed_time(e1) / reps:7.1f} us | total {m[5]:8.0f} cyc  tiles {m[6]:.1f} kiters {m[7]:.0f} | prod wait-empty {m[0]:7.0f}  mma wait-full {m[1]:7.0f} "
          f"wait-tmem {m[2]:7.0f} | epi wait-full {m[3]:7.0f} busy {m[4]:7.0f}  (per tile: epi busy {m[4] / max(m[6], 1):6.0f}, per k-iter total {m[5] / max(m[7], 1):5.0f})")

run("layer1 fprop 3x3 64->64 @64 +stats", 32, 64, 64, 64, 3, 1, 1, stats=True)
run("layer1 fprop 3x3 64->64 @64", 32, 64, 64, 64, 3, 1, 1)
run("layer1 dgrad 3x3 64->64 @64", 32, 64, 64, 64, 3, 1, 1, transposed=1, mn=1)
run("layer2 fprop 3x3 128->128 @32", 32, 128, 128, 32, 3, 1, 1, stats=True)
run("layer3 fprop 3x3 256->256 @16", 32, 256, 256, 16, 3, 1, 1, stats=True)
run("layer4 fprop 3x3 512->512 @8", 32, 512, 512, 8, 3, 1, 1, stats=True)
run("deconv3 fprop 256->256 @32->64", 32, 256, 256, 32, 4, 2, 1, transposed=1, stats=True)
run("deconv3 fprop (no stats)", 32, 256, 256, 32, 4, 2, 1, transposed=1)
run("1x1 256->256 @64", 32, 256, 256, 64, 1, 1, 0)
