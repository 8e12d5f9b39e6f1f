"""Run one tensor-core conv shape a few times (target for `ncu -k regex:conv_halo ...`).  usage: run_one_conv.py layer1|layer1d|layer2|layer3|layer4|deconv3 [reps]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from awr_b200 import _lib as L
lib = L.lib()
SHAPES = {"layer1": (32, 64, 64, 64, 3, 1, 1, 0, 0, True), "layer1d": (32, 64, 64, 64, 3, 1, 1, 1, 1, False), "layer2": (32, 128, 128, 32, 3, 1, 1, 0, 0, True),
          "layer3": (32, 256, 256, 16, 3, 1, 1, 0, 0, True), "layer4": (32, 512, 512, 8, 3, 1, 1, 0, 0, True), "deconv3": (32, 256, 256, 32, 4, 2, 1, 1, 0, True)}
N, Ci, Co, H, k, s, pad, transposed, mn, stats = SHAPES[sys.argv[1]]
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 4
Ho = (H - 1) * s - 2 * pad + k if transposed and not mn else ((H + 2 * pad - k) // s + 1)
x = torch.randn(N, H, H, Co if mn else Ci, device="cuda").bfloat16()
w = (torch.randn(k, k, Co, Ci, device="cuda") * 0.05).bfloat16()
y = torch.empty(N, Ho, Ho, Ci if mn else Co, device="cuda", dtype=torch.bfloat16)
st = L.acc_zeros(2 * Co, "cuda") if stats else None
for _ in range(reps):
    if not mn:
        L.check(lib.awr_conv_tc(x.data_ptr(), w.data_ptr(), None, y.data_ptr(), None if st is None else st.data_ptr(), N, H, H, Ci, Ho, Ho, Co, k, k, s, pad,
                                transposed, 1, Ci, Co * Ci, 0, 0, 0, L.stream()), "conv")
    else:
        L.check(lib.awr_conv_tc(x.data_ptr(), w.data_ptr(), None, y.data_ptr(), None, N, H, H, Co, Ho, Ho, Ci, k, k, s, pad,
                                transposed, Ci, 1, Co * Ci, 0, 0, 0, L.stream()), "conv")
torch.cuda.synchronize()
print("ok", y.float().abs().mean().item())
