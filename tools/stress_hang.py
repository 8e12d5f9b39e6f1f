"""Stress: repeat the bench legs that launch the tensor-core kernels back to back (PDL chains), to catch rare deadlocks.  usage: stress_hang.py [reps]"""
import faulthandler, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import awr_b200
from awr_b200.trainer import FusedTrainer
import bench
faulthandler.dump_traceback_later(int(os.environ.get("HANG_S", "240")), exit=True)
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 100
dev = torch.device("cuda", 0)
torch.manual_seed(1)
net = awr_b200.get_deconv_net(18, 14, 2, precision="bf16").to(dev)
tr = FusedTrainer(net, 32, 128, 1.0, 1.0, 1.0, use_graph=True)
img = torch.rand(32, 1, 128, 128, device=dev) * 2 - 1
jt = torch.rand(32, 14, 3, device=dev) - 0.5
for _ in range(5):
    tr.load_batch(img, jt); tr.run_step()
torch.cuda.synchronize()
t0 = time.time()
for r in range(reps):
    ms, fl, n = bench.conv_class_time(tr, reps=3)
    if r % 20 == 0:
        print(f"rep {r}: conv class {ms:.4f} ms ({time.time() - t0:.1f} s)", flush=True)
    for _ in range(3):
        tr.load_batch(img, jt); tr.run_step()
torch.cuda.synchronize()
print("stress ok", flush=True)
