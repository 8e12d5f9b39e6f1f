"""Hardware probe: where do cycles go in the MMA-issuer loop?  (commit per group, barrier wait per group, TMA-fed ring)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from awr_b200 import _lib as L
import _dbglib
lib = _dbglib.lib()          # hardware probes live in libawr_b200_debug.so (make debug)
G = torch.randn(65536, 64, device="cuda").bfloat16()
def run(N, per, mode, stages, rows, grid=148, groups=2000):
    out = torch.zeros(grid, dtype=torch.int64, device="cuda")
    L.check(lib.awr_debug_umma_pipe(G.data_ptr(), G.shape[0], out.data_ptr(), N, per, groups, mode, stages, rows, grid, L.stream()), "pipe")
    torch.cuda.synchronize()
    cyc = out.float().mean().item() / groups
    ideal = per * max(128 * N / 256, 48)
    print(f"N {N:3d} per {per} mode {mode} ({'commit ' if mode & 1 else ''}{'wait ' if mode & 2 else ''}{'tma ' + str(rows) + ' rows' if mode & 4 else ''}) stages {stages}: "
          f"{cyc:7.1f} cycles/group  (MMA-only {ideal:.0f})  TMA {rows * 128 / cyc if mode & 4 else 0:5.1f} B/clk")
for N in (64, 128):
    for per in (4, 8):
        for stages in (1, 2, 4):
            run(N, per, 0, stages, 64)
        run(N, per, 1, 4, 64)
        run(N, per, 3, 4, 64)
        run(N, per, 5, 4, 256)
