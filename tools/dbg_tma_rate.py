"""Hardware probe: per-SM TMA throughput for 128-byte-row boxes (the conv kernels' operand tiles) by box size, ring depth, issue pattern
and data placement.  Run: make -C awr-adaptive-weighting-regression_b200 debug && gpurun -- python tools/dbg_tma_rate.py"""
import torch
import _dbglib
from awr_b200 import _lib as L
lib = _dbglib.lib()
g_rows = 1 << 20                      # 128 MiB matrix of 128-byte rows
G = torch.randn(g_rows, 64, device="cuda").bfloat16()
groups = 256
MODES = {0: "one issuing thread", 1: "two issuing threads", 2: "two tensor maps alternating", 3: "1-D bulk copies"}
for grid in (1, 148):
    for mode in (0, 1, 2, 3):
        for rows in (16, 64, 128, 256):
            for stages in (4,):
                for name, stride, span in (("same rows x9 (L2)", 0, 9), ("private rows streaming (HBM)", 7000, 1 << 20)):
                    out = torch.zeros(grid, dtype=torch.int64, device="cuda")
                    for _ in range(2):
                        L.check(lib.awr_debug_tma_rate(G.data_ptr(), g_rows, out.data_ptr(), rows, stages, groups, stride, span, mode, grid, L.stream()), "tma_rate")
                    torch.cuda.synchronize()
                    cyc = out.float().mean().item()
                    print(f"grid {grid:3d} {MODES[mode]:28s} box {rows:3d} rows x128B stages {stages}: {groups * rows * 128 / cyc:6.1f} B/clk/SM  ({cyc / groups:6.0f} cycles/box)  {name}")
