"""Hardware probe: tcgen05.mma rate (M=128, K=16, bf16, SS operands) with the halo-tile kernel's A-operand pattern -- window start not
1024-byte aligned, 8-row groups `sbo` bytes apart (the halo row pitch) -- against the dense canonical tile (sbo = 1024, start 0).
Run:  make -C awr-adaptive-weighting-regression_b200 debug && gpurun -- python tools/dbg_umma_rate2.py"""
import torch
import _dbglib
from awr_b200 import _lib as L
lib = _dbglib.lib()
iters = 2000
for grid in (1, 148):
    for N in (64, 128):
        for sbo, start, tap in ((1024, 0, 0), (1024, 384, 0), (1024, 0, 128), (2048, 0, 0), (2304, 0, 0), (2304, 0, 128), (2304, 128, 128), (2304, 2432, 128),
                                (2176, 0, 128), (3072, 0, 0), (3072, 0, 128), (2560, 0, 128)):
            out = torch.zeros(grid, dtype=torch.int64, device="cuda")
            L.check(lib.awr_debug_umma_rate2(out.data_ptr(), N, sbo, start, tap, iters, grid, L.stream()), "rate2")
            torch.cuda.synchronize()
            cyc = out.float().mean().item() / (iters * 8)
            print(f"grid {grid:3d} N {N:3d} sbo {sbo:5d} start {start:5d} tap_step {tap:4d}: {cyc:6.1f} cycles/MMA (ideal {128 * N / 256:.0f})")
