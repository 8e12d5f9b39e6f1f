"""Hardware probe: cycles per tcgen05.mma (cta_group::1, M=128, K=16, bf16, SS operands) vs N and number of accumulators."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from awr_b200 import _lib as L
import _dbglib
lib = _dbglib.lib()          # hardware probes live in libawr_b200_debug.so (make debug)
for grid in (1, 148):
    for N in (64, 128, 256):
        for nacc in (1, 2, 4):
            if nacc * N > 512: continue
            for shift in (0, 3):
                out = torch.zeros(grid, dtype=torch.int64, device="cuda")
                iters = 2000
                L.check(lib.awr_debug_umma_rate(out.data_ptr(), N, nacc, iters, shift, grid, L.stream()), "rate")
                torch.cuda.synchronize()
                cyc = out.float().mean().item() / (iters * 4)
                print(f"grid {grid:3d} N {N:3d} nacc {nacc} a_row_shift {shift}: {cyc:6.1f} cycles/MMA  (ideal {128 * N / 256:.0f})  -> {2 * 128 * N * 16 / cyc:.0f} flop/clk/SM")
