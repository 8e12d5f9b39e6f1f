"""Hardware probe: how far can the issuing thread run ahead of the tensor pipe?  For bursts of n MMAs (M=128, K=16, bf16) prints the cycles until
the LAST tcgen05.mma has been issued and until the burst has completed: issue time ~ 0 means the MMAs are queued, issue time ~ completion
means tcgen05.mma blocks the issuing thread (queue depth ~ where the two curves meet).  Run: make debug && gpurun -- python tools/dbg_umma_queue.py"""
import torch
import _dbglib
from awr_b200 import _lib as L
lib = _dbglib.lib()
for N in (64, 128, 256):
    for taps in (1, 2, 3, 4, 6, 8, 16, 32):
        out = torch.zeros(2, dtype=torch.int64, device="cuda")
        for _ in range(2):
            L.check(lib.awr_debug_umma_rate2(out.data_ptr(), N, 2304, 0, 128, taps, 1, L.stream()), "rate2")
        torch.cuda.synchronize()
        done, issued = out[0].item(), out[1].item()
        print(f"N {N:3d} burst {taps * 8:3d} MMAs: issued after {issued:6d} cycles, complete after {done:6d}  (ideal {taps * 8 * max(32, 128 * N // 256):6d} at the tensor rate)")
