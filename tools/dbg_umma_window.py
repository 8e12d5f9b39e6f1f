"""Hardware probe: row-shifted / strided windows of a SWIZZLE_128B smem tile as UMMA A operands (see csrc/debug_umma.cu)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from awr_b200 import _lib as L
import _dbglib
lib = _dbglib.lib()          # hardware probes live in libawr_b200_debug.so (make debug)
rows = 256
g = torch.Generator().manual_seed(0)
G = torch.randn(rows, 64, generator=g).bfloat16()
Bm = torch.randn(64, 64, generator=g).bfloat16()
Gd, Bd = G.cuda(), Bm.cuda()
for sbo in (8, 10, 16):
    for r0 in (0, 1, 3, 8, 13, 37):
        if r0 + 15 * sbo + 8 > rows:
            continue
        idx = torch.tensor([r0 + (m // 8) * sbo + (m % 8) for m in range(128)])
        ref = G[idx].float() @ Bm.float().t()
        out = []
        for mode in (0, 1):
            D = torch.full((128, 64), float("nan"), device="cuda")
            L.check(lib.awr_debug_umma_window(Gd.data_ptr(), Bd.data_ptr(), D.data_ptr(), rows, r0, sbo, mode, L.stream()), "probe")
            torch.cuda.synchronize()
            out.append((D.cpu() - ref).abs().max().item())
        print(f"sbo_rows {sbo:2d} r0 {r0:2d}: max|err| base_offset=0 -> {out[0]:.3e}   base_offset=(addr>>7)&7 -> {out[1]:.3e}   (ref max {ref.abs().max():.1f})")
