#!/bin/bash
set -u
O=gpurun_out; mkdir -p $O
timeout 200 python -m pytest tests/test_trainer_gpu.py -x -q --timeout 120 > $O/fin_pytest.log 2>&1; echo "pytest exit $?" >> $O/fin_pytest.log; tail -4 $O/fin_pytest.log
python - <<'PY' > gpurun_out/fin_h2d.log 2>&1
import torch, time
for mb in (2, 16, 64):
    h = torch.empty(mb * 2**20 // 4).pin_memory(); d = torch.empty_like(h, device="cuda")
    for _ in range(3): d.copy_(h, non_blocking=True)
    torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): d.copy_(h, non_blocking=True)
    e1.record(); torch.cuda.synchronize()
    print(mb, "MiB pinned H2D", e0.elapsed_time(e1) / 10, "ms ->", mb * 2**20 / (e0.elapsed_time(e1) / 10 * 1e-3) / 1e9, "GB/s")
PY
cat gpurun_out/fin_h2d.log
bash tools/profile_final.sh r01_final2
