#!/bin/bash
set -u
O=gpurun_out; mkdir -p $O
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/check_dp_overlap.py > $O/dp2_check.log 2>&1; echo "check $?"; tail -3 $O/dp2_check.log
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 200 --warmup 5 > $O/dp2_bench.json 2> $O/dp2_bench.err; echo "bench $?"
python - <<'PY'
import json
d = json.loads([l for l in open("gpurun_out/dp2_bench.json").read().strip().splitlines() if l.startswith("{")][-1])
print(d["n_gpus"], d["value"], d["ms_per_step"], d["e2e"], d["config"]["parallelism"])
PY
