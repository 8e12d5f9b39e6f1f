#!/bin/bash
set -u
O=gpurun_out; mkdir -p $O
timeout 300 python -m pytest tests/test_trainer_gpu.py -x -q --timeout 120 > $O/ab5_pytest.log 2>&1; echo "pytest exit $?" >> $O/ab5_pytest.log; tail -6 $O/ab5_pytest.log
timeout 200 python bench.py --steps 200 --cpu-steps 0 > $O/ab5_bench_new.json 2> $O/ab5_bench_new.err; echo "bench new $?"
timeout 200 python bench.py --steps 50 --cpu-steps 0 --net hourglass_1 > $O/ab5_bench_hg1.json 2> $O/ab5_bench_hg1.err; echo "bench hg1 $?"
timeout 200 python bench.py --steps 50 --cpu-steps 0 --net resnet_50 > $O/ab5_bench_r50.json 2> $O/ab5_bench_r50.err; echo "bench r50 $?"
timeout 250 python bench.py --steps 20 --cpu-steps 0 --net resnet_50 --img-size 256 --batch 64 > $O/ab5_bench_r50_256.json 2> $O/ab5_bench_r50_256.err; echo "bench r50_256 $?"
for f in new hg1 r50 r50_256; do python - <<PY
import json
try:
    d = json.loads(open("$O/ab5_bench_$f.json").read().strip().splitlines()[-1])
    print("$f", d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], "conv TF", d["roofline"]["achieved"], d["launches_per_step"], d["loss"])
except Exception as e:
    print("$f", "ERR", e)
PY
done
tail -3 $O/ab5_bench_*.err
