#!/bin/bash
set -u
O=gpurun_out; mkdir -p $O
timeout 400 python -m pytest tests -m gpu -x -q --timeout 120 > $O/hg_pytest.log 2>&1; echo "pytest exit $?" >> $O/hg_pytest.log; tail -4 $O/hg_pytest.log
timeout 200 python bench.py --steps 50 --cpu-steps 0 --net hourglass_1 > $O/hg_bench_hg1.json 2> $O/hg_bench_hg1.err; echo "bench hg1 $?"
timeout 200 python bench.py --steps 100 --cpu-steps 0 > $O/hg_bench_new.json 2> $O/hg_bench_new.err; echo "bench new $?"
for f in hg1 new; do python - <<PY
import json
try:
    d = json.loads(open("$O/hg_bench_$f.json").read().strip().splitlines()[-1])
    print("$f", d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], d["launches_per_step"], d["loss"])
    print({k: v["ms_per_step"] for k, v in d["kernel_classes"].items()})
except Exception as e:
    print("$f", "ERR", e)
PY
done
