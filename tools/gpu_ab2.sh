#!/bin/bash
set -u
O=gpurun_out; mkdir -p $O
timeout 300 python -m pytest tests -m gpu -x -q > $O/ab2_pytest.log 2>&1; echo "pytest exit $?" >> $O/ab2_pytest.log; tail -4 $O/ab2_pytest.log
timeout 200 python bench.py --steps 100 --cpu-steps 0 --layers ab2_layers_new.md > $O/ab2_bench_new.json 2> $O/ab2_bench_new.err; echo "bench new $?"
AWR_STEM_WGRAD=old timeout 200 python bench.py --steps 100 --cpu-steps 0 --no-parity > $O/ab2_bench_oldwgrad.json 2> $O/ab2_bench_oldwgrad.err; echo "bench oldwgrad $?"
for f in new oldwgrad; do python - <<PY
import json
try:
    d = json.loads(open("$O/ab2_bench_$f.json").read().strip().splitlines()[-1])
    print("$f", d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["achieved"], d["roofline_head"]["pair_us"], (d.get("parity") or {}).get("uvd_max_abs_diff"))
    print({k: v["ms_per_step"] for k, v in d["kernel_classes"].items()})
except Exception as e:
    print("$f", "ERR", e)
PY
done
